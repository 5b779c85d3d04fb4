"""PRIME (process reinforcement through implicit rewards, arXiv:2502.01456) building blocks: an implicit process reward model is a
language model whose token-level reward is ``beta * (log pi_prm - log pi_ref)``; it is trained online with a sequence-level
cross-entropy against the outcome label, and its dense rewards are combined with the outcome reward through leave-one-out baselines.
Counterpart of the reference recipe's core algorithms (3rdparty/verl/recipe/prime/prime_core_algos.py:20-150); the model / trainer
wiring is the generic one (a ZeRO engine for the PRM, ``data/reward_manager.py::PrimeRewardManager`` for the outcome scores)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .core_algos import masked_whiten


def implicit_process_rewards(prm_log_prob: torch.Tensor, ref_log_prob: torch.Tensor, response_mask: torch.Tensor, beta: float = 0.05,
                             granularity: str = "token") -> torch.Tensor:
    """[b, R] token-level rewards ``beta * log(pi_prm / pi_ref)``; ``granularity='whole'`` puts the sequence sum on the last token."""
    q = (prm_log_prob - ref_log_prob) * response_mask * beta
    if granularity == "token":
        return q
    if granularity != "whole":
        raise ValueError("granularity must be 'token' or 'whole'")
    out = torch.zeros_like(q)
    last = (response_mask.sum(-1).long() - 1).clamp(min=0)
    out[torch.arange(q.shape[0]), last] = q.sum(-1)
    return out * (response_mask.sum(-1, keepdim=True) > 0)


def prm_ce_loss(token_level_scores: torch.Tensor, acc: torch.Tensor, response_mask: torch.Tensor, beta: float = 0.05) -> torch.Tensor:
    """Train the implicit PRM: sigmoid(beta * sum_t log-ratio) should predict the outcome label ``acc`` in [0, 1]."""
    logit = (token_level_scores * response_mask).sum(-1) * beta
    return F.binary_cross_entropy_with_logits(logit, acc.to(logit.dtype))


def _rloo(values: torch.Tensor, n: int) -> torch.Tensor:
    """Leave-one-out baseline inside groups of ``n`` consecutive rows: v_i * n/(n-1) - sum_group / (n-1)."""
    g = values.view(-1, n)
    return (g * (n / (n - 1)) - g.sum(-1, keepdim=True) / (n - 1)).reshape(-1)


def prime_advantages(response_mask: torch.Tensor, n_samples: int, acc: torch.Tensor | None = None, process_rewards: torch.Tensor | None = None,
                     reward_gt_coef: float = 5.0, reward_dpo_coef: float = 5.0) -> tuple[torch.Tensor, torch.Tensor]:
    """(advantages, returns) [b, R]: the outcome reward (on the last token) and the dense implicit rewards each get their own
    leave-one-out baseline over the ``n_samples`` responses of a prompt (rows adjacent), are summed with their coefficients, turned
    into returns-to-go and whitened."""
    if n_samples < 2:
        raise ValueError("leave-one-out baselines need at least 2 samples per prompt")
    mask = response_mask.float()
    total = torch.zeros_like(mask)
    if process_rewards is not None and reward_dpo_coef != 0.0:
        r = process_rewards * mask
        mean_per_seq = r.sum(-1) / mask.sum(-1).clamp(min=1)                   # baseline from the per-token mean reward of the others
        base = (mean_per_seq.view(-1, n_samples).sum(-1, keepdim=True) / (n_samples - 1)).expand(-1, n_samples).reshape(-1, 1)
        total = total + (r * (n_samples / (n_samples - 1)) - base * mask) * reward_dpo_coef
    if acc is not None and reward_gt_coef != 0.0:
        last = (mask.sum(-1).long() - 1).clamp(min=0)
        out = torch.zeros_like(mask)
        out[torch.arange(mask.shape[0]), last] = _rloo(acc.float(), n_samples)
        total = total + out * (mask.sum(-1, keepdim=True) > 0) * reward_gt_coef
    returns = (total * mask).flip(-1).cumsum(-1).flip(-1)
    return masked_whiten(returns, mask), returns


def pairwise_accuracy(token_level_scores: torch.Tensor, acc: torch.Tensor, response_mask: torch.Tensor, n_samples: int) -> float:
    """How often the implicit sequence reward orders two samples of a prompt like their outcome labels (0.5 when no pair differs)."""
    s = (token_level_scores * response_mask).sum(-1).view(-1, n_samples)
    a = acc.float().view(-1, n_samples)
    ds, da = s.unsqueeze(2) - s.unsqueeze(1), a.unsqueeze(2) - a.unsqueeze(1)
    iu = torch.triu(torch.ones(n_samples, n_samples, dtype=torch.bool), diagonal=1)
    ds, da = ds[:, iu], da[:, iu]
    w = da.abs()
    if float(w.sum()) == 0.0:
        return 0.5
    return float((((ds > 0) == (da > 0)).float() * w).sum() / w.sum())


class ImplicitPRM:
    """The implicit process reward model as a trainable component: a policy-shaped LM (``engine``, a ZeRO engine) scored against a
    frozen reference (``ref_engine``).  ``score`` returns dense token rewards for a rollout batch, ``update`` takes one optimizer
    step of the sequence-level cross-entropy against the outcome labels ``acc`` (reference: recipe/prime/prime_dp_rm.py — compute_rm_score /
    update_rm).  Micro-batching and packing are the PPO actor's."""

    def __init__(self, engine, ref_engine, beta: float = 0.05, granularity: str = "token", micro_batch_size: int = 8):
        from ..workers.actor import ActorConfig, DataParallelPPOActor

        self.beta, self.granularity = beta, granularity
        rt = ActorConfig(ppo_mini_batch_size=micro_batch_size, ppo_micro_batch_size_per_gpu=micro_batch_size)
        self.prm = DataParallelPPOActor(engine, rt)
        self.ref = DataParallelPPOActor(ref_engine, rt)
        self.engine = engine

    @torch.no_grad()
    def _ref_log_prob(self, data):
        lp, _ = self.ref.compute_log_prob(data, calculate_entropy=False)
        return lp

    @torch.no_grad()
    def score(self, data) -> torch.Tensor:
        """[b, R] implicit process rewards of the batch's responses."""
        lp, _ = self.prm.compute_log_prob(data, calculate_entropy=False)
        return implicit_process_rewards(lp, self._ref_log_prob(data), data.batch["response_mask"].float(), self.beta, self.granularity)

    def update(self, data, acc: torch.Tensor) -> dict:
        """One optimizer step on the whole batch (gradients accumulated over micro-batches)."""
        ref_lp = self._ref_log_prob(data)
        n = len(data)
        mbs = self.prm._micro_batches(data, self.prm.config.ppo_micro_batch_size_per_gpu, None, False)
        total = 0.0
        for mb, idx in mbs:
            rows = torch.tensor(idx, dtype=torch.long, device=ref_lp.device)
            _, lp = self.prm._forward_micro_batch(mb, 1.0, calculate_entropy=False)
            mask = mb.batch["response_mask"].float()
            loss = prm_ce_loss(lp - ref_lp[rows], acc[rows].to(lp.device), mask, self.beta) * (len(idx) / n)
            loss.backward()
            total += float(loss.detach())
        st = self.engine.step()
        return {"reward_model/prm_ce_loss": total, "reward_model/grad_norm": st["grad_norm"]}
