"""PPO / GRPO family algorithms: advantage estimators, KL controllers, loss aggregation, policy losses,
value loss and KL penalties.

Capability parity with the reference's ``core_algos`` (3rdparty/verl/verl/trainer/ppo/core_algos.py:
estimators :212-755, KL controllers :150-209, agg_loss :772-809, policy losses :887-1366, value loss
:1370-1409, kl_penalty :1412-1473); formulas per SURVEY Appendix E.  Group statistics are vectorised
with ``torch.unique`` + ``index_add`` instead of per-sample Python dict loops.
"""
from __future__ import annotations

from enum import Enum
from typing import Callable

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "AdvantageEstimator", "register_adv_est", "get_adv_estimator_fn", "register_policy_loss", "get_policy_loss_fn",
    "AdaptiveKLController", "FixedKLController", "get_kl_controller", "masked_sum", "masked_mean", "masked_var",
    "masked_whiten", "agg_loss", "compute_policy_loss_vanilla", "compute_value_loss", "kl_penalty",
    "compute_gae_advantage_return", "compute_grpo_outcome_advantage", "compute_rewards",
]


# ---------------------------------------------------------------------------------------------
# masked statistics (reference: 3rdparty/verl/verl/utils/torch_functional.py:163-223)
# ---------------------------------------------------------------------------------------------
def masked_sum(x: torch.Tensor, mask: torch.Tensor, dim=None) -> torch.Tensor:
    return (x * mask).sum() if dim is None else (x * mask).sum(dim)


def masked_mean(x: torch.Tensor, mask: torch.Tensor, dim=None) -> torch.Tensor:
    if dim is None:
        return (x * mask).sum() / (mask.sum() + 1e-8)
    return (x * mask).sum(dim) / (mask.sum(dim) + 1e-8)


def masked_var(x: torch.Tensor, mask: torch.Tensor, unbiased: bool = True) -> torch.Tensor:
    mean = masked_mean(x, mask)
    var = masked_mean((x - mean) ** 2, mask)
    if unbiased:
        n = mask.sum()
        if n.item() == 0:
            raise ValueError("masked_var: mask is empty")
        if n.item() == 1:
            raise ValueError("masked_var: a single valid element has no unbiased variance")
        var = var * (n / (n - 1))
    return var


def masked_whiten(x: torch.Tensor, mask: torch.Tensor, shift_mean: bool = True) -> torch.Tensor:
    mean, var = masked_mean(x, mask), masked_var(x, mask)
    out = (x - mean) * torch.rsqrt(var + 1e-8)
    return out if shift_mean else out + mean


# ---------------------------------------------------------------------------------------------
# registries
# ---------------------------------------------------------------------------------------------
class AdvantageEstimator(str, Enum):
    GAE = "gae"
    GRPO = "grpo"
    GRPO_VECTORIZED = "grpo_vectorized"
    GRPO_PASSK = "grpo_passk"
    REINFORCE_PLUS_PLUS = "reinforce_plus_plus"
    REINFORCE_PLUS_PLUS_BASELINE = "reinforce_plus_plus_baseline"
    REMAX = "remax"
    RLOO = "rloo"
    RLOO_VECTORIZED = "rloo_vectorized"
    OPO = "opo"
    GPG = "gpg"


_ADV: dict[str, Callable] = {}
_LOSS: dict[str, Callable] = {}


def _name(x) -> str:
    return x.value if isinstance(x, Enum) else str(x)


def register_adv_est(name):
    def deco(fn):
        key = _name(name)
        if key in _ADV and _ADV[key] is not fn:
            raise ValueError(f"advantage estimator {key!r} already registered")
        _ADV[key] = fn
        return fn

    return deco


def get_adv_estimator_fn(name) -> Callable:
    key = _name(name)
    if key not in _ADV:
        raise ValueError(f"unknown advantage estimator {key!r}; have {sorted(_ADV)}")
    return _ADV[key]


def register_policy_loss(name: str):
    def deco(fn):
        _LOSS[name] = fn
        return fn

    return deco


def get_policy_loss_fn(name: str) -> Callable:
    if name not in _LOSS:
        raise ValueError(f"unknown policy loss {name!r}; have {sorted(_LOSS)}")
    return _LOSS[name]


# ---------------------------------------------------------------------------------------------
# KL controllers
# ---------------------------------------------------------------------------------------------
class AdaptiveKLController:
    """Proportional controller of Ziegler et al. (arXiv:1909.08593)."""

    def __init__(self, init_kl_coef: float, target_kl: float, horizon: float):
        self.value, self.target, self.horizon = init_kl_coef, target_kl, horizon

    def update(self, current_kl: float, n_steps: int) -> None:
        err = float(np.clip(current_kl / self.target - 1.0, -0.2, 0.2))
        self.value *= 1.0 + err * n_steps / self.horizon


class FixedKLController:
    def __init__(self, kl_coef: float):
        self.value = kl_coef

    def update(self, current_kl: float, n_steps: int) -> None:
        return None


def get_kl_controller(kl_ctrl) -> AdaptiveKLController | FixedKLController:
    typ = kl_ctrl["type"] if isinstance(kl_ctrl, dict) else kl_ctrl.type
    get = (lambda k: kl_ctrl[k]) if isinstance(kl_ctrl, dict) else (lambda k: getattr(kl_ctrl, k))
    if typ == "fixed":
        return FixedKLController(get("kl_coef"))
    if typ == "adaptive":
        if get("horizon") <= 0:
            raise ValueError("adaptive KL controller needs horizon > 0")
        return AdaptiveKLController(get("kl_coef"), get("target_kl"), get("horizon"))
    raise NotImplementedError(f"kl controller type {typ!r}")


# ---------------------------------------------------------------------------------------------
# group helpers
# ---------------------------------------------------------------------------------------------
def _group_ids(index, device) -> tuple[torch.Tensor, int]:
    """Map arbitrary hashable group labels (uids) to dense ids 0..G-1."""
    arr = np.asarray(index)
    _, inv = np.unique(arr.astype(str) if arr.dtype == object else arr, return_inverse=True)
    inv_t = torch.as_tensor(inv.reshape(-1), device=device, dtype=torch.long)
    return inv_t, int(inv_t.max().item()) + 1 if inv_t.numel() else 0


def _group_mean_std(scores: torch.Tensor, gid: torch.Tensor, g: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-sample group mean / unbiased std / count; singleton groups get mean 0, std 1."""
    cnt = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, torch.ones_like(scores))
    s = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, scores)
    mean = s / cnt.clamp_min(1)
    sq = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, (scores - mean[gid]) ** 2)
    std = torch.sqrt(sq / (cnt - 1).clamp_min(1))
    single = cnt <= 1
    mean = torch.where(single, torch.zeros_like(mean), mean)
    std = torch.where(single, torch.ones_like(std), std)
    return mean[gid], std[gid], cnt[gid]


# ---------------------------------------------------------------------------------------------
# advantage estimators.  All return (advantages, returns), shape [bs, response_len].
# ---------------------------------------------------------------------------------------------
@register_adv_est(AdvantageEstimator.GAE)
@torch.no_grad()
def compute_gae_advantage_return(token_level_rewards, values, response_mask, gamma, lam, **_):
    t_len = token_level_rewards.shape[-1]
    next_v = torch.zeros_like(values[:, 0])
    last = torch.zeros_like(values[:, 0])
    adv_rev = []
    for t in reversed(range(t_len)):
        m = response_mask[:, t]
        delta = token_level_rewards[:, t] + gamma * next_v - values[:, t]
        cand = delta + gamma * lam * last
        # masked positions carry the running state through unchanged
        next_v = values[:, t] * m + (1 - m) * next_v
        last = cand * m + (1 - m) * last
        adv_rev.append(last)
    adv = torch.stack(adv_rev[::-1], dim=1)
    returns = adv + values
    adv = masked_whiten(adv, response_mask)
    return adv, returns


@register_adv_est(AdvantageEstimator.GRPO)
@torch.no_grad()
def compute_grpo_outcome_advantage(token_level_rewards, response_mask, index, epsilon: float = 1e-6,
                                   norm_adv_by_std_in_grpo: bool = True, config=None, **_):
    scores = token_level_rewards.sum(-1)
    gid, g = _group_ids(index, scores.device)
    mean, std, _ = _group_mean_std(scores, gid, g)
    a = (scores - mean) / (std + epsilon) if norm_adv_by_std_in_grpo else scores - mean
    a = a.unsqueeze(-1) * response_mask
    return a, a


register_adv_est(AdvantageEstimator.GRPO_VECTORIZED)(compute_grpo_outcome_advantage)


@register_adv_est(AdvantageEstimator.GRPO_PASSK)
@torch.no_grad()
def compute_grpo_passk_outcome_advantage(token_level_rewards, response_mask, index, epsilon: float = 1e-6,
                                         norm_adv_by_std_in_grpo: bool = True, config=None, **_):
    """Pass@k: only the best sample of a group gets (r_max - r_second)[/std]."""
    if config is not None:
        norm_adv_by_std_in_grpo = (config.get("norm_adv_by_std_in_grpo", True) if hasattr(config, "get")
                                   else getattr(config, "norm_adv_by_std_in_grpo", True))
    scores = token_level_rewards.sum(-1)
    gid, g = _group_ids(index, scores.device)
    adv = torch.zeros_like(scores)
    for k in range(g):
        members = (gid == k).nonzero(as_tuple=True)[0]
        if members.numel() < 2:
            raise ValueError("Pass@k needs at least 2 samples per group")
        r = scores[members]
        top, top_i = torch.topk(r, 2)
        a = top[0] - top[1]
        if norm_adv_by_std_in_grpo:
            a = a / (r.std() + epsilon)
        adv[members[top_i[0]]] = a
    adv = adv.unsqueeze(-1) * response_mask
    return adv, adv


@register_adv_est(AdvantageEstimator.REINFORCE_PLUS_PLUS_BASELINE)
@torch.no_grad()
def compute_reinforce_plus_plus_baseline_outcome_advantage(token_level_rewards, response_mask, index, epsilon=1e-6,
                                                           config=None, **_):
    scores = token_level_rewards.sum(-1)
    gid, g = _group_ids(index, scores.device)
    mean, _, _ = _group_mean_std(scores, gid, g)
    a = (scores - mean).unsqueeze(-1).expand_as(response_mask) * response_mask
    a = masked_whiten(a, response_mask) * response_mask
    return a, a


@register_adv_est(AdvantageEstimator.RLOO)
@torch.no_grad()
def compute_rloo_outcome_advantage(token_level_rewards, response_mask, index, epsilon=1e-6, config=None, **_):
    scores = token_level_rewards.sum(-1)
    gid, g = _group_ids(index, scores.device)
    cnt = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, torch.ones_like(scores))[gid]
    tot = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, scores)[gid]
    # n/(n-1) * (r - mean) == r - mean of the OTHER samples; singleton groups keep their raw score
    loo = (cnt * scores - tot) / (cnt - 1).clamp_min(1)
    a = torch.where(cnt > 1, loo, scores)
    a = a.unsqueeze(-1) * response_mask
    return a, a


@register_adv_est(AdvantageEstimator.RLOO_VECTORIZED)
@torch.no_grad()
def compute_rloo_vectorized_outcome_advantage(token_level_rewards, response_mask, index, epsilon=1e-6, config=None, **_):
    scores = token_level_rewards.sum(-1)
    gid, g = _group_ids(index, scores.device)
    cnt = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, torch.ones_like(scores))[gid]
    tot = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, scores)[gid]
    a = ((cnt * scores - tot) / (cnt - 1).clamp_min(1)) * (cnt > 1)
    a = a.unsqueeze(-1) * response_mask
    return a, a


@register_adv_est(AdvantageEstimator.OPO)
@torch.no_grad()
def compute_opo_outcome_advantage(token_level_rewards, response_mask, index, epsilon=1e-6, config=None, **_):
    """Length-weighted group baseline (arXiv:2505.23585)."""
    lens = response_mask.sum(-1).to(token_level_rewards.dtype)
    scores = token_level_rewards.sum(-1)
    gid, g = _group_ids(index, scores.device)
    cnt = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, torch.ones_like(scores))
    num = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, lens * scores)
    den = torch.zeros(g, device=scores.device, dtype=scores.dtype).index_add_(0, gid, lens)
    base = torch.where(cnt > 1, num / den.clamp_min(1e-8), torch.zeros_like(num))
    a = (scores - base[gid]).unsqueeze(-1) * response_mask
    return a, a


@register_adv_est(AdvantageEstimator.REINFORCE_PLUS_PLUS)
@torch.no_grad()
def compute_reinforce_plus_plus_outcome_advantage(token_level_rewards, response_mask, config=None, gamma=None, **_):
    if gamma is None:
        gamma = config.gamma if config is not None and hasattr(config, "gamma") else 1.0
    returns = torch.zeros_like(token_level_rewards)
    running = torch.zeros_like(token_level_rewards[:, 0])
    for t in reversed(range(token_level_rewards.shape[1])):
        running = token_level_rewards[:, t] + gamma * running
        returns[:, t] = running
        running = running * response_mask[:, t]      # reset after the sequence ends
    adv = masked_whiten(returns, response_mask) * response_mask
    return adv, returns


@register_adv_est(AdvantageEstimator.REMAX)
@torch.no_grad()
def compute_remax_outcome_advantage(token_level_rewards, reward_baselines, response_mask, config=None, **_):
    returns = (token_level_rewards * response_mask).flip(-1).cumsum(-1).flip(-1)
    adv = returns - reward_baselines.unsqueeze(-1) * response_mask
    return adv, returns


@register_adv_est(AdvantageEstimator.GPG)
@torch.no_grad()
def compute_gpg_outcome_advantage(token_level_rewards, response_mask, index, epsilon=1e-6, f_norm: float = 1.0,
                                  alpha: float = 1.0, config=None, **_):
    scores = token_level_rewards.sum(-1)
    bsz = scores.shape[0]
    alpha_t = bsz / torch.count_nonzero(scores).clamp(min=1)
    gid, g = _group_ids(index, scores.device)
    mean, _, _ = _group_mean_std(scores, gid, g)
    a = alpha_t * (scores - mean) / f_norm
    a = a.unsqueeze(-1) * response_mask
    return a, a


def compute_rewards(token_level_scores, old_log_prob, ref_log_prob, kl_ratio):
    return token_level_scores - (old_log_prob - ref_log_prob) * kl_ratio


# ---------------------------------------------------------------------------------------------
# loss aggregation
# ---------------------------------------------------------------------------------------------
def agg_loss(loss_mat: torch.Tensor, loss_mask: torch.Tensor, loss_agg_mode: str) -> torch.Tensor:
    if loss_agg_mode == "token-mean":
        return masked_mean(loss_mat, loss_mask)
    # sequence means run over sequences that still have loss tokens: fully masked rows (padded duplicates, DAPO-filtered
    # groups, rollout-correction vetoes) must not dilute the mean (reference: verl core_algos.agg_loss seq_mask)
    if loss_agg_mode == "seq-mean-token-sum":
        seq_mask = (loss_mask.sum(-1) > 0).to(loss_mat.dtype)
        return masked_mean((loss_mat * loss_mask).sum(-1), seq_mask)
    if loss_agg_mode == "seq-mean-token-mean":
        seq_mask = (loss_mask.sum(-1) > 0).to(loss_mat.dtype)
        return masked_mean((loss_mat * loss_mask).sum(-1) / (loss_mask.sum(-1) + 1e-8), seq_mask)
    if loss_agg_mode == "seq-mean-token-sum-norm":
        return (loss_mat * loss_mask).sum(-1).sum() / loss_mask.shape[-1]
    raise ValueError(f"invalid loss_agg_mode {loss_agg_mode!r}")


def _cfg(config, name, default=None):
    if config is None:
        return default
    if isinstance(config, dict):
        v = config.get(name, default)
    else:
        v = getattr(config, name, default)
    return default if v is None else v


def _clip_bounds(config, cliprange=None, lo=None, hi=None):
    c = cliprange if cliprange is not None else _cfg(config, "clip_ratio", 0.2)
    lo = lo if lo is not None else _cfg(config, "clip_ratio_low", c)
    hi = hi if hi is not None else _cfg(config, "clip_ratio_high", c)
    return c, lo, hi


# ---------------------------------------------------------------------------------------------
# policy losses.  Signature: (old_log_prob, log_prob, advantages, response_mask, loss_agg_mode,
#                              config, rollout_is_weights) -> (loss, metrics dict of floats)
# ---------------------------------------------------------------------------------------------
@register_policy_loss("vanilla")
def compute_policy_loss_vanilla(old_log_prob, log_prob, advantages, response_mask, loss_agg_mode="token-mean",
                                config=None, rollout_is_weights=None, cliprange=None, cliprange_low=None,
                                cliprange_high=None, clip_ratio_c=None):
    """Dual-clip PPO (arXiv:1912.09729)."""
    _, lo, hi = _clip_bounds(config, cliprange, cliprange_low, cliprange_high)
    c = clip_ratio_c if clip_ratio_c is not None else _cfg(config, "clip_ratio_c", 3.0)
    if c <= 1.0:
        raise ValueError(f"clip_ratio_c must be > 1, got {c}")
    neg_kl = torch.clamp(log_prob - old_log_prob, -20.0, 20.0)
    ratio = torch.exp(neg_kl)
    ppo_kl = masked_mean(-neg_kl, response_mask)
    l1 = -advantages * ratio
    l2 = -advantages * torch.clamp(ratio, 1 - lo, 1 + hi)
    clip1 = torch.maximum(l1, l2)
    clipfrac = masked_mean((l2 > l1).float(), response_mask)
    l3 = -advantages * c
    clip2 = torch.minimum(l3, clip1)
    clipfrac_lower = masked_mean(((clip1 > l3) & (advantages < 0)).float(), response_mask)
    losses = torch.where(advantages < 0, clip2, clip1)
    if rollout_is_weights is not None:
        losses = losses * rollout_is_weights
    loss = agg_loss(losses, response_mask, loss_agg_mode)
    return loss, {"actor/pg_clipfrac": clipfrac.detach().item(), "actor/ppo_kl": ppo_kl.detach().item(),
                  "actor/pg_clipfrac_lower": clipfrac_lower.detach().item()}


@register_policy_loss("gspo")
def compute_policy_loss_gspo(old_log_prob, log_prob, advantages, response_mask, loss_agg_mode="seq-mean-token-mean",
                             config=None, rollout_is_weights=None):
    """Sequence-level importance ratio (arXiv:2507.18071)."""
    _, lo, hi = _clip_bounds(config)
    neg_kl = log_prob - old_log_prob
    seq_len = response_mask.sum(-1).clamp(min=1)
    seq_kl = (neg_kl * response_mask).sum(-1) / seq_len
    log_ratio = torch.clamp(log_prob - log_prob.detach() + seq_kl.detach().unsqueeze(-1), max=10.0)
    ratio = torch.exp(log_ratio)
    l1 = -advantages * ratio
    l2 = -advantages * torch.clamp(ratio, 1 - lo, 1 + hi)
    losses = torch.maximum(l1, l2)
    if rollout_is_weights is not None:
        losses = losses * rollout_is_weights
    loss = agg_loss(losses, response_mask, "seq-mean-token-mean")
    return loss, {"actor/pg_clipfrac": masked_mean((l2 > l1).float(), response_mask).item(),
                  "actor/ppo_kl": masked_mean(-neg_kl, response_mask).item(), "actor/pg_clipfrac_lower": 0.0}


@register_policy_loss("gpg")
def compute_policy_loss_gpg(old_log_prob, log_prob, advantages, response_mask, loss_agg_mode="token-mean",
                            config=None, rollout_is_weights=None):
    losses = -log_prob * advantages
    if rollout_is_weights is not None:
        losses = losses * rollout_is_weights
    return agg_loss(losses, response_mask, loss_agg_mode), {}


@register_policy_loss("cispo")
def compute_policy_loss_cispo(old_log_prob, log_prob, advantages, response_mask, loss_agg_mode="token-mean",
                              config=None, rollout_is_weights=None):
    """CISPO (MiniMax-M1, arXiv:2506.13585): REINFORCE weighted by the CLIPPED, stop-gradient importance ratio — every token keeps
    a gradient (no PPO-style zeroing of clipped tokens), the clip only bounds its weight."""
    _, lo, hi = _clip_bounds(config)
    neg_kl = torch.clamp(log_prob - old_log_prob, -20.0, 20.0)
    ratio = torch.exp(neg_kl)
    w = torch.clamp(ratio, 1 - lo, 1 + hi).detach()
    losses = -w * advantages * log_prob
    if rollout_is_weights is not None:
        losses = losses * rollout_is_weights
    clipped = ((ratio < 1 - lo) | (ratio > 1 + hi)).float()
    return agg_loss(losses, response_mask, loss_agg_mode), {"actor/pg_clipfrac": masked_mean(clipped, response_mask).item(),
                                                            "actor/ppo_kl": masked_mean(-neg_kl, response_mask).item()}


@register_policy_loss("sppo")
def compute_policy_loss_sppo(old_log_prob, log_prob, advantages, response_mask, loss_agg_mode="token-mean",
                             config=None, rollout_is_weights=None):
    """SPPO (self-play preference optimisation, arXiv:2405.00675; reference recipe/sppo/dp_actor.py:34-57): regress the sequence
    log-ratio onto ``eta * reward``: ``(sum_t (log pi - log pi_old) - eta * R)^2`` averaged over non-empty samples.  The per-sequence
    reward is read from ``advantages`` (its masked mean, e.g. an outcome score minus 1/2 broadcast over the response)."""
    eta = _policy_loss_cfg(config, "sppo_eta", 1.0)
    mask = response_mask.float()
    log_ratio = ((log_prob - old_log_prob) * mask).sum(-1)
    reward = (advantages * mask).sum(-1) / mask.sum(-1).clamp(min=1)
    have = (mask.sum(-1) > 0).float()
    loss = (((log_ratio - eta * reward) ** 2) * have).sum() / have.sum().clamp(min=1)
    return loss, {"actor/sppo_log_ratio": (log_ratio * have).sum().item() / max(1.0, have.sum().item()),
                  "actor/ppo_kl": masked_mean(old_log_prob - log_prob, response_mask).item()}


def pair_preferences(scores: torch.Tensor) -> torch.Tensor:
    """Rows come in consecutive pairs (two samples of one prompt); True marks the preferred row of each pair (ties: the first)."""
    if scores.shape[0] % 2:
        raise ValueError("pairwise preference needs an even number of rows (rollout.n = 2, samples of a prompt adjacent)")
    s = scores.view(-1, 2)
    first = s[:, 0] >= s[:, 1]
    return torch.stack([first, ~first], dim=1).reshape(-1)


def dpo_loss(pi_chosen, pi_rejected, ref_chosen, ref_rejected, beta: float = 0.1, label_smoothing: float = 0.0,
             loss_type: str = "sigmoid") -> torch.Tensor:
    """Per-pair DPO (sigmoid, optional label smoothing) or IPO loss on sequence log-probabilities."""
    logits = (pi_chosen - pi_rejected) - (ref_chosen - ref_rejected)
    if loss_type == "sigmoid":
        return -F.logsigmoid(beta * logits) * (1 - label_smoothing) - F.logsigmoid(-beta * logits) * label_smoothing
    if loss_type == "ipo":
        return (logits - 1 / (2 * beta)) ** 2
    raise ValueError(f"unsupported dpo loss_type {loss_type!r} (sigmoid | ipo)")


@register_policy_loss("online_dpo")
def compute_policy_loss_online_dpo(old_log_prob, log_prob, advantages, response_mask, loss_agg_mode="token-mean",
                                   config=None, rollout_is_weights=None):
    """Online DPO / SPIN-style self-play (reference recipe/spin/core_algos.py:59-158): the two samples of every prompt form a pair,
    the one with the larger sequence score (read from ``advantages``) is 'chosen'; the reference policy of the DPO logit is the
    behaviour policy of this iteration (``old_log_prob``).  Needs rollout.n = 2 and whole pairs per micro-batch."""
    beta = _policy_loss_cfg(config, "dpo_beta", 0.1)
    mask = response_mask.float()
    seq_lp, seq_old = (log_prob * mask).sum(-1), (old_log_prob * mask).sum(-1)
    score = (advantages * mask).sum(-1) / mask.sum(-1).clamp(min=1)
    chosen = pair_preferences(score)
    losses = dpo_loss(seq_lp[chosen], seq_lp[~chosen], seq_old[chosen], seq_old[~chosen], beta,
                      _policy_loss_cfg(config, "dpo_label_smoothing", 0.0), _policy_loss_cfg(config, "dpo_loss_type", "sigmoid"))
    decided = (score[chosen] != score[~chosen]).float()                  # tied pairs carry no preference
    loss = (losses * decided).sum() / decided.sum().clamp(min=1)
    margin = ((seq_lp[chosen] - seq_old[chosen]) - (seq_lp[~chosen] - seq_old[~chosen])).detach()
    return loss, {"actor/dpo_margin": margin.mean().item(), "actor/dpo_acc": (margin > 0).float().mean().item(),
                  "actor/ppo_kl": masked_mean(old_log_prob - log_prob, response_mask).item()}


def _policy_loss_cfg(config, name, default):
    pl = _cfg(config, "policy_loss", None)
    return _cfg(pl, name, default)


@register_policy_loss("clip_cov")
def compute_policy_loss_clip_cov(old_log_prob, log_prob, advantages, response_mask, loss_agg_mode="token-mean",
                                 config=None, rollout_is_weights=None):
    """Clip-Cov (Entropy-Mechanism-of-RL): randomly detach a small share of high-covariance tokens."""
    ratio_cfg = _policy_loss_cfg(config, "clip_cov_ratio", 0.0002)
    ub = _policy_loss_cfg(config, "clip_cov_ub", 5.0)
    lb = _policy_loss_cfg(config, "clip_cov_lb", 1.0)
    _, lo, hi = _clip_bounds(config)
    neg_kl = log_prob - old_log_prob
    ratio = torch.exp(neg_kl)
    l1 = -advantages * ratio
    l2 = -advantages * torch.clamp(ratio, 1 - lo, 1 + hi)
    valid = response_mask > 0
    clipped = (l2 > l1) & valid
    cov = (advantages - masked_mean(advantages, response_mask)) * (log_prob - masked_mean(log_prob.detach(), response_mask))
    cov = cov.detach().masked_fill(~valid | clipped, float("-inf"))
    n_clip = max(int(ratio_cfg * response_mask.sum().item()), 1)
    cand = ((cov < ub) & (cov > lb) & valid).nonzero()
    corr = torch.ones_like(advantages)
    if cand.shape[0] > 0:
        pick = cand[torch.randperm(cand.shape[0], device=cand.device)[: min(n_clip, cand.shape[0])]]
        corr[pick[:, 0], pick[:, 1]] = 0
    losses = torch.maximum(l1, l2) * corr
    if rollout_is_weights is not None:
        losses = losses * rollout_is_weights
    loss = agg_loss(losses, response_mask, loss_agg_mode)
    return loss, {"actor/pg_clipfrac": masked_mean((corr == 0).float(), response_mask).item(),
                  "actor/ppo_kl": masked_mean(-neg_kl, response_mask).item()}


@register_policy_loss("kl_cov")
def compute_policy_loss_kl_cov(old_log_prob, log_prob, advantages, response_mask, loss_agg_mode="token-mean",
                               config=None, rollout_is_weights=None):
    """KL-Cov: add a |KL| penalty on the top-k covariance tokens."""
    k_ratio = _policy_loss_cfg(config, "kl_cov_ratio", 0.0002)
    coef = _policy_loss_cfg(config, "ppo_kl_coef", 1.0)
    neg_kl = log_prob - old_log_prob
    ratio = torch.exp(neg_kl)
    base = -advantages * ratio
    with_kl = base + coef * neg_kl.abs()
    valid = response_mask > 0
    losses = base
    n_valid = int(valid.sum().item())
    if n_valid > 0:
        adv_v, lp_v = advantages[valid].detach(), log_prob[valid].detach()
        cov = (adv_v - adv_v.mean()) * (lp_v - lp_v.mean())
        k = max(1, int(n_valid * k_ratio))
        top = torch.topk(cov, k, largest=True).indices
        flat_idx = valid.reshape(-1).nonzero(as_tuple=True)[0][top]
        sel = torch.zeros(valid.numel(), dtype=torch.bool, device=valid.device)
        sel[flat_idx] = True
        losses = torch.where(sel.reshape(valid.shape), with_kl, base)
    if rollout_is_weights is not None:
        losses = losses * rollout_is_weights
    loss = agg_loss(losses, response_mask, loss_agg_mode)
    return loss, {"actor/ppo_kl": masked_mean(neg_kl.abs(), response_mask).item()}


@register_policy_loss("geo_mean")
def compute_policy_loss_geo_mean(old_log_prob, log_prob, advantages, response_mask, loss_agg_mode="token-mean",
                                 config=None, rollout_is_weights=None):
    """GMPO (arXiv:2507.20673): geometric-mean sequence ratio with token-level log-ratio clipping."""
    _, lo, hi = _clip_bounds(config)
    neg_kl = log_prob - old_log_prob
    sgn = torch.sign(advantages)
    clamped = torch.clamp(neg_kl, -lo, hi)
    chosen = sgn * torch.minimum(sgn * neg_kl, sgn * clamped)
    n_tok = response_mask.sum(-1)
    ratio = torch.exp((chosen * response_mask).sum(-1) / (n_tok + 1e-8))
    adv_seq = (advantages * response_mask).sum(-1) / (n_tok + 1e-8)
    losses = -adv_seq * ratio
    if rollout_is_weights is not None:
        w = torch.exp((torch.log(rollout_is_weights + 1e-10) * response_mask).sum(-1) / (n_tok + 1e-8))
        losses = losses * w
    clipped = neg_kl != clamped
    return losses.mean(), {
        "actor/pg_clipfrac": masked_mean((clipped & (advantages > 0)).float(), response_mask).item(),
        "actor/ppo_kl": masked_mean(-neg_kl, response_mask).item(),
        "actor/pg_clipfrac_lower": masked_mean((clipped & (advantages < 0)).float(), response_mask).item(),
    }


@register_policy_loss("rollout_correction")
def compute_policy_loss_rollout_correction(old_log_prob, log_prob, advantages, response_mask,
                                           loss_agg_mode="token-mean", config=None, rollout_is_weights=None):
    """Pure importance-weighted policy gradient (no PPO clip): -A * w * log pi, w from the rollout
    correction helper (policy gradient evaluated on-policy w.r.t. old_log_prob == log_prob.detach())."""
    w = rollout_is_weights if rollout_is_weights is not None else torch.ones_like(advantages)
    ratio = torch.exp(torch.clamp(log_prob - old_log_prob, -20.0, 20.0))
    losses = -advantages * ratio * w
    loss = agg_loss(losses, response_mask, loss_agg_mode)
    return loss, {"actor/ppo_kl": masked_mean(old_log_prob - log_prob, response_mask).item(), "actor/pg_clipfrac": 0.0}


def compute_entropy_loss(entropy: torch.Tensor, response_mask: torch.Tensor, loss_agg_mode: str = "token-mean"):
    return agg_loss(entropy, response_mask, loss_agg_mode)


def compute_value_loss(vpreds, returns, values, response_mask, cliprange_value: float, loss_agg_mode="token-mean"):
    clipped = torch.clamp(vpreds, values - cliprange_value, values + cliprange_value)
    l1 = (vpreds - returns) ** 2
    l2 = (clipped - returns) ** 2
    loss = 0.5 * agg_loss(torch.maximum(l1, l2), response_mask, loss_agg_mode)
    clipfrac = masked_mean((l2 > l1).float(), response_mask)
    return loss, clipfrac


def _kl_forward(logprob, ref_logprob, kind: str) -> torch.Tensor:
    if kind in ("kl", "k1"):
        return logprob - ref_logprob
    if kind == "abs":
        return (logprob - ref_logprob).abs()
    if kind in ("mse", "k2"):
        return 0.5 * (logprob - ref_logprob) ** 2
    if kind in ("low_var_kl", "k3"):
        d = torch.clamp(ref_logprob - logprob, -20.0, 20.0)
        return torch.clamp(torch.exp(d) - d - 1.0, -10.0, 10.0)
    if kind == "full":
        raise NotImplementedError("full-distribution KL needs logits")
    raise NotImplementedError(f"kl penalty {kind!r}")


def kl_penalty(logprob: torch.Tensor, ref_logprob: torch.Tensor, kl_penalty: str) -> torch.Tensor:
    """Token-level KL estimators; a trailing '+' uses the estimator's value with the k2 gradient
    (straight-through), which is an unbiased gradient estimator."""
    if kl_penalty.endswith("+"):
        val = _kl_forward(logprob, ref_logprob, kl_penalty[:-1])
        grad = _kl_forward(logprob, ref_logprob, "k2")
        return grad - grad.detach() + val.detach()
    return _kl_forward(logprob, ref_logprob, kl_penalty)
