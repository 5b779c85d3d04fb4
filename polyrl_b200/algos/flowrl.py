"""FlowRL (arXiv:2509.15207): match the policy to the reward DISTRIBUTION with a trajectory-balance objective instead of
maximising expected reward.  A small head predicts log Z(x) from the mean prompt hidden state and every sampled response is pushed
towards  log Z(x) + mean_t log pi(y_t) = beta * r(x, y) + mean_t log pi_ref(y_t),  weighted by a clipped sequence importance ratio.

Reference behaviour: 3rdparty/verl/recipe/flowrl/flowrl_actor.py (``ProjZModule`` :36-58, prompt-hidden pooling :256-276,
``compute_flowrl_objective`` :417-490).  Own implementation on this framework's engines: the head is a plain replicated module
(a few hidden x hidden matrices) outside the ZeRO engine, its gradients are all-reduced over the data-parallel group at the
optimizer step.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class LogZHead(torch.nn.Module):
    """(Linear -> GELU -> LayerNorm -> Dropout) x (num_layers - 1) -> Linear(hidden, 1) on the mean prompt hidden state."""

    def __init__(self, hidden_size: int, num_layers: int = 3, dropout: float = 0.1):
        super().__init__()
        layers: list[torch.nn.Module] = []
        for _ in range(max(0, num_layers - 1)):
            layers += [torch.nn.Linear(hidden_size, hidden_size), torch.nn.GELU(), torch.nn.LayerNorm(hidden_size),
                       torch.nn.Dropout(dropout)]
        layers.append(torch.nn.Linear(hidden_size, 1))
        self.net = torch.nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net(x.float()).squeeze(-1)


def _masked_mean(x: torch.Tensor, m: torch.Tensor, dim: int = -1) -> torch.Tensor:
    return (x * m).sum(dim) / m.sum(dim).clamp(min=1e-8)


def flowrl_objective(log_prob, ref_log_prob, old_log_prob, log_z, reward, response_mask, clip_low: float = 0.2,
                     clip_high: float = 0.2, beta: float = 15.0):
    """Trajectory-balance loss on a micro-batch.  ``reward`` is the token-level advantage tensor (group-normalised outcome reward
    broadcast over the response).  Tokens whose ratio already left the trust region in the direction their reward pushes are
    dropped from the per-sequence averages (the CISPO-style mask of the reference).  -> (loss, metrics)"""
    mask = response_mask.float()
    ratio = torch.exp(log_prob - old_log_prob)
    drop = ((reward > 0) & (ratio > 1.0 + clip_high)) | ((reward < 0) & (ratio < 1.0 - clip_low))
    keep = mask * (~drop).float()
    avg_lp = _masked_mean(log_prob, keep)
    avg_ref = _masked_mean(ref_log_prob, keep)
    seq_r = _masked_mean(reward, keep)
    delta = log_z + avg_lp - beta * seq_r - avg_ref
    imp_raw = torch.exp(((log_prob - old_log_prob) * keep).sum(-1)).detach()
    imp = imp_raw.clamp(1.0 - clip_low, 1.0 + clip_high)
    loss = (imp * delta.pow(2)).mean()
    tot = mask.sum().clamp(min=1.0)
    metrics = {
        "actor/flowrl_loss": loss.detach().item(), "actor/log_z": log_z.detach().mean().item(),
        "actor/flowrl_residual": delta.detach().abs().mean().item(),
        "actor/importance_weight_raw": imp_raw.mean().item(), "actor/importance_weight": imp.mean().item(),
        "actor/ppo_kl": ((old_log_prob - log_prob).detach() * mask).sum().item() / tot.item(),
        "actor/ref_kl": ((ref_log_prob - log_prob).detach() * mask).sum().item() / tot.item(),
        "actor/cispo_mask_ratio": ((mask * drop.float()).sum() / tot).item(),
    }
    return loss, metrics
