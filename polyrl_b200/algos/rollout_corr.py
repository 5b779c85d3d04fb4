"""Rollout-correction: importance-sampling weights and rejection masks that correct the mismatch between the
rollout engine's sampling distribution (``rollout_log_probs``, e.g. fp8 weights / different kernels) and the
trainer's policy (``old_log_probs``), plus off-policy diagnostics.

Capability parity with verl's helper (3rdparty/verl/verl/trainer/ppo/rollout_corr_helper.py:82-879): truncated
IS at token / sequence / geometric-mean level, rejection sampling (mask) with upper/lower thresholds, per-token
veto, and the ``rollout_corr/*`` metrics.
"""
from __future__ import annotations

import torch

from .core_algos import masked_mean

SAFETY = 20.0


def _level_log_ratio(log_ratio: torch.Tensor, mask: torch.Tensor, level: str) -> torch.Tensor:
    if level == "token":
        return log_ratio
    if level == "sequence":
        return (log_ratio * mask).sum(-1, keepdim=True).expand_as(log_ratio)
    if level == "geometric":
        return ((log_ratio * mask).sum(-1, keepdim=True) / mask.sum(-1, keepdim=True).clamp_min(1)).expand_as(log_ratio)
    raise ValueError(f"unknown rollout-correction level {level!r} (token | sequence | geometric)")


@torch.no_grad()
def compute_rollout_correction(old_log_prob: torch.Tensor, rollout_log_prob: torch.Tensor, response_mask: torch.Tensor,
                               rollout_is: str | None = None, rollout_is_threshold: float = 2.0, rollout_rs: str | None = None,
                               rollout_rs_threshold: float | None = None, rollout_rs_threshold_lower: float | None = None,
                               rollout_token_veto_threshold: float | None = None):
    """-> (is_weights or None, modified_response_mask, metrics).

    * ``rollout_is``: truncated importance weights ``min(exp(lvl(log pi_train - log pi_rollout)), threshold)``.
    * ``rollout_rs``: rejection — tokens / sequences whose ratio leaves ``[lower, upper]`` are masked out.
    * ``rollout_token_veto_threshold``: a single catastrophic token (ratio < threshold) rejects its sequence.
    """
    mask = response_mask.float()
    log_ratio = torch.clamp(old_log_prob - rollout_log_prob, -SAFETY, SAFETY)
    metrics: dict[str, float] = {}
    weights = None
    if rollout_is:
        w = torch.exp(_level_log_ratio(log_ratio, mask, rollout_is))
        metrics["rollout_corr/is_mean_raw"] = masked_mean(w, mask).item()
        metrics["rollout_corr/is_truncated_frac"] = masked_mean((w > rollout_is_threshold).float(), mask).item()
        weights = torch.clamp(w, max=rollout_is_threshold) * mask
        metrics["rollout_corr/is_mean"] = masked_mean(weights, mask).item()
        ess_num = (weights.sum()) ** 2
        metrics["rollout_corr/is_eff_sample_size"] = (ess_num / ((weights ** 2).sum() + 1e-8) / mask.sum().clamp_min(1)).item()
    new_mask = response_mask
    if rollout_rs:
        upper = rollout_rs_threshold if rollout_rs_threshold is not None else rollout_is_threshold
        lower = rollout_rs_threshold_lower if rollout_rs_threshold_lower is not None else 1.0 / upper
        r = torch.exp(_level_log_ratio(log_ratio, mask, rollout_rs))
        keep = (r <= upper) & (r >= lower)
        metrics["rollout_corr/rs_masked_frac"] = 1.0 - masked_mean(keep.float(), mask).item()
        new_mask = response_mask * keep.to(response_mask.dtype)
    if rollout_token_veto_threshold is not None:
        bad = ((torch.exp(log_ratio) < rollout_token_veto_threshold) & (mask > 0)).any(-1, keepdim=True)
        metrics["rollout_corr/veto_seq_frac"] = bad.float().mean().item()
        new_mask = new_mask * (~bad).to(new_mask.dtype)
    # off-policy diagnostics
    metrics["rollout_corr/kl"] = masked_mean(-log_ratio, mask).item()                        # E_rollout[log pi_r - log pi_t]
    metrics["rollout_corr/k3_kl"] = masked_mean(torch.exp(log_ratio) - 1 - log_ratio, mask).item()
    metrics["rollout_corr/log_ppl_diff"] = (masked_mean(-old_log_prob, mask) - masked_mean(-rollout_log_prob, mask)).item()
    metrics["rollout_corr/chi2_token"] = (masked_mean(torch.exp(2 * log_ratio), mask) - 1).item()
    return weights, new_mask, metrics
