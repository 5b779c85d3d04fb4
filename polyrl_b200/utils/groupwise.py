"""Group-wise statistics over flat score vectors (one score per rollout, groups = the n samples of a prompt), as used by the
group-relative advantage estimators and the group filters.  Capability parity: 3rdparty/verl/verl/utils/groupwise.py
(``as_torch_index``, ``group_mean_std``)."""
from __future__ import annotations

import numpy as np
import torch


def as_torch_index(ids, device=None) -> torch.Tensor:
    """Arbitrary group labels (uid strings, ints, numpy / torch arrays) -> dense int64 indices 0..G-1 in order of first appearance."""
    if torch.is_tensor(ids) and ids.dtype in (torch.int32, torch.int64, torch.int16, torch.uint8):
        _, inv = torch.unique(ids.reshape(-1).cpu(), sorted=False, return_inverse=True)
        labels = inv.tolist()
    else:
        labels = np.asarray(ids, dtype=object).reshape(-1).tolist()
    seen: dict = {}
    idx = [seen.setdefault(x if not torch.is_tensor(x) else x.item(), len(seen)) for x in labels]
    return torch.tensor(idx, dtype=torch.long, device=device)


def group_mean_std(scores: torch.Tensor, gidx: torch.Tensor, eps: float = 1e-6, device=None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(mean, std, count) per group, fp32, in three ``index_add_`` passes (no Python loop over groups).  std is the unbiased sample
    standard deviation with its variance floored at ``eps``; groups of one sample get mean 0 / std 1 so their advantage is the raw score."""
    dev = torch.device(device) if device is not None else scores.device
    s = scores.reshape(-1).to(dev, torch.float32)
    g = gidx.reshape(-1).to(dev, torch.long)
    if s.numel() != g.numel():
        raise ValueError(f"scores and gidx length mismatch: {s.numel()} vs {g.numel()}")
    if s.numel() == 0:
        e = torch.empty(0, device=dev, dtype=torch.float32)
        return e, e.clone(), e.clone()
    n = int(g.max().item()) + 1
    cnt = torch.zeros(n, device=dev).index_add_(0, g, torch.ones_like(s))
    s1 = torch.zeros(n, device=dev).index_add_(0, g, s)
    s2 = torch.zeros(n, device=dev).index_add_(0, g, s * s)
    c = cnt.clamp_min(1.0)
    mean = s1 / c
    std = ((s2 - s1 * s1 / c) / (cnt - 1.0).clamp_min(1.0)).clamp_min(eps).sqrt()
    single = cnt <= 1.0
    return mean.masked_fill(single, 0.0), std.masked_fill(single, 1.0), cnt
