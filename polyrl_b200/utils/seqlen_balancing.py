"""Sequence-length balancing: Karmarkar–Karp (largest differencing) and greedy k-way partitioners,
token-budget dynamic micro-batching, and DP balance statistics.

Capability parity: 3rdparty/verl/verl/utils/seqlen_balancing.py:37-416 (partitioners :37-202,
``log_seqlen_unbalance`` :205-250, ``rearrange_micro_batches`` :261-334).
"""
from __future__ import annotations

import heapq
from typing import Sequence

import torch
import torch.distributed as dist


def greedy_partition(seqlens: Sequence[int], k: int, equal_size: bool = False) -> list[list[int]]:
    """Assign items (largest first) to the currently lightest bin."""
    n = len(seqlens)
    if equal_size and n % k != 0:
        raise ValueError(f"{n} items cannot be split into {k} equal-size partitions")
    cap = n // k if equal_size else n
    order = sorted(range(n), key=lambda i: -seqlens[i])
    bins: list[list[int]] = [[] for _ in range(k)]
    sums = [0] * k
    for i in order:
        cands = [b for b in range(k) if len(bins[b]) < cap]
        b = min(cands, key=lambda j: (sums[j], j))
        bins[b].append(i)
        sums[b] += seqlens[i]
    return bins


def karmarkar_karp(seqlens: Sequence[int], k: int, equal_size: bool = False) -> list[list[int]]:
    """k-way largest-differencing method.  A *state* is k bins sorted by sum (descending); two states
    merge by pairing the heaviest bin of one with the lightest of the other.  With ``equal_size`` items
    are first grouped k at a time so every bin ends with the same count."""
    n = len(seqlens)
    if equal_size and n % k != 0:
        raise ValueError(f"{n} items cannot be split into {k} equal-size partitions")
    order = sorted(range(n), key=lambda i: seqlens[i])
    states = []   # heap of (-spread, tiebreak, bins) ; bins = list of (sum, [indices]) sorted desc by sum
    counter = 0

    def push(bins):
        nonlocal counter
        bins = sorted(bins, key=lambda b: -b[0])
        heapq.heappush(states, (-(bins[0][0] - bins[-1][0]), counter, bins))
        counter += 1

    if equal_size:
        for off in range(0, n, k):
            grp = order[off:off + k]
            push([(seqlens[i], [i]) for i in grp])
    else:
        for i in order:
            push([(seqlens[i], [i])] + [(0, []) for _ in range(k - 1)])
    if not states:
        return [[] for _ in range(k)]
    while len(states) > 1:
        _, _, a = heapq.heappop(states)
        _, _, b = heapq.heappop(states)
        merged = [(a[j][0] + b[k - 1 - j][0], a[j][1] + b[k - 1 - j][1]) for j in range(k)]
        push(merged)
    bins = states[0][2]
    return [sorted(b[1]) for b in bins]


def get_seqlen_balanced_partitions(seqlens: Sequence[int], k_partitions: int, equal_size: bool) -> list[list[int]]:
    if len(seqlens) < k_partitions:
        raise ValueError(f"cannot split {len(seqlens)} items into {k_partitions} non-empty partitions")
    parts = karmarkar_karp(list(seqlens), k_partitions, equal_size)
    seen = sorted(i for p in parts for i in p)
    assert seen == list(range(len(seqlens)))
    if any(len(p) == 0 for p in parts):      # degenerate KK outcome: fall back to greedy with non-empty bins
        parts = greedy_partition(list(seqlens), k_partitions, equal_size)
    return [sorted(p) for p in parts]


def log_seqlen_unbalance(seqlen_list: Sequence[int], partitions: list[list[int]], prefix: str) -> dict:
    k = len(partitions)
    n = len(seqlen_list)
    per = n // k
    naive = [sum(seqlen_list[i * per:(i + 1) * per]) for i in range(k)]
    bal = [sum(seqlen_list[i] for i in p) for p in partitions]
    return {
        f"{prefix}/min": min(naive), f"{prefix}/max": max(naive), f"{prefix}/minmax_diff": max(naive) - min(naive),
        f"{prefix}/balanced_min": min(bal), f"{prefix}/balanced_max": max(bal), f"{prefix}/mean": sum(naive) / k,
    }


def ceildiv(a: int, b: int) -> int:
    return -(-a // b)


def rearrange_micro_batches(seq_lens: Sequence[int], max_token_len: int, dp_group=None, same_micro_num_in_dp: bool = True,
                            min_num_micro_batch: int | None = None, num_batches_divided_by: int | None = None
                            ) -> list[list[int]]:
    """Split sample indices into the fewest micro-batches whose token totals fit ``max_token_len``,
    balanced with Karmarkar–Karp; the number of micro-batches is equalised across the DP group
    (all-reduce MAX) so collectives inside the model stay matched."""
    if not len(seq_lens):
        return []
    if max(seq_lens) > max_token_len:
        raise ValueError(f"max_token_len={max_token_len} is smaller than the longest sequence {max(seq_lens)}")
    total = int(sum(seq_lens))
    num = min(len(seq_lens), ceildiv(total, max_token_len))
    if min_num_micro_batch is not None:
        num = max(num, min(min_num_micro_batch, len(seq_lens)))
    if num_batches_divided_by:
        num = ceildiv(num, num_batches_divided_by) * num_batches_divided_by
    num = max(1, min(num, len(seq_lens)))
    while True:
        parts = get_seqlen_balanced_partitions(seq_lens, num, equal_size=False)
        if max(sum(seq_lens[i] for i in p) for p in parts) <= max_token_len or num >= len(seq_lens):
            break
        num += 1
    if same_micro_num_in_dp and dist.is_initialized() and dist.get_world_size(dp_group) > 1:
        backend = dist.get_backend(dp_group)
        dev = "cuda" if "nccl" in str(backend) else "cpu"
        t = torch.tensor([num], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=dp_group)
        synced = min(int(t.item()), len(seq_lens))
        if synced != num:
            num = synced
            parts = get_seqlen_balanced_partitions(seq_lens, num, equal_size=False)
    # heavier (quadratic attention cost) micro-batches first -> better tail behaviour
    parts.sort(key=lambda p: -sum(seq_lens[i] ** 2 for i in p))
    return parts
