"""Ulysses sequence parallelism for the packed (remove-padding) training path.

Outside attention every rank of the SP group holds a contiguous 1/sp slice of the packed token stream (activation
memory ∝ T/sp); around attention an all-to-all swaps the sharded dimension: [T/sp, heads, d] -> [T, heads/sp, d],
varlen attention runs on full sequences with a subset of the heads, and a second all-to-all swaps back.
KV heads are replicated when ``sp > n_kv``.

Capability parity: 3rdparty/verl/verl/utils/ulysses.py:62-321 (gather_seq_scatter_heads / gather_heads_scatter_seq,
``SeqAllToAll`` autograd, ``ulysses_pad_and_slice_inputs``, ``Gather`` with grad scaling) and
verl/models/transformers/monkey_patch.py:49-117 (KV-head replication, constraints :273-280).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class PeerAllToAll:
    """All-to-all over NVLink peer memory: every rank WRITES chunk j straight into rank j's receive slot with the peer-copy
    kernel (csrc/p2p.cu), bracketed by device-side barriers — no NCCL, no staging, stream-ordered and graph-safe.
    Replaces ``dist.all_to_all`` on list-of-chunks (reference: verl/utils/ulysses.py:133-152) on one NVSwitch domain.
    Two receive buffers alternate so a rank may start the next exchange while peers still read the previous one."""

    def __init__(self, heap, max_bytes: int):
        self.heap, self.world, self.rank = heap, heap.world, heap.rank
        self.slot_bytes = ((max_bytes + 255) // 256) * 256
        self.bufs = [heap.alloc(self.world * self.slot_bytes, torch.uint8) for _ in range(2)]
        self.turn = 0
        self.C = heap._C

    def __call__(self, parts: list[torch.Tensor]) -> list[torch.Tensor]:
        nbytes = parts[0].numel() * parts[0].element_size()
        if nbytes > self.slot_bytes:
            raise ValueError(f"all-to-all chunk of {nbytes} B exceeds the symmetric slot ({self.slot_bytes} B)")
        buf = self.bufs[self.turn]
        self.turn ^= 1
        self.heap.barrier()                                    # peers finished reading this buffer two exchanges ago
        for j, p in enumerate(parts):
            self.C.p2p_copy(p.data_ptr(), buf.ptrs[j] + self.rank * self.slot_bytes, nbytes, 16)
        self.heap.barrier()                                    # every rank's chunks have landed
        flat = buf.local
        return [flat[i * self.slot_bytes:i * self.slot_bytes + nbytes].view(parts[0].dtype).view(parts[0].shape)
                for i in range(self.world)]


_PEER_A2A: dict = {}


def enable_peer_all_to_all(group, heap, max_bytes: int) -> None:
    """Route this SP group's all-to-alls through NVLink peer memory (``heap`` must span exactly ``group``)."""
    _PEER_A2A[id(group)] = PeerAllToAll(heap, max_bytes)


def _all_to_all(x: torch.Tensor, scatter_dim: int, gather_dim: int, group) -> torch.Tensor:
    """Split ``x`` into sp chunks along ``scatter_dim``, exchange, concatenate along ``gather_dim``."""
    sp = dist.get_world_size(group)
    if sp == 1:
        return x
    parts = [p.contiguous() for p in x.chunk(sp, dim=scatter_dim)]
    peer = _PEER_A2A.get(id(group))
    if peer is not None and x.is_cuda:
        return torch.cat(peer(parts), dim=gather_dim)
    if x.is_cuda:
        out = [torch.empty_like(parts[0]) for _ in range(sp)]
        dist.all_to_all(out, parts, group=group)
    else:                                   # gloo has no all_to_all: emulate with all_gather (CPU plumbing tests)
        rank = dist.get_rank(group)
        stacked = torch.stack(parts)        # [sp, ...] what I send to each peer
        gathered = [torch.empty_like(stacked) for _ in range(sp)]
        dist.all_gather(gathered, stacked, group=group)
        out = [g[rank] for g in gathered]   # what each peer sends to me
    return torch.cat(out, dim=gather_dim)


class SeqAllToAll(torch.autograd.Function):
    """Differentiable all-to-all: the backward is the inverse exchange."""

    @staticmethod
    def forward(ctx, x, scatter_dim: int, gather_dim: int, group):
        ctx.dims, ctx.group = (scatter_dim, gather_dim), group
        return _all_to_all(x, scatter_dim, gather_dim, group)

    @staticmethod
    def backward(ctx, g):
        s, ga = ctx.dims
        return _all_to_all(g.contiguous(), ga, s, ctx.group), None, None, None


def gather_seq_scatter_heads(x: torch.Tensor, group, seq_dim: int = 0, head_dim: int = 1) -> torch.Tensor:
    """[T/sp, H, d] -> [T, H/sp, d]"""
    return SeqAllToAll.apply(x, head_dim, seq_dim, group)


def gather_heads_scatter_seq(x: torch.Tensor, group, seq_dim: int = 0, head_dim: int = 1) -> torch.Tensor:
    """[T, H/sp, d] -> [T/sp, H, d]"""
    return SeqAllToAll.apply(x, seq_dim, head_dim, group)


class GatherSeq(torch.autograd.Function):
    """All-gather along the sequence dim; backward keeps the local slice (optionally scaled by sp, matching the
    reference's ``Gather(grad_scaler=True)``)."""

    @staticmethod
    def forward(ctx, x, group, grad_scale: bool):
        sp, rank = dist.get_world_size(group), dist.get_rank(group)
        ctx.meta = (sp, rank, x.shape[0], grad_scale)
        if sp == 1:
            return x
        parts = [torch.empty_like(x) for _ in range(sp)]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, dim=0)

    @staticmethod
    def backward(ctx, g):
        sp, rank, n, scale = ctx.meta
        if sp == 1:
            return g, None, None
        out = g[rank * n:(rank + 1) * n]
        return (out * sp if scale else out), None, None


def ulysses_pad_and_slice(input_ids: torch.Tensor, position_ids: torch.Tensor, group, pad_id: int = 0):
    """Pad the packed stream to a multiple of sp and return this rank's slice plus the pad size."""
    sp, rank = dist.get_world_size(group), dist.get_rank(group)
    t = input_ids.shape[0]
    pad = (-t) % sp
    if pad:
        input_ids = torch.cat([input_ids, input_ids.new_full((pad,), pad_id)])
        position_ids = torch.cat([position_ids, position_ids.new_zeros(pad)])
    n = input_ids.shape[0] // sp
    return input_ids[rank * n:(rank + 1) * n], position_ids[rank * n:(rank + 1) * n], pad


def validate_sp(num_heads: int, num_kv_heads: int, sp: int) -> None:
    if num_heads % sp != 0:
        raise ValueError(f"num_heads ({num_heads}) must be divisible by ulysses_sequence_parallel_size ({sp})")
    if not (num_kv_heads % sp == 0 or sp % num_kv_heads == 0):
        raise ValueError(f"num_kv_heads ({num_kv_heads}) must divide or be divisible by sp ({sp})")


def ulysses_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int, group,
                      total_tokens: int, causal: bool = True, scale: float | None = None) -> torch.Tensor:
    """q:[T/sp, nq, d], k/v:[T/sp, nkv, d] (local token slice, all heads) -> [T/sp, nq, d].

    ``cu_seqlens`` describe the FULL packed stream (identical on every SP rank); the tail padding tokens (beyond
    ``total_tokens``) form a dummy sequence."""
    from ..ops.attention import attention_varlen

    sp = dist.get_world_size(group)
    nkv = k.shape[1]
    if sp > nkv:                     # replicate kv heads so they can be scattered sp ways
        rep = sp // nkv
        k, v = k.repeat_interleave(rep, dim=1), v.repeat_interleave(rep, dim=1)
    qf = gather_seq_scatter_heads(q, group)
    kf = gather_seq_scatter_heads(k, group)
    vf = gather_seq_scatter_heads(v, group)
    t_full = qf.shape[0]
    cu = cu_seqlens
    if t_full > total_tokens:        # padding tail = one extra sequence
        cu = torch.cat([cu_seqlens, cu_seqlens.new_tensor([t_full])])
        max_seqlen = max(max_seqlen, t_full - total_tokens)
    o = attention_varlen(qf, kf, vf, cu, max_seqlen, causal=causal, scale=scale)
    return gather_heads_scatter_seq(o, group)
