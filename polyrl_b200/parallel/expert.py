"""Expert parallelism for the mixture-of-experts family: the experts of every block are split over the ranks of ``ep_group``; tokens
travel to the rank that owns their expert and back (dispatch / combine all-to-all), everything else (attention, router, norms,
embeddings) is replicated and data-parallel over the same ranks.

    tokens [T_r, H] --router top-k--> assignments sorted by expert --counts all-to-all--> dispatch all-to-all (variable splits)
      --> local experts (fused SwiGLU projections per expert) --> combine all-to-all (reverse splits) --> weighted scatter-add

Transport: ``torch.distributed.all_to_all_single`` with split sizes on NCCL; the padded all-gather emulation on gloo (CPU tests);
or — ``enable_peer_dispatch`` — direct NVLink peer writes of every destination's rows into its receive window (``PeerVarAllToAll``
over the symmetric heap, the variable-size sibling of ``ulysses.PeerAllToAll``).  Both directions are differentiable (the backward
of an exchange is the reverse exchange).

Capability parity: the reference reaches MoE models through Megatron's expert parallelism (``expert_model_parallel_size``,
3rdparty/verl/verl/workers/megatron_workers.py:210-211 and mcore's token dispatcher).  ``EPTrainEngine`` is the matching training
engine: AdamW over replicated dense parameters (gradients averaged over the group) and rank-local expert slices.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .. import ops
from ..models import transformer as tf
from ..models.config import ModelConfig
from ..ops import reference as ref

_PEER: dict = {}


class PeerVarAllToAll:
    """Variable-size all-to-all over NVLink peer memory.  Rank r writes the rows destined to rank j into j's receive window at
    slot r (fixed ``slot_bytes`` stride, so no offsets have to be negotiated); two windows alternate like ``PeerAllToAll``."""

    def __init__(self, heap, slot_bytes: int):
        self.heap, self.world, self.rank = heap, heap.world, heap.rank
        self.slot_bytes = ((slot_bytes + 255) // 256) * 256
        self.bufs = [heap.alloc(self.world * self.slot_bytes, torch.uint8) for _ in range(2)]
        self.turn = 0
        self.C = heap._C

    def __call__(self, x: torch.Tensor, send_counts: list[int], recv_counts: list[int]) -> torch.Tensor:
        row = x.shape[1] * x.element_size()
        if max(max(send_counts), max(recv_counts)) * row > self.slot_bytes:
            raise ValueError("expert dispatch exceeds the symmetric slot; raise max_tokens_per_peer")
        buf = self.bufs[self.turn]
        self.turn ^= 1
        self.heap.barrier()
        lo = 0
        for j, n in enumerate(send_counts):
            if n:
                self.C.p2p_copy(x.data_ptr() + lo * row, buf.ptrs[j] + self.rank * self.slot_bytes, n * row, 16)
            lo += n
        self.heap.barrier()
        flat = buf.local
        parts = [flat[i * self.slot_bytes:i * self.slot_bytes + n * row].view(x.dtype).view(n, x.shape[1]) for i, n in enumerate(recv_counts)]
        return torch.cat(parts, dim=0)


def enable_peer_dispatch(group, heap, max_tokens_per_peer: int, hidden_size: int, elem_bytes: int = 2) -> None:
    """Route this EP group's dispatch / combine through NVLink peer memory (``heap`` must span exactly ``group``)."""
    _PEER[id(group)] = PeerVarAllToAll(heap, max_tokens_per_peer * hidden_size * elem_bytes)


def _exchange(x: torch.Tensor, send_counts: list[int], recv_counts: list[int], group) -> torch.Tensor:
    """Rows of ``x`` are grouped by destination rank (``send_counts``); returns the rows received, grouped by source rank."""
    ep = dist.get_world_size(group)
    if ep == 1:
        return x
    peer = _PEER.get(id(group))
    if peer is not None and x.is_cuda:
        return peer(x.contiguous(), send_counts, recv_counts)
    if x.is_cuda:
        out = x.new_empty(sum(recv_counts), x.shape[1])
        dist.all_to_all_single(out, x.contiguous(), recv_counts, send_counts, group=group)
        return out
    # gloo: pad every destination block to the largest one and all-gather
    rank = dist.get_rank(group)
    m = torch.tensor([max(send_counts)], dtype=torch.long)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    m = int(m.item())
    padded = x.new_zeros(ep, m, x.shape[1])
    lo = 0
    for j, n in enumerate(send_counts):
        padded[j, :n] = x[lo:lo + n]
        lo += n
    gathered = [torch.empty_like(padded) for _ in range(ep)]
    dist.all_gather(gathered, padded, group=group)
    return torch.cat([gathered[i][rank, :n] for i, n in enumerate(recv_counts)], dim=0)


class _VarAllToAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, send_counts, recv_counts, group):
        ctx.send, ctx.recv, ctx.group = send_counts, recv_counts, group
        return _exchange(x, send_counts, recv_counts, group)

    @staticmethod
    def backward(ctx, g):
        return _exchange(g.contiguous(), ctx.recv, ctx.send, ctx.group), None, None, None


def local_expert_range(cfg: ModelConfig, ep: int, rank: int) -> tuple[int, int]:
    if cfg.num_experts % ep:
        raise ValueError(f"{cfg.num_experts} experts are not divisible by the expert-parallel size {ep}")
    per = cfg.num_experts // ep
    return rank * per, (rank + 1) * per


def ep_moe_mlp(cfg: ModelConfig, w: dict, h: torch.Tensor, group) -> torch.Tensor:
    """``models.transformer.moe_mlp`` with the experts split over ``group``: ``w['ewgu']`` / ``w['ewd']`` hold this rank's experts
    only, ``w['router']`` is the full (replicated) router."""
    ep, rank = dist.get_world_size(group), dist.get_rank(group)
    per = cfg.num_experts // ep
    t, k = h.shape[0], cfg.num_experts_per_tok
    top_w, top_e, _ = tf.moe_route(cfg, w, h)
    flat_e = top_e.reshape(-1)
    order = torch.argsort(flat_e, stable=True)                    # grouped by expert == grouped by destination rank
    tok = torch.div(order, k, rounding_mode="floor")
    counts = torch.bincount(flat_e, minlength=cfg.num_experts)   # [E] assignments per global expert
    # every rank learns how many rows each peer sends to each of its local experts
    all_counts = [torch.empty_like(counts) for _ in range(ep)]
    dist.all_gather(all_counts, counts, group=group)
    mine = torch.stack([c[rank * per:(rank + 1) * per] for c in all_counts])          # [ep (source), per (local expert)]
    send_counts = counts.view(ep, per).sum(-1).tolist()
    recv_counts = mine.sum(-1).tolist()
    xs = _VarAllToAll.apply(h.index_select(0, tok), send_counts, recv_counts, group)   # rows grouped by source, then by expert
    # regroup by local expert: row (source s, expert e, i) -> expert-major
    src_exp = mine.tolist()
    seg, lo = [], 0
    for s in range(ep):
        for e in range(per):
            seg.append((e, s, lo, src_exp[s][e]))
            lo += src_exp[s][e]
    perm = torch.cat([torch.arange(a, a + n, device=h.device) for e, s, a, n in sorted(seg) if n] or [torch.zeros(0, dtype=torch.long, device=h.device)])
    xe = xs.index_select(0, perm)
    outs, lo = [], 0
    for e in range(per):
        n = int(mine[:, e].sum())
        if n:
            a = ops.linear_swiglu(xe[lo:lo + n], w["ewgu"][e])
            outs.append(ops.linear(a, w["ewd"][e]))
            lo += n
    ye = torch.cat(outs, dim=0) if outs else xe.new_zeros(0, h.shape[1])
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(perm.numel(), device=h.device)
    ys = _VarAllToAll.apply(ye.index_select(0, inv), recv_counts, send_counts, group)   # back to the token owners, assignment order
    y = ys * top_w.reshape(-1)[order].to(h.dtype).unsqueeze(-1)
    y = torch.zeros(t, h.shape[1], dtype=y.dtype, device=h.device).index_add(0, tok, y)
    s = tf.shared_expert(w, h)                                    # replicated (data-parallel) like the rest of the dense part
    return y if s is None else y + s


def ep_layer_forward(cfg: ModelConfig, w: dict, x: torch.Tensor, inp: tf.PackedInput, group) -> torch.Tensor:
    """A decoder block whose MoE MLP is expert-parallel: attention through the ordinary block code with the MLP swapped.  Dense
    blocks of a mixed stack run the ordinary block."""
    if "router" not in w:
        return tf.layer_forward(cfg, w, x, inp)
    dense = {k: v for k, v in w.items() if k not in ("router", "ewgu", "ewd")}
    dense["_skip_mlp"] = True
    xa = tf.layer_forward(cfg, dense, x, inp)                     # attention half (+ residual); see layer_forward
    hh = ops.rmsnorm(xa, w["ln2"], cfg.rms_norm_eps)
    return xa + ep_moe_mlp(cfg, w, hh, group)


class EPTrainEngine:
    """Expert-parallel (= data-parallel for the dense part) training engine for MoE models."""

    EXPERT = ("ewgu", "ewd")

    def __init__(self, cfg: ModelConfig, device, ep_group, lr: float = 1e-6, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.01, max_grad_norm: float = 1.0, seed: int = 0, dtype=None):
        if not cfg.num_experts:
            raise ValueError("expert parallelism needs a mixture-of-experts model")
        if cfg.sandwich_norm:
            raise NotImplementedError("expert parallelism for blocks with output norms (sandwich_norm)")
        self.cfg, self.device, self.group = cfg, torch.device(device), ep_group
        self.ep, self.rank = dist.get_world_size(ep_group), dist.get_rank(ep_group)
        self.e_lo, self.e_hi = local_expert_range(cfg, self.ep, self.rank)
        self.dtype = dtype or (torch.bfloat16 if self.device.type == "cuda" else torch.float32)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.opt_step = 0
        full = tf.TransformerWeights(cfg, device="cpu", dtype=torch.float32, init=True, seed=seed)
        self.load_full(full.embed, full.layers, full.head)

    def is_expert(self, name: str) -> bool:
        return name.split(".")[-1] in self.EXPERT

    def load_full(self, embed: dict, layers: list[dict], head: dict) -> None:
        named = {**{k: v.clone() for k, v in embed.items()}, **{k: v.clone() for k, v in head.items()}}
        for i, w in enumerate(layers):
            for k, v in w.items():
                named[f"layers.{i}.{k}"] = (v[self.e_lo:self.e_hi] if k in self.EXPERT else v).clone()
        self.master = {k: v.to(self.device, torch.float32) for k, v in named.items()}
        self.params = {k: v.to(self.dtype).clone().requires_grad_(True) for k, v in self.master.items()}
        self.exp_avg = {k: torch.zeros_like(v) for k, v in self.master.items()}
        self.exp_avg_sq = {k: torch.zeros_like(v) for k, v in self.master.items()}

    def layer(self, i: int) -> dict:
        pre = f"layers.{i}."
        return {k[len(pre):]: v for k, v in self.params.items() if k.startswith(pre)}

    def forward_hidden(self, inp: tf.PackedInput) -> torch.Tensor:
        x = F.embedding(inp.input_ids, self.params["embed"])
        for i in range(self.cfg.num_layers):
            x = ep_layer_forward(self.cfg, self.layer(i), x, inp, self.group)
        return x

    def forward_logprob_entropy(self, inp: tf.PackedInput, labels: torch.Tensor, temperature: float = 1.0):
        h = ops.rmsnorm(self.forward_hidden(inp), self.params["norm_w"], self.cfg.rms_norm_eps)
        w = self.params["embed"] if self.cfg.tie_word_embeddings else self.params["lm_head"]
        return ops.linear_cross_entropy(h, w, labels, float(temperature))

    def zero_grad(self) -> None:
        for p in self.params.values():
            p.grad = None

    def step(self, lr: float | None = None) -> dict:
        """Dense gradients are averaged over the group (its ranks see different data).  Expert gradients already hold the sum over
        every rank's tokens (they arrived through the dispatch) and are divided by the group size for the same mean-over-ranks."""
        lr = self.lr if lr is None else lr
        grads = {}
        for k, p in self.params.items():
            g = p.grad.float() if p.grad is not None else torch.zeros_like(self.master[k])
            if not self.is_expert(k):
                dist.all_reduce(g, group=self.group)
            grads[k] = g / self.ep
        sq_e = sum((g.pow(2).sum() for k, g in grads.items() if self.is_expert(k)), torch.zeros((), device=self.device))
        sq_d = sum((g.pow(2).sum() for k, g in grads.items() if not self.is_expert(k)), torch.zeros((), device=self.device))
        dist.all_reduce(sq_e, group=self.group)
        norm = float((sq_e + sq_d).sqrt().item())
        self.last_grads = grads                    # mean gradients of the step just taken (tests / diagnostics)
        skipped = not math.isfinite(norm)
        if not skipped:
            self.opt_step += 1
            coef = min(1.0, self.max_grad_norm / (norm + 1e-6)) if self.max_grad_norm > 0 else 1.0
            b1, b2 = self.betas
            with torch.no_grad():
                for k, g in grads.items():
                    ref.adamw_step(self.master[k], self.exp_avg[k], self.exp_avg_sq[k], g, lr, b1, b2, self.eps, self.weight_decay,
                                   self.opt_step, coef)
                    self.params[k].copy_(self.master[k].to(self.dtype))
        self.zero_grad()
        return {"grad_norm": norm, "skipped": skipped}

    def gather_full_state(self) -> dict[str, torch.Tensor]:
        out = {}
        for k, v in self.master.items():
            if self.is_expert(k):
                parts = [torch.empty_like(v) for _ in range(self.ep)]
                dist.all_gather(parts, v.contiguous(), group=self.group)
                out[k] = torch.cat(parts, dim=0)
            else:
                out[k] = v.clone()
        return out
