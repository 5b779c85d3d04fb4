"""Tensor-parallel TRAINING (Megatron-style column / row parallel blocks + vocab-parallel fused linear-CE), composable with
data parallelism: a (dp x tp) mesh where every TP rank owns 1/tp of each block's projections, fp32 master weights and Adam
moments (optimizer memory / tp), gradients are averaged over the DP group.

The PolyRL stream path only trains with FSDP; TP/PP training exists in the reference through verl's Megatron backend
(3rdparty/verl/verl/workers/megatron_workers.py:204-211, verl/utils/megatron/tensor_parallel.py:109-186 vocab-parallel
entropy, Triton linear-CE TP mode verl/utils/kernel/kernels.py:642-694).  This module is the B200 take on that capability for
models whose blocks are too wide for one GPU's activations: the same fused kernels (tcgen05 GEMMs, fused SwiGLU, fused
linear-CE) run on the local slices, the two all-reduces per block are autograd functions, the LM head is vocab-parallel.

Equivalence contract (tests/tp_train_worker.py): TP=2 (x DP) training steps == single-process training on the same data.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from .. import ops
from ..models import transformer as tf
from ..models.config import ModelConfig
from ..ops import reference as ref
from ..ops.attention import attention_varlen


class _CopyToTP(torch.autograd.Function):
    """Identity forward, all-reduce backward (input of a column-parallel linear)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, group=ctx.group)
        return g, None


class _ReduceFromTP(torch.autograd.Function):
    """All-reduce forward, identity backward (output of a row-parallel linear)."""

    @staticmethod
    def forward(ctx, x, group):
        x = x.contiguous()
        dist.all_reduce(x, group=group)
        return x

    @staticmethod
    def backward(ctx, g):
        return g, None


def _all_gather_rows(x: torch.Tensor, group) -> torch.Tensor:
    parts = [torch.empty_like(x) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts, dim=0)


def _reduce_scatter_rows(x: torch.Tensor, group) -> torch.Tensor:
    tp, r = dist.get_world_size(group), dist.get_rank(group)
    n = x.shape[0] // tp
    if x.is_cuda:
        out = x.new_empty(n, *x.shape[1:])
        dist.reduce_scatter_tensor(out, x.contiguous(), group=group)
        return out
    x = x.contiguous().clone()                   # gloo has no reduce-scatter: all-reduce and keep this rank's rows
    dist.all_reduce(x, group=group)
    return x[r * n:(r + 1) * n].contiguous()


class _GatherFromSP(torch.autograd.Function):
    """Megatron sequence parallelism: activations between the blocks' GEMM regions are sharded by token rows.  Entering a
    column-parallel region: all-gather rows forward, reduce-scatter the gradient backward."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return _all_gather_rows(x, group)

    @staticmethod
    def backward(ctx, g):
        return _reduce_scatter_rows(g, ctx.group), None


class _ScatterToSP(torch.autograd.Function):
    """Leaving a row-parallel region: reduce-scatter rows forward (replaces the all-reduce), all-gather the gradient backward."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return _reduce_scatter_rows(x, group)

    @staticmethod
    def backward(ctx, g):
        return _all_gather_rows(g, ctx.group), None


class _NoReduceBackward(torch.autograd.Function):
    """Identity whose gradient is divided by tp: placed between an all-gather (whose backward SUMS the ranks' gradients) and a
    consumer whose backward already delivers the complete, identical gradient on every rank (the vocab-parallel head)."""

    @staticmethod
    def forward(ctx, x, tp):
        ctx.tp = tp
        return x

    @staticmethod
    def backward(ctx, g):
        return g / ctx.tp, None


def local_config(cfg: ModelConfig, tp: int) -> ModelConfig:
    if cfg.num_experts:
        raise NotImplementedError("tensor parallelism over mixture-of-experts blocks (shard the experts with ZeRO or pipeline stages)")
    if cfg.num_heads % tp or cfg.num_kv_heads % tp or (cfg.intermediate_size // 128) % tp or cfg.vocab_size % tp:
        raise ValueError("heads, kv heads, intermediate_size/128 and vocab_size must be divisible by the TP size")
    d = cfg.to_dict()
    d.update(num_heads=cfg.num_heads // tp, num_kv_heads=cfg.num_kv_heads // tp, intermediate_size=cfg.intermediate_size // tp,
             head_dim=cfg.head_dim)
    out = ModelConfig(**d)
    out.hidden_size = cfg.hidden_size
    return out


def shard_layer(cfg: ModelConfig, w: dict, rank: int, tp: int) -> dict:
    """Column-parallel qkv / gate-up (rows), row-parallel o / down (columns); norms replicated."""
    d, nq, nkv = cfg.head_dim, cfg.num_heads // tp, cfg.num_kv_heads // tp
    q0, k0, v0 = rank * nq * d, cfg.q_size + rank * nkv * d, cfg.q_size + cfg.kv_size + rank * nkv * d
    out = {"ln1": w["ln1"].clone(), "ln2": w["ln2"].clone()}
    for k in ("q_norm", "k_norm", "ln_attn", "ln_mlp"):
        if k in w:
            out[k] = w[k].clone()
    out["wqkv"] = torch.cat([w["wqkv"][q0:q0 + nq * d], w["wqkv"][k0:k0 + nkv * d], w["wqkv"][v0:v0 + nkv * d]]).contiguous()
    if "bqkv" in w:
        out["bqkv"] = torch.cat([w["bqkv"][q0:q0 + nq * d], w["bqkv"][k0:k0 + nkv * d], w["bqkv"][v0:v0 + nkv * d]]).contiguous()
    out["wo"] = w["wo"][:, q0:q0 + nq * d].contiguous()
    f2 = 2 * cfg.intermediate_size // tp
    out["wgu"] = w["wgu"][rank * f2:(rank + 1) * f2].contiguous()
    fl = cfg.intermediate_size // tp
    out["wd"] = w["wd"][:, rank * fl:(rank + 1) * fl].contiguous()
    return out


def tp_layer_forward(lcfg: ModelConfig, w: dict, x: torch.Tensor, inp: tf.PackedInput, group, sequence_parallel: bool = False) -> torch.Tensor:
    """One decoder block on this rank's head / channel slice.  x is replicated across the TP group, or — ``sequence_parallel`` —
    holds this rank's T/tp token rows: norms and residual adds then run on 1/tp of the tokens, the all-reduces become an
    all-gather into / a reduce-scatter out of each GEMM region (same bytes on the wire, 1/tp of the norm/residual activations)."""
    nq, nkv, d = lcfg.num_heads, lcfg.num_kv_heads, lcfg.head_dim
    # GLM-4: the summed projection outputs are normalised before the residual add
    post = (lambda y, k: ops.rmsnorm(y, w[k], lcfg.rms_norm_eps)) if lcfg.sandwich_norm else (lambda y, k: y)
    if sequence_parallel:
        h = _GatherFromSP.apply(ops.rmsnorm(x, w["ln1"], lcfg.rms_norm_eps), group)
        t = h.shape[0]
        qkv = tf.qk_norm_(lcfg, w, ops.linear(h, w["wqkv"], bias=w.get("bqkv")))
        pos, cos, sin = inp.rope_inputs(lcfg, x.device)
        qkv = ops.apply_rope(qkv, pos, cos, sin, nq, nkv, d)
        a = attention_varlen(qkv[:, : nq * d].view(t, nq, d), qkv[:, nq * d:(nq + nkv) * d].view(t, nkv, d),
                             qkv[:, (nq + nkv) * d:].view(t, nkv, d), inp.cu_seqlens, inp.max_seqlen, causal=True).reshape(t, nq * d)
        x = x + post(_ScatterToSP.apply(ops.linear(a, w["wo"]), group), "ln_attn")
        h = _GatherFromSP.apply(ops.rmsnorm(x, w["ln2"], lcfg.rms_norm_eps), group)
        return x + post(_ScatterToSP.apply(ops.linear(ops.linear_swiglu(h, w["wgu"]), w["wd"]), group), "ln_mlp")
    t = x.shape[0]
    h = _CopyToTP.apply(ops.rmsnorm(x, w["ln1"], lcfg.rms_norm_eps), group)
    qkv = tf.qk_norm_(lcfg, w, ops.linear(h, w["wqkv"], bias=w.get("bqkv")))
    pos, cos, sin = inp.rope_inputs(lcfg, x.device)
    qkv = ops.apply_rope(qkv, pos, cos, sin, nq, nkv, d)
    q = qkv[:, : nq * d].view(t, nq, d)
    k = qkv[:, nq * d:(nq + nkv) * d].view(t, nkv, d)
    v = qkv[:, (nq + nkv) * d:].view(t, nkv, d)
    a = attention_varlen(q, k, v, inp.cu_seqlens, inp.max_seqlen, causal=True).reshape(t, nq * d)
    x = x + post(_ReduceFromTP.apply(ops.linear(a, w["wo"]), group), "ln_attn")
    h = _CopyToTP.apply(ops.rmsnorm(x, w["ln2"], lcfg.rms_norm_eps), group)
    a = ops.linear_swiglu(h, w["wgu"])
    return x + post(_ReduceFromTP.apply(ops.linear(a, w["wd"]), group), "ln_mlp")


class TPTrainEngine:
    """(dp x tp) training engine for the llama family.  Parameters live as fp32 masters + compute-dtype copies of the LOCAL
    slices; ``step()`` averages gradients over ``dp_group``, clips by the global norm (sharded params counted once per TP
    rank, replicated params once) and applies AdamW."""

    REPLICATED = ("ln1", "ln2", "ln_attn", "ln_mlp", "embed", "norm_w", "q_norm", "k_norm")

    def __init__(self, cfg: ModelConfig, device, tp_group, dp_group=None, lr: float = 1e-6, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.01, max_grad_norm: float = 1.0, seed: int = 0, dtype=None, sequence_parallel: bool = False):
        if cfg.family != "llama":
            raise NotImplementedError("TP training is implemented for the llama / qwen block family")
        if cfg.mlp_act != "swiglu":
            raise NotImplementedError("TP training is implemented for SwiGLU MLPs (the xIELU family trains on the ZeRO / pipeline engines)")
        if cfg.attn_type != "mha":
            raise NotImplementedError("TP training is not implemented for MLA attention (use the ZeRO, pipeline or expert-parallel engine)")
        self.sequence_parallel = sequence_parallel      # Megatron SP (reference: megatron.yaml sequence_parallel, utils/megatron/sequence_parallel.py)
        self.cfg, self.device = cfg, torch.device(device)
        self.tp_group, self.dp_group = tp_group, dp_group
        self.tp = dist.get_world_size(tp_group)
        self.tp_rank = dist.get_rank(tp_group)
        self.dp = dist.get_world_size(dp_group) if dp_group is not None else 1
        self.lcfg = local_config(cfg, self.tp)
        self.dtype = dtype or (torch.bfloat16 if self.device.type == "cuda" else torch.float32)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.opt_step = 0
        # same initialisation as the ZeRO engine / TransformerWeights (seeded full tensors), then sliced
        full = tf.TransformerWeights(cfg, device="cpu", dtype=torch.float32, init=True, seed=seed)
        self.load_full(full.embed, full.layers, full.head)

    # ------------------------------------------------------------------ parameters
    def load_full(self, embed: dict, layers: list[dict], head: dict) -> None:
        cfg, r, tp = self.cfg, self.tp_rank, self.tp
        vs = cfg.vocab_size // tp
        self.vocab_start = r * vs
        named: dict[str, torch.Tensor] = {"embed": embed["embed"].clone(), "norm_w": head["norm_w"].clone()}
        lm = embed["embed"] if cfg.tie_word_embeddings else head["lm_head"]
        named["lm_head"] = lm[r * vs:(r + 1) * vs].clone()                       # vocab-parallel (also when tied: own copy)
        for i, w in enumerate(layers):
            for k, v in shard_layer(cfg, w, r, tp).items():
                named[f"layers.{i}.{k}"] = v
        self.master = {k: v.to(self.device, torch.float32) for k, v in named.items()}
        self.params = {k: v.to(self.dtype).clone().requires_grad_(True) for k, v in self.master.items()}
        self.exp_avg = {k: torch.zeros_like(v) for k, v in self.master.items()}
        self.exp_avg_sq = {k: torch.zeros_like(v) for k, v in self.master.items()}

    def layer(self, i: int) -> dict:
        pre = f"layers.{i}."
        return {k[len(pre):]: v for k, v in self.params.items() if k.startswith(pre)}

    def is_replicated(self, name: str) -> bool:
        return name.split(".")[-1] in self.REPLICATED

    # ------------------------------------------------------------------ forward
    def forward_hidden(self, inp: tf.PackedInput) -> torch.Tensor:
        if not self.sequence_parallel:
            x = torch.nn.functional.embedding(inp.input_ids, self.params["embed"])
            for i in range(self.cfg.num_layers):
                x = tp_layer_forward(self.lcfg, self.layer(i), x, inp, self.tp_group)
            return x
        # sequence parallel: pad the packed stream to a multiple of tp (the pad rows form a dummy tail sequence for attention), embed
        # and keep this rank's rows; the final norm runs on the local rows, then the rows are gathered for the vocab-parallel head
        t = inp.input_ids.shape[0]
        pad = (-t) % self.tp
        ids, full = inp.input_ids, inp
        if pad:
            ids = torch.cat([ids, ids.new_full((pad,), self.cfg.pad_token_id)])
            pos = torch.cat([inp.position_ids, inp.position_ids.new_zeros(pad)])
            cu = torch.cat([inp.cu_seqlens, inp.cu_seqlens.new_tensor([t + pad])])
            full = tf.PackedInput(ids, pos, cu, max(inp.max_seqlen, pad))
        n = ids.shape[0] // self.tp
        x = torch.nn.functional.embedding(ids[self.tp_rank * n:(self.tp_rank + 1) * n], self.params["embed"])
        for i in range(self.cfg.num_layers):
            x = tp_layer_forward(self.lcfg, self.layer(i), x, full, self.tp_group, sequence_parallel=True)
        self._sp_rows = (t, pad)
        return x

    def forward_logprob_entropy(self, inp: tf.PackedInput, labels: torch.Tensor, temperature: float = 1.0):
        if self.sequence_parallel:
            x = self.forward_hidden(inp)
            h = _GatherFromSP.apply(ops.rmsnorm(x, self.params["norm_w"], self.cfg.rms_norm_eps), self.tp_group)[: self._sp_rows[0]]
            # the gathered rows feed a vocab-parallel head whose backward already all-reduces dH: scale-free hand-over
            return ops.linear_cross_entropy_tp(_NoReduceBackward.apply(h, self.tp), self.params["lm_head"], labels, float(temperature),
                                               self.tp_group, self.vocab_start)
        h = ops.rmsnorm(self.forward_hidden(inp), self.params["norm_w"], self.cfg.rms_norm_eps)
        # vocab-parallel fused linear-CE: its backward all-reduces dH over the TP group, so every rank continues with the
        # complete (identical) gradient — no extra copy-to-TP wrapper here
        return ops.linear_cross_entropy_tp(h, self.params["lm_head"], labels, float(temperature), self.tp_group, self.vocab_start)

    # ------------------------------------------------------------------ optimizer
    def zero_grad(self) -> None:
        for p in self.params.values():
            p.grad = None

    def step(self, lr: float | None = None) -> dict:
        lr = self.lr if lr is None else lr
        grads = {}
        for k, p in self.params.items():
            g = p.grad.float() if p.grad is not None else torch.zeros_like(self.master[k])
            if self.sequence_parallel and self.is_replicated(k):
                dist.all_reduce(g, group=self.tp_group)       # norms / embedding saw only this rank's token rows
            if self.dp > 1:
                dist.all_reduce(g, group=self.dp_group)
                g /= self.dp
            grads[k] = g
        # tied embeddings: the input embedding also receives the (vocab-parallel) head gradient
        if self.cfg.tie_word_embeddings:
            vs = self.cfg.vocab_size // self.tp
            full = torch.zeros_like(grads["embed"])
            full[self.vocab_start:self.vocab_start + vs] = grads["lm_head"]
            dist.all_reduce(full, group=self.tp_group)
            grads["embed"] = grads["embed"] + full
            grads["lm_head"] = grads["embed"][self.vocab_start:self.vocab_start + vs].clone()
        sq_shard = sum((g.pow(2).sum() for k, g in grads.items() if not self.is_replicated(k) and not (self.cfg.tie_word_embeddings and k == "lm_head")),
                       torch.zeros((), device=self.device))
        sq_rep = sum((g.pow(2).sum() for k, g in grads.items() if self.is_replicated(k)), torch.zeros((), device=self.device))
        dist.all_reduce(sq_shard, group=self.tp_group)
        norm = float((sq_shard + sq_rep).sqrt().item())
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

    # ------------------------------------------------------------------ export
    def gather_full_state(self) -> dict[str, torch.Tensor]:
        """Reassemble the unsharded parameters (same naming as ``Zero3Engine.named_full_tensors``) on every rank."""
        cfg, tp = self.cfg, self.tp
        d, nq, nkv = cfg.head_dim, cfg.num_heads // tp, cfg.num_kv_heads // tp

        def gather(t: torch.Tensor) -> list[torch.Tensor]:
            parts = [torch.empty_like(t) for _ in range(tp)]
            dist.all_gather(parts, t.contiguous(), group=self.tp_group)
            return parts

        out = {"embed": self.master["embed"].clone(), "norm_w": self.master["norm_w"].clone()}
        if not cfg.tie_word_embeddings:
            out["lm_head"] = torch.cat(gather(self.master["lm_head"]))
        for i in range(cfg.num_layers):
            pre = f"layers.{i}."
            out[pre + "ln1"], out[pre + "ln2"] = self.master[pre + "ln1"].clone(), self.master[pre + "ln2"].clone()
            for k in ("q_norm", "k_norm", "ln_attn", "ln_mlp"):
                if pre + k in self.master:
                    out[pre + k] = self.master[pre + k].clone()
            qkv = gather(self.master[pre + "wqkv"])
            out[pre + "wqkv"] = torch.cat([torch.cat([p[: nq * d] for p in qkv]), torch.cat([p[nq * d:(nq + nkv) * d] for p in qkv]),
                                           torch.cat([p[(nq + nkv) * d:] for p in qkv])])
            if pre + "bqkv" in self.master:
                b = gather(self.master[pre + "bqkv"])
                out[pre + "bqkv"] = torch.cat([torch.cat([p[: nq * d] for p in b]), torch.cat([p[nq * d:(nq + nkv) * d] for p in b]),
                                               torch.cat([p[(nq + nkv) * d:] for p in b])])
            out[pre + "wo"] = torch.cat(gather(self.master[pre + "wo"]), dim=1)
            out[pre + "wgu"] = torch.cat(gather(self.master[pre + "wgu"]))
            out[pre + "wd"] = torch.cat(gather(self.master[pre + "wd"]), dim=1)
        return out
