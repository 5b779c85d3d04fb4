"""Tensor-parallel rollout: weight slicing, the fused GEMM -> all-reduce over NVLink peer memory, and lock-step
driving of the TP ranks.

Megatron-style sharding (what SGLang does for the reference's ``rollout.tensor_model_parallel_size``,
rlboost/verl_stream/workers/rollout/sglang_rollout/sglang_rollout_remote.py:152): QKV and gate/up are
column-parallel (each rank owns nq/W query heads, nkv/W kv heads and F/W intermediate channels — its KV cache
only holds its heads), O and down projections are row-parallel and need an all-reduce.  The reference stack
issues that all-reduce as a separate NCCL / custom-all-reduce kernel after a cuBLAS GEMM; here the decode path
uses ``skinny_gemm_tp``: the GEMM's epilogue pushes its partial tile into every peer's inbox with NVLink
stores, signals per-tile flags and reduces in rank order — one kernel, no NCCL (SURVEY §2.5c "SGLang TP
all-reduce x2 per layer per decode step").  Prefill (large M) uses the NCCL all-reduce.

Lock-step: all TP ranks run identical engines on identical request streams with identical seeds; hidden states
are bit-identical after the rank-ordered reduction, so sampling agrees without any host communication.  Rank 0
broadcasts the request/abort commands it receives over a gloo control group.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .. import ops
from ..models import transformer as tf
from ..models.config import ModelConfig
from ..ops.attention import attention_varlen
from ..ops.functional import rope_and_cache_, _skinny_buffers


class TPContext:
    def __init__(self, cfg: ModelConfig, heap, tp_rank: int, tp_world: int, max_batch: int, group=None):
        if cfg.num_heads % tp_world or cfg.num_kv_heads % tp_world:
            raise ValueError("heads and kv heads must be divisible by the TP size")
        if cfg.num_experts:
            if cfg.num_experts % tp_world:
                raise ValueError(f"{cfg.num_experts} experts cannot be split over {tp_world} TP ranks")
        elif (cfg.intermediate_size // 128) % tp_world:
            raise ValueError("intermediate_size/128 must be divisible by the TP size")
        self.cfg, self.heap, self.rank, self.world, self.group = cfg, heap, tp_rank, tp_world, group
        self.max_batch = max_batch
        if heap is None:            # CPU / gloo (plumbing tests) or explicit library mode: all-reduce through torch.distributed
            self.ar_mode, self.C = "nccl", None
            return
        bp = ((min(max_batch, 256) + 15) // 16) * 16
        tiles = (cfg.hidden_size + 127) // 128
        self.inbox_half = tp_world * tiles * bp * 128
        self.inbox = heap.alloc(2 * self.inbox_half, torch.float32)
        self.flags = heap.alloc(tp_world * tiles + 64, torch.int32)
        self.epoch_state = torch.zeros(2, device=heap.device, dtype=torch.int32)   # [epoch, done-counter] (device side)
        self.C = ops.ext()
        # row-parallel all-reduce on the decode path (measured, profiles/r1_tp2_rowparallel_*.json):
        #   "oneshot" (default): library GEMM straight into a symmetric buffer -> device barrier -> one-shot peer-memory
        #             all-reduce kernel (every rank reads all partials over NVLink, rank-ordered fp32 sum, fused residual)
        #   "fused":  the swap-AB split-K GEMM whose epilogue pushes fp32 tiles into the peers' inboxes (one kernel)
        #   "nccl":   library GEMM + NCCL all-reduce (baseline)
        import os

        #   "nvls":   library GEMM into a multicast-mapped buffer -> ONE kernel: rank barrier + `multimem.ld_reduce` of the whole
        #             buffer inside the NVSwitch (+ fused residual); needs the VMM/multicast heap (parallel/nvls.py)
        self.ar_mode = os.environ.get("POLYRL_TP_ALLREDUCE", "auto")
        rows = ((min(max_batch, 256) + 7) // 8) * 8
        self.nvls_ar = None
        if self.ar_mode == "auto" and tp_world <= 2:
            self.ar_mode = "oneshot"            # measured at TP=2: peer-pointer one-shot 18.2 us vs multimem 21.8 us (profiles/)
        if self.ar_mode in ("auto", "nvls"):
            if getattr(heap, "multicast", False):
                from ..parallel.nvls import NvlsOneShotAllReduce

                self.nvls_ar = NvlsOneShotAllReduce(heap, rows * cfg.hidden_size)
                self.ar_mode = "nvls"
            elif self.ar_mode == "nvls":
                raise RuntimeError("POLYRL_TP_ALLREDUCE=nvls needs a multicast-capable symmetric heap (parallel/nvls.py)")
            else:
                self.ar_mode = "oneshot"
        self.ar_bufs = [heap.alloc(rows * cfg.hidden_size, torch.bfloat16) for _ in range(2)]   # alternate: see row_parallel
        self.ar_turn = 0
        heap.barrier()

    def local_cfg(self) -> ModelConfig:
        """Model config as seen by one TP rank (for KV-cache sizing and the attention kernels)."""
        c = self.cfg
        d = c.to_dict()
        # mixture of experts: attention is head-sharded as usual, the experts are SHARDED WHOLE over the TP ranks (rank r owns
        # experts [r E/W, (r+1) E/W)); activations are replicated inside a TP group, so no dispatch all-to-all is needed — every
        # rank runs the tokens routed to its experts and the partial outputs meet in the same all-reduce a dense row-parallel
        # down-projection uses
        d.update(num_heads=c.num_heads // self.world, num_kv_heads=c.num_kv_heads // self.world,
                 intermediate_size=c.intermediate_size if c.num_experts else c.intermediate_size // self.world, head_dim=c.head_dim)
        out = ModelConfig(**d)
        out.hidden_size = c.hidden_size
        return out

    # ---------------------------------------------------------------- weight slicing
    def shard_layer(self, w: dict) -> dict:
        """Slice one layer's full weights for this rank.  Row/column slices that are contiguous (or strided with a
        row pitch TMA can express) are views; the fused qkv needs a compact copy (three disjoint row ranges)."""
        c, r, W = self.cfg, self.rank, self.world
        d, nq, nkv = c.head_dim, c.num_heads // W, c.num_kv_heads // W
        q0, k0, v0 = r * nq * d, c.q_size + r * nkv * d, c.q_size + c.kv_size + r * nkv * d
        out = {"ln1": w["ln1"], "ln2": w["ln2"]}
        for k in ("q_norm", "k_norm", "ln_attn", "ln_mlp"):   # per-head gains (Qwen3) and output norms (GLM-4): replicated
            if k in w:
                out[k] = w[k]
        if c.attn_type == "mla":
            # latent attention: the heads are split (query / key-value up-projections by head rows, output projection by head
            # columns); the latent bottlenecks — and therefore the paged latent cache — are replicated on every TP rank
            up, dv = c.qk_nope_head_dim + c.v_head_dim, c.v_head_dim
            out["kv_a_ln"] = w["kv_a_ln"]
            out["wkv_up"] = w["wkv_up"][r * nq * up:(r + 1) * nq * up]
            out["wo"] = w["wo"][:, r * nq * dv:(r + 1) * nq * dv]
            if c.q_lora_rank:
                out["wa"], out["q_a_ln"] = w["wa"], w["q_a_ln"]
                out["wq_up"] = w["wq_up"][q0:q0 + nq * d]
            else:                                                     # plain query projection fused into ``wa``: local head rows + the latent rows
                out["wa"] = torch.cat([w["wa"][q0:q0 + nq * d], w["wa"][c.q_size:]], dim=0).contiguous()
        else:
            out["wqkv"] = torch.cat([w["wqkv"][q0:q0 + nq * d], w["wqkv"][k0:k0 + nkv * d], w["wqkv"][v0:v0 + nkv * d]], dim=0).contiguous()
            if "bqkv" in w:
                out["bqkv"] = torch.cat([w["bqkv"][q0:q0 + nq * d], w["bqkv"][k0:k0 + nkv * d], w["bqkv"][v0:v0 + nkv * d]]).contiguous()
            out["wo"] = w["wo"][:, q0:q0 + nq * d]                      # [H, nq_local*d] view, row pitch = full
        if "router" in w:                                            # a sparse block (MoE stacks may also hold dense blocks)
            el = c.num_experts // W
            out["router"] = w["router"]                               # replicated: every rank routes every token identically
            if "router_bias" in w:
                out["router_bias"] = w["router_bias"]
            out["ewgu"], out["ewd"] = w["ewgu"][r * el:(r + 1) * el], w["ewd"][r * el:(r + 1) * el]
            if "swgu" in w:                                           # shared expert: column / row parallel like a dense MLP
                fs = c.shared_expert_intermediate_size // W
                if c.shared_expert_intermediate_size % (128 * W) == 0:
                    out["swgu"], out["swd"] = w["swgu"][r * 2 * fs:(r + 1) * 2 * fs], w["swd"][:, r * fs:(r + 1) * fs]
                elif r == 0:                                          # too narrow to split into whole 128-row blocks: rank 0 runs it
                    out["swgu"], out["swd"] = w["swgu"], w["swd"]
                if "sgate" in w and "swgu" in out:
                    out["sgate"] = w["sgate"]
            return out
        if "wup" in w:                                               # xIELU MLP: element-wise activation, plain column / row split
            fl = c.intermediate_size // W
            out["wup"], out["wd"] = w["wup"][r * fl:(r + 1) * fl], w["wd"][:, r * fl:(r + 1) * fl]
            out["act_ap"], out["act_an"] = w["act_ap"], w["act_an"]
            return out
        f2 = 2 * c.intermediate_size // W
        out["wgu"] = w["wgu"][r * f2:(r + 1) * f2]                  # contiguous rows (block-interleaved layout)
        fl = c.intermediate_size // W
        out["wd"] = w["wd"][:, r * fl:(r + 1) * fl]                 # [H, F_local] view
        return out

    # ---------------------------------------------------------------- row-parallel linear
    def row_parallel(self, x_local: torch.Tensor, w_local: torch.Tensor, residual: torch.Tensor | None,
                     decode: bool = True) -> torch.Tensor:
        b = x_local.shape[0]
        small = self.heap is not None and x_local.is_cuda and decode and b <= 256 and b <= self.max_batch
        if small and self.ar_mode == "nvls":
            h = w_local.shape[0]
            part = self.nvls_ar.next_input(b * h).view(b, h)
            torch.mm(x_local, w_local.t(), out=part)
            out = torch.empty(b, h, device=x_local.device, dtype=torch.bfloat16)
            return self.nvls_ar.reduce(out, residual.contiguous() if residual is not None else None)
        if small and self.ar_mode == "oneshot":
            # Two symmetric buffers alternate: a rank may overwrite buffer p for call c+2 only after passing the barrier of
            # call c+1, which every peer reaches after it finished READING buffer p in call c.  (Decode issues an even
            # number of calls per step, so captured graphs replay a consistent alternation.)
            buf = self.ar_bufs[self.ar_turn]
            self.ar_turn ^= 1
            h = w_local.shape[0]
            part = buf.local[: b * h].view(b, h)
            torch.mm(x_local, w_local.t(), out=part)
            self.heap.barrier()
            out = torch.empty(b, h, device=x_local.device, dtype=torch.bfloat16)
            self.C.p2p_allreduce(buf.ptrs, 0, out, residual.contiguous() if residual is not None else None, 0)
            return out
        if small and self.ar_mode == "fused":
            ws, ctr = _skinny_buffers(x_local.device)
            res = residual.contiguous() if residual is not None else None
            return self.C.skinny_gemm_tp(x_local.contiguous(), w_local, res, ws, ctr, 0, self.inbox.ptrs, self.flags.ptrs,
                                         self.rank, self.epoch_state, self.inbox_half)
        y = torch.nn.functional.linear(x_local, w_local)
        dist.all_reduce(y, group=self.group)
        return y + residual if residual is not None else y


def allreduce_partial(tp: TPContext, part: torch.Tensor, residual: torch.Tensor | None, decode: bool) -> torch.Tensor:
    """Sum a [b, H] partial result over the TP ranks (+ residual) through the same transports as ``row_parallel``."""
    b, h = part.shape
    small = tp.heap is not None and part.is_cuda and decode and b <= 256 and b <= tp.max_batch
    if small and tp.ar_mode == "nvls":
        tp.nvls_ar.next_input(b * h).view(b, h).copy_(part)
        out = torch.empty(b, h, device=part.device, dtype=torch.bfloat16)
        return tp.nvls_ar.reduce(out, residual.contiguous() if residual is not None else None)
    if small and tp.ar_mode in ("oneshot", "fused"):
        buf = tp.ar_bufs[tp.ar_turn]
        tp.ar_turn ^= 1
        buf.local[: b * h].view(b, h).copy_(part)
        tp.heap.barrier()
        out = torch.empty(b, h, device=part.device, dtype=torch.bfloat16)
        tp.C.p2p_allreduce(buf.ptrs, 0, out, residual.contiguous() if residual is not None else None, 0)
        return out
    y = part.contiguous()
    dist.all_reduce(y, group=tp.group)
    return y + residual if residual is not None else y


def moe_partial(tp: TPContext, cfg: ModelConfig, w: dict, h: torch.Tensor, static: bool) -> torch.Tensor:
    """This rank's share of a mixture-of-experts MLP: the (replicated) router picks top-k experts per token, the rank computes
    the assignments that fall on ITS experts and returns the weighted partial sum [T, H] (zero rows for tokens routed elsewhere).
    ``static`` keeps every shape fixed and avoids host reads (CUDA-graph decode): assignments are scattered into an
    [E_local + 1, T, H] buffer whose last row collects the other ranks' assignments and is never multiplied."""
    import torch.nn.functional as F

    from ..ops import reference as ref

    t, k, e = h.shape[0], cfg.num_experts_per_tok, cfg.num_experts
    el = e // tp.world
    e0 = tp.rank * el
    top_w, top_e, _ = tf.moe_route(cfg, w, h)
    flat_e = top_e.reshape(-1)
    order = torch.argsort(flat_e, stable=True)
    sorted_e = flat_e[order]
    tok = torch.div(order, k, rounding_mode="floor")
    gate = top_w.reshape(-1)[order].to(h.dtype)
    mine = (sorted_e >= e0) & (sorted_e < e0 + el)
    if static:
        counts = torch.zeros(e, dtype=torch.long, device=h.device).scatter_add_(0, flat_e, torch.ones_like(flat_e))
        start = counts.cumsum(0) - counts
        slot = torch.arange(t * k, device=h.device) - start[sorted_e]
        row = torch.where(mine, sorted_e - e0, torch.full_like(sorted_e, el))
        x = h.new_zeros(el + 1, t, h.shape[1])
        x[row, slot] = h.index_select(0, tok)
        gu = torch.bmm(x[:el], w["ewgu"].to(h.dtype).transpose(1, 2))
        a = ref.swiglu(gu.reshape(el * t, -1)).reshape(el, t, -1)
        y = torch.bmm(a, w["ewd"].to(h.dtype).transpose(1, 2))
        y = torch.cat([y, y.new_zeros(1, t, h.shape[1])], dim=0)[row, slot] * gate.unsqueeze(-1)
        return _add_shared(w, h, torch.zeros(t, h.shape[1], dtype=y.dtype, device=h.device).index_add(0, tok, y))
    counts = torch.bincount(flat_e, minlength=e).tolist()
    xs = h.index_select(0, tok)
    out = torch.zeros(t, h.shape[1], dtype=h.dtype, device=h.device)
    lo = 0
    for ei, n in enumerate(counts):
        if n and e0 <= ei < e0 + el:
            a = ops.linear_swiglu(xs[lo:lo + n], w["ewgu"][ei - e0])
            y = ops.linear(a, w["ewd"][ei - e0]) * gate[lo:lo + n].unsqueeze(-1)
            out = out.index_add(0, tok[lo:lo + n], y.to(out.dtype))
        lo += n
    return _add_shared(w, h, out)


def _add_shared(w: dict, h: torch.Tensor, part: torch.Tensor) -> torch.Tensor:
    """+ this rank's column slice of the shared expert (its sigmoid gate distributes over the partial sums)."""
    s = tf.shared_expert(w, h)
    return part if s is None else part + s.to(part.dtype)


def _row_out(tp: TPContext, lcfg: ModelConfig, w: dict, a: torch.Tensor, wname: str, norm: str, x: torch.Tensor, decode: bool) -> torch.Tensor:
    """Row-parallel projection + residual; GLM-4 normalises the (all-reduced) projection output before the residual add."""
    if lcfg.sandwich_norm:
        return x + ops.rmsnorm(tp.row_parallel(a, w[wname], None, decode=decode), w[norm], lcfg.rms_norm_eps)
    return tp.row_parallel(a, w[wname], x, decode=decode)


def _mlp_tp(tp: TPContext, lcfg: ModelConfig, w: dict, x: torch.Tensor, decode: bool) -> torch.Tensor:
    h = ops.rmsnorm(x, w["ln2"], lcfg.rms_norm_eps)
    if "router" in w:
        if lcfg.sandwich_norm:
            raise NotImplementedError("mixture-of-experts blocks with output norms under tensor-parallel rollout")
        part = moe_partial(tp, tp.cfg, w, h, static=decode and x.is_cuda and tf._MOE_STATIC_DECODE())
        return allreduce_partial(tp, part.to(x.dtype), x, decode)
    if "wup" in w:
        return _row_out(tp, lcfg, w, tf.xielu(ops.linear(h, w["wup"]), w["act_ap"], w["act_an"]), "wd", "ln_mlp", x, decode)
    return _row_out(tp, lcfg, w, ops.linear_swiglu(h, w["wgu"]), "wd", "ln_mlp", x, decode)


def layer_decode_tp(tp: TPContext, lcfg: ModelConfig, w: dict, x: torch.Tensor, inp: tf.DecodeInput, k_cache, v_cache) -> torch.Tensor:
    nq, nkv, d = lcfg.num_heads, lcfg.num_kv_heads, lcfg.head_dim
    h = ops.rmsnorm(x, w["ln1"], lcfg.rms_norm_eps)
    if lcfg.attn_type == "mla":         # local heads over the replicated latent cache, then the usual row-parallel output projection
        from ..models import mla

        x = _row_out(tp, lcfg, w, mla.attention_decode(lcfg, w, h, inp, k_cache), "wo", "ln_attn", x, True)
        return _mlp_tp(tp, lcfg, w, x, decode=True)
    qkv = tf.qk_norm_(lcfg, w, ops.linear(h, w["wqkv"], bias=w.get("bqkv")))
    cos, sin = tf.rope_tables(lcfg, x.device)
    rope_and_cache_(qkv, inp.position_ids, cos, sin, nq, nkv, d, k_cache, v_cache, inp.slot_mapping, k_cache.shape[2])
    a = ops.paged_decode_attention(qkv[:, : nq * d], k_cache, v_cache, inp.block_table, inp.seq_lens, nq, nkv, d, d ** -0.5,
                                   max_len=inp.max_len)
    x = _row_out(tp, lcfg, w, a, "wo", "ln_attn", x, True)
    return _mlp_tp(tp, lcfg, w, x, decode=True)


def layer_prefill_tp(tp: TPContext, lcfg: ModelConfig, w: dict, x: torch.Tensor, inp: tf.PackedInput, kv_sink) -> torch.Tensor:
    nq, nkv, d = lcfg.num_heads, lcfg.num_kv_heads, lcfg.head_dim
    t = x.shape[0]
    h = ops.rmsnorm(x, w["ln1"], lcfg.rms_norm_eps)
    if lcfg.attn_type == "mla":
        from ..models import mla

        x = _row_out(tp, lcfg, w, mla.attention_forward(lcfg, w, h, inp, kv_sink), "wo", "ln_attn", x, False)
        return _mlp_tp(tp, lcfg, w, x, decode=False)
    qkv = tf.qk_norm_(lcfg, w, ops.linear(h, w["wqkv"], bias=w.get("bqkv")))
    pos, cos, sin = inp.rope_inputs(lcfg, x.device)              # 1-D positions, or per-token M-RoPE tables (VLM prompts)
    qkv = rope_and_cache_(qkv, pos, cos, sin, nq, nkv, d, *kv_sink)
    q = qkv[:, : nq * d].view(t, nq, d)
    k = qkv[:, nq * d:(nq + nkv) * d].view(t, nkv, d)
    v = qkv[:, (nq + nkv) * d:].view(t, nkv, d)
    a = attention_varlen(q, k, v, inp.cu_seqlens, inp.max_seqlen, causal=True).reshape(t, nq * d)
    x = _row_out(tp, lcfg, w, a, "wo", "ln_attn", x, False)
    return _mlp_tp(tp, lcfg, w, x, decode=False)
