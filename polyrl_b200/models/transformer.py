"""Functional decoder-only transformer over packed (remove-padding) token streams.

The model is a set of *parameter specs* plus pure functions of (config, weights, activations); it owns
no storage.  Storage and sharding belong to ``parallel/zero3.py`` (training) and ``rollout/engine.py``
(inference), which hand these functions dict views into their flat buffers — that is what lets the
trainer and the colocated rollout engine alias the same bf16 weights with zero copies.

Forward semantics match the reference's remove-padding path (3rdparty/verl/verl/workers/actor/
dp_actor.py:86-273: unpad -> [1, total_nnz] -> HF decoder with varlen flash-attn -> lm_head), with the
HF per-op kernels replaced by the fused ops in ``polyrl_b200.ops``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from .. import ops
from ..ops import reference as ref
from ..ops.attention import attention_varlen
from ..ops.functional import rope_and_cache_
from .config import ModelConfig

ParamSpec = list[tuple[str, tuple[int, ...]]]


# --------------------------------------------------------------------------------------------
# parameter specs
# --------------------------------------------------------------------------------------------
def is_sparse_layer(cfg: ModelConfig, i: int | None) -> bool:
    """Block i carries a mixture-of-experts MLP (``i=None``: a generic block of the model — sparse for every MoE config)."""
    return bool(cfg.num_experts) and (i is None or i not in cfg.dense_layers)


def layer_param_specs(cfg: ModelConfig, i: int | None = None) -> ParamSpec:
    """Parameters of decoder block ``i`` (MoE stacks may mix sparse and dense blocks: ``cfg.dense_layers``)."""
    h, f = cfg.hidden_size, cfg.intermediate_size
    qkv = cfg.q_size + 2 * cfg.kv_size
    if cfg.family == "llama":
        if cfg.attn_type == "mla":          # latent attention: fused first-stage projection + up-projections (models/mla.py)
            from . import mla

            specs: ParamSpec = [("ln1", (h,)), *mla.param_specs(cfg), ("ln2", (h,))]
        else:
            specs = [("ln1", (h,)), ("wqkv", (qkv, h))]
            if cfg.qkv_bias:
                specs.append(("bqkv", (qkv,)))
            if cfg.qk_norm:
                specs += [("q_norm", (cfg.head_dim,)), ("k_norm", (cfg.head_dim,))]
            specs += [("wo", (h, cfg.q_size)), ("ln2", (h,))]
        if cfg.sandwich_norm:
            specs += [("ln_attn", (h,)), ("ln_mlp", (h,))]
        if is_sparse_layer(cfg, i):    # router + stacked experts (gate/up rows interleaved per expert like the dense ``wgu``)
            fm, fs = cfg.moe_intermediate_size, cfg.shared_expert_intermediate_size
            specs += [("router", (cfg.num_experts, h)), ("ewgu", (cfg.num_experts, 2 * fm, h)), ("ewd", (cfg.num_experts, h, fm))]
            if cfg.moe_router == "sigmoid_group":
                specs.append(("router_bias", (cfg.num_experts,)))        # selection-only score correction (no gradient reaches it)
            if fs:
                specs += [("swgu", (2 * fs, h)), ("swd", (h, fs))]
                if cfg.shared_expert_gate:
                    specs.append(("sgate", (1, h)))
        elif cfg.mlp_act == "xielu":   # Apertus: up -> xIELU (two learned scalars) -> down
            specs += [("wup", (f, h)), ("wd", (h, f)), ("act_ap", (1,)), ("act_an", (1,))]
        else:
            specs += [("wgu", (2 * f, h)), ("wd", (h, f))]
        return specs
    return [("ln1_w", (h,)), ("ln1_b", (h,)), ("wqkv", (qkv, h)), ("bqkv", (qkv,)), ("wo", (h, h)), ("bo", (h,)),
            ("ln2_w", (h,)), ("ln2_b", (h,)), ("wfc", (f, h)), ("bfc", (f,)), ("wproj", (h, f)), ("bproj", (h,))]


def embed_param_specs(cfg: ModelConfig) -> ParamSpec:
    specs: ParamSpec = [("embed", (cfg.vocab_size, cfg.hidden_size))]
    if cfg.family == "gpt2":
        specs.append(("pos_embed", (cfg.max_position_embeddings, cfg.hidden_size)))
    return specs


def head_param_specs(cfg: ModelConfig) -> ParamSpec:
    h = cfg.hidden_size
    specs: ParamSpec = [("norm_w", (h,))]
    if cfg.family == "gpt2":
        specs.append(("norm_b", (h,)))
    if cfg.value_head:
        # row 0 / element 0 are the value projection and its bias (HF ``score`` of a ...ForTokenClassification checkpoint with
        # num_labels = 1); padded to 8 rows for alignment
        specs += [("v_head", (8, h)), ("v_head_b", (8,))]
    elif not cfg.tie_word_embeddings:
        specs.append(("lm_head", (cfg.vocab_size, h)))
    return specs


def init_param_(name: str, t: torch.Tensor, cfg: ModelConfig, generator: torch.Generator | None = None) -> None:
    """Random init (weights ~ N(0, 0.02), norm gains 1, biases 0), in place, any dtype/device."""
    base = name.split(".")[-1]
    if base.startswith("lora_"):                    # peft default: A ~ U(-1/sqrt(in), 1/sqrt(in)), B = 0
        if base.endswith("_B"):
            t.zero_()
        else:
            bound = 1.0 / math.sqrt(t.shape[-1])
            t.uniform_(-bound, bound, generator=generator)
        return
    if base in ("act_ap", "act_an"):                # HF XIELUActivation: softplus^-1 of 0.8 and of 0.8 - beta
        t.fill_(math.log(math.expm1(0.8 if base == "act_ap" else 0.8 - XIELU_BETA)))
        return
    if base in ("ln1", "ln2", "ln_attn", "ln_mlp", "norm_w", "ln1_w", "ln2_w", "q_norm", "k_norm", "merger_ln_w", "q_a_ln", "kv_a_ln"):
        t.fill_(1.0)
    elif base.startswith("b") or base.endswith("_b") or base.startswith("merger_b") or base == "router_bias":
        t.zero_()
    else:
        std = 0.02
        if base in ("wo", "wd", "wproj", "ewd", "swd"):
            std = 0.02 / math.sqrt(2 * cfg.num_layers)
        t.normal_(0.0, std, generator=generator)
        if base == "v_head":
            t[1:].zero_()


# --------------------------------------------------------------------------------------------
# rotary tables (cached per device)
# --------------------------------------------------------------------------------------------
_ROPE_CACHE: dict = {}


def rotary_dim(cfg: ModelConfig) -> int:
    """Features of a head that rotate (``partial_rotary_factor``; the rest passes through unchanged)."""
    if cfg.attn_type == "mla":
        return cfg.qk_rope_head_dim
    return int(cfg.head_dim * cfg.partial_rotary_factor) if cfg.partial_rotary_factor < 1.0 else cfg.head_dim


def rope_feature_perm(cfg: ModelConfig) -> torch.Tensor | None:
    """perm [head_dim] with ours[j] = checkpoint[perm[j]] inside every q / k head, or None when the layouts agree.  Our kernels
    rotate the pair (j, j + d/2) with table column j.  A checkpoint that pairs (2j, 2j+1) (``rope_interleaved``) or rotates only its
    first r features gets its rotating pairs moved to columns j < r/2 and the pass-through features to the columns whose table
    entries are cos = 1, sin = 0.  q . k is invariant under a common permutation of the head features, so attention is unchanged."""
    d, r = cfg.head_dim, rotary_dim(cfg)
    if not cfg.rope_interleaved and r == d:
        return None
    perm = torch.empty(d, dtype=torch.long)
    j = torch.arange(r // 2)
    first, second = (2 * j, 2 * j + 1) if cfg.rope_interleaved else (j, j + r // 2)
    perm[: r // 2], perm[d // 2: d // 2 + r // 2] = first, second
    rest = torch.arange(r, d)
    perm[r // 2: d // 2], perm[d // 2 + r // 2:] = rest[: (d - r) // 2], rest[(d - r) // 2:]
    return perm


def _permute_heads(t: torch.Tensor, perm: torch.Tensor, d: int, inverse: bool = False) -> torch.Tensor:
    """Apply ``perm`` to the feature index inside every head of a [n_heads * d, ...] projection weight or bias."""
    idx = torch.argsort(perm) if inverse else perm
    return t.reshape(-1, d, *t.shape[1:])[:, idx.to(t.device)].reshape(t.shape)


def scaled_inv_freq(cfg: ModelConfig) -> tuple[torch.Tensor, float]:
    """(inverse frequencies [head_dim / 2], attention factor) under ``cfg.rope_scaling`` — the long-context schemes HF checkpoints
    declare (transformers ``modeling_rope_utils``: linear, llama3, yarn).  The factor multiplies cos and sin (YaRN's temperature)."""
    import math

    d, base = rotary_dim(cfg), float(cfg.rope_theta)
    inv = 1.0 / (base ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    rs = cfg.rope_scaling
    if not rs or rs.get("rope_type", "default") == "default":
        return inv, 1.0
    kind, factor = rs["rope_type"], float(rs.get("factor", 1.0))
    if kind == "linear":
        return inv / factor, 1.0
    orig = float(rs.get("original_max_position_embeddings") or cfg.max_position_embeddings)
    if kind == "llama3":
        lo_f, hi_f = float(rs.get("low_freq_factor", 1.0)), float(rs.get("high_freq_factor", 4.0))
        wavelen = 2 * math.pi / inv
        smooth = ((orig / wavelen - lo_f) / (hi_f - lo_f)).clamp(0.0, 1.0)
        scaled = torch.where(wavelen > orig / lo_f, inv / factor, inv)                       # long wavelengths: fully interpolated
        mid = (wavelen <= orig / lo_f) & (wavelen >= orig / hi_f)
        return torch.where(mid, (1 - smooth) * inv / factor + smooth * inv, scaled), 1.0
    if kind == "yarn":
        beta_fast, beta_slow = float(rs.get("beta_fast", 32.0)), float(rs.get("beta_slow", 1.0))

        def corr_dim(rot: float) -> float:
            return d * math.log(orig / (rot * 2 * math.pi)) / (2 * math.log(base))

        low, high = corr_dim(beta_fast), corr_dim(beta_slow)
        if rs.get("truncate", True):
            low, high = math.floor(low), math.ceil(high)
        low, high = max(low, 0), min(high, d - 1)
        if low == high:
            high += 0.001
        ramp = ((torch.arange(d // 2, dtype=torch.float32) - low) / (high - low)).clamp(0.0, 1.0)
        extrap_w = 1.0 - ramp
        inv_s = (inv / factor) * (1 - extrap_w) + inv * extrap_w

        def mscale(scale: float, m: float = 1.0) -> float:
            return 1.0 if scale <= 1 else 0.1 * m * math.log(scale) + 1.0

        af = rs.get("attention_factor")
        if af is None:
            ms, msa = rs.get("mscale"), rs.get("mscale_all_dim")
            af = mscale(factor, ms) / mscale(factor, msa) if (ms and msa) else mscale(factor)
        return inv_s, float(af)
    raise NotImplementedError(f"rope scaling type {kind!r} (supported: linear, llama3, yarn)")


def rope_tables(cfg: ModelConfig, device) -> tuple[torch.Tensor, torch.Tensor]:
    rs_key = tuple(sorted((k, str(v)) for k, v in (cfg.rope_scaling or {}).items()))
    key = (cfg.head_dim, cfg.rope_theta, cfg.max_position_embeddings, str(device), rs_key, rotary_dim(cfg))
    if key not in _ROPE_CACHE:
        if not cfg.rope_scaling and rotary_dim(cfg) == cfg.head_dim:
            _ROPE_CACHE[key] = ref.rope_tables(cfg.max_position_embeddings, cfg.head_dim, cfg.rope_theta, device=device)
        else:
            inv, af = scaled_inv_freq(cfg)
            fr = torch.outer(torch.arange(cfg.max_position_embeddings, dtype=torch.float32), inv)
            cos, sin = fr.cos() * af, fr.sin() * af
            pad = cfg.head_dim // 2 - cos.shape[1]           # pass-through features: identity rotation
            if pad:
                cos, sin = F.pad(cos, (0, pad), value=1.0), F.pad(sin, (0, pad), value=0.0)
            _ROPE_CACHE[key] = (cos.to(device).contiguous(), sin.to(device).contiguous())
    return _ROPE_CACHE[key]


def qk_norm_(cfg: ModelConfig, w: dict, qkv: torch.Tensor) -> torch.Tensor:
    """Qwen3: RMS-normalise every q and k head (over head_dim) with learned gains, before RoPE.  Library ops on purpose (rows of
    128 elements are a poor fit for the row-per-CTA norm kernel); returns a new fused qkv tensor."""
    if "q_norm" not in w:
        return qkv
    t = qkv.shape[0]
    nq, nkv, d = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    q = ref.rmsnorm(qkv[:, : nq * d].reshape(t, nq, d), w["q_norm"], cfg.rms_norm_eps)
    k = ref.rmsnorm(qkv[:, nq * d:(nq + nkv) * d].reshape(t, nkv, d), w["k_norm"], cfg.rms_norm_eps)
    return torch.cat([q.reshape(t, nq * d), k.reshape(t, nkv * d), qkv[:, (nq + nkv) * d:]], dim=-1)


def _MOE_STATIC_DECODE() -> bool:
    """``POLYRL_MOE_STATIC_DECODE=1``: decode steps of MoE models use the static-shape expert path (graph-capturable)."""
    import os

    return os.environ.get("POLYRL_MOE_STATIC_DECODE", "0") == "1"


def moe_load_balance_loss(cfg: ModelConfig, probs: torch.Tensor, top_e: torch.Tensor) -> torch.Tensor:
    """Switch-style auxiliary loss of one block: ``E * sum_e f_e * P_e`` with f_e the fraction of (token, slot) assignments routed to
    expert e and P_e the mean router probability of e (HF ``load_balancing_loss_func`` evaluated per layer)."""
    e = cfg.num_experts
    f = torch.bincount(top_e.reshape(-1), minlength=e).to(probs.dtype) / max(1, top_e.shape[0])      # summed over the k slots
    return e * (f * probs.mean(0)).sum()


def moe_route(cfg: ModelConfig, w: dict, h: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Router of a sparse block on rows h:[T,H] -> (weights [T,k] fp32, expert ids [T,k], scores [T,E] fp32 for the balance loss).
    ``softmax``: fp32 softmax over all experts, top-k, optional renormalisation (Mixtral / Qwen MoE).  ``sigmoid_group``: sigmoid
    scores; experts are CHOSEN on scores + ``router_bias`` inside the ``topk_group`` best of ``n_group`` expert groups (a group's
    rank = sum of its two best biased scores) but WEIGHTED by the unbiased scores, renormalised and scaled by
    ``routed_scaling_factor`` (DeepSeek-V3 / GLM-4.5; HF Glm4MoeMoE.route_tokens_to_experts)."""
    k, e = cfg.num_experts_per_tok, cfg.num_experts
    if cfg.moe_router == "sigmoid_group":
        scores = torch.sigmoid(F.linear(h.float(), w["router"].float()))
        choice = scores + w["router_bias"].float()
        if cfg.n_group > 1:
            grp = choice.view(-1, cfg.n_group, e // cfg.n_group).topk(2, dim=-1)[0].sum(-1)
            keep = torch.zeros_like(grp).scatter_(1, grp.topk(cfg.topk_group, dim=-1)[1], 1.0).bool()
            choice = choice.masked_fill(~keep.unsqueeze(-1).expand(-1, cfg.n_group, e // cfg.n_group).reshape(-1, e), 0.0)
        top_e = choice.topk(k, dim=-1)[1]
        top_w = scores.gather(1, top_e)
        if cfg.norm_topk_prob:
            top_w = top_w / (top_w.sum(-1, keepdim=True) + 1e-20)
        return top_w * cfg.routed_scaling_factor, top_e, scores / scores.sum(-1, keepdim=True)   # balance term on normalised scores
    probs = torch.softmax(F.linear(h, w["router"].to(h.dtype)).float(), dim=-1)
    choice = probs
    if cfg.n_group > 1:        # DeepSeek-V2 group_limited_greedy: experts of the topk_group groups with the best single expert
        grp = probs.view(-1, cfg.n_group, e // cfg.n_group).max(-1).values
        keep = torch.zeros_like(grp).scatter_(1, grp.topk(cfg.topk_group, dim=-1)[1], 1.0).bool()
        choice = probs.masked_fill(~keep.unsqueeze(-1).expand(-1, cfg.n_group, e // cfg.n_group).reshape(-1, e), 0.0)
    top_w, top_e = torch.topk(choice, k, dim=-1)
    if cfg.norm_topk_prob:
        top_w = top_w / top_w.sum(-1, keepdim=True)
    if cfg.routed_scaling_factor != 1.0:
        top_w = top_w * cfg.routed_scaling_factor
    return top_w, top_e, probs


def shared_expert(w: dict, h: torch.Tensor) -> torch.Tensor | None:
    """The always-on expert of a sparse block (None when the block has none); Qwen2-MoE scales it by a per-token sigmoid gate."""
    if "swgu" not in w:
        return None
    y = ops.linear(ops.linear_swiglu(h, w["swgu"]), w["swd"])
    if "sgate" in w:
        y = y * torch.sigmoid(F.linear(h, w["sgate"].to(h.dtype)))
    return y


def moe_mlp(cfg: ModelConfig, w: dict, h: torch.Tensor, aux_out: list | None = None) -> torch.Tensor:
    """Token-choice top-k mixture of SwiGLU experts on rows h:[T,H] -> [T,H] (Qwen3-MoE semantics: fp32 softmax over all experts,
    top-k, optional renormalisation of the k weights).  Tokens are grouped by expert and every expert with work runs the same
    fused projections as a dense MLP (``ops.linear_swiglu`` / ``ops.linear``) on its rows; results are scattered back weighted.
    Reference capability: HF Qwen3MoeSparseMoeBlock through verl's FSDP path (verl/models/transformers/ + fsdp_workers.py)."""
    t = h.shape[0]
    k = cfg.num_experts_per_tok
    top_w, top_e, probs = moe_route(cfg, w, h)                                    # [T, k]
    if aux_out is not None:
        aux_out.append(moe_load_balance_loss(cfg, probs, top_e))
    flat_e = top_e.reshape(-1)
    order = torch.argsort(flat_e, stable=True)                                     # assignments grouped by expert
    tok = torch.div(order, k, rounding_mode="floor")
    counts = torch.bincount(flat_e, minlength=cfg.num_experts).tolist()
    gate_w = top_w.reshape(-1)[order].to(h.dtype)
    xs = h.index_select(0, tok)
    outs, lo = [], 0
    for e, n in enumerate(counts):
        if n == 0:
            continue
        a = ops.linear_swiglu(xs[lo:lo + n], w["ewgu"][e])
        outs.append(ops.linear(a, w["ewd"][e]))
        lo += n
    y = torch.cat(outs, dim=0) * gate_w.unsqueeze(-1)
    y = torch.zeros(t, h.shape[1], dtype=y.dtype, device=h.device).index_add(0, tok, y)
    s = shared_expert(w, h)
    return y if s is None else y + s


def moe_mlp_static(cfg: ModelConfig, w: dict, h: torch.Tensor) -> torch.Tensor:
    """``moe_mlp`` with static shapes and no host synchronisation, so a decode step over a fixed batch can be captured in a CUDA
    graph: assignments are sorted by expert, scattered into an [E, C, H] buffer with capacity C = T (an expert can receive at most
    every token once, nothing is ever dropped) and all experts run as two batched GEMMs.  Padded rows cost FLOPs but no weight
    traffic beyond reading each expert once, which a decode step pays anyway — meant for T <= a few hundred rows."""
    t, k, e = h.shape[0], cfg.num_experts_per_tok, cfg.num_experts
    top_w, top_e, _ = moe_route(cfg, w, h)
    flat_e = top_e.reshape(-1)
    order = torch.argsort(flat_e, stable=True)
    sorted_e = flat_e[order]
    tok = torch.div(order, k, rounding_mode="floor")
    # scatter_add instead of bincount: bincount reads its max back to the host, which a CUDA-graph capture cannot do
    counts = torch.zeros(e, dtype=torch.long, device=flat_e.device).scatter_add_(0, flat_e, torch.ones_like(flat_e))
    start = counts.cumsum(0) - counts
    slot = torch.arange(t * k, device=h.device) - start[sorted_e]                 # position of the assignment inside its expert
    x = h.new_zeros(e, t, h.shape[1])
    x[sorted_e, slot] = h.index_select(0, tok)
    gu = torch.bmm(x, w["ewgu"].to(h.dtype).transpose(1, 2))                      # [E, T, 2f]
    a = ref.swiglu(gu.reshape(e * t, -1)).reshape(e, t, -1)
    y = torch.bmm(a, w["ewd"].to(h.dtype).transpose(1, 2))[sorted_e, slot]          # [T*k, H] in assignment order
    y = y * top_w.reshape(-1)[order].to(h.dtype).unsqueeze(-1)
    y = torch.zeros(t, h.shape[1], dtype=y.dtype, device=h.device).index_add(0, tok, y)
    s = shared_expert(w, h)
    return y if s is None else y + s


@dataclass
class PackedInput:
    """A micro-batch with padding removed: tokens of all sequences concatenated."""
    input_ids: torch.Tensor        # [T] int64
    position_ids: torch.Tensor     # [T] int64
    cu_seqlens: torch.Tensor       # [B+1] int32
    max_seqlen: int
    sp_group: object = None        # Ulysses SP: ids/positions are this rank's slice, cu_seqlens describe the full stream
    total_tokens: int = 0          # un-padded length of the full packed stream (SP only)
    # prefill behind a cached KV prefix (rollout engine's prefix cache): (block idx [Tk], in-page offset [Tk], cu_seqlens_k
    # [B+1], max_k) describing EVERY key position (cached + new) of every sequence in the paged cache
    kv_gather: object = None
    # vision-language inputs (models/vision.py): 3-axis M-RoPE positions [3, T] and the tower's merged patch embeddings that replace
    # the embedding rows at the image placeholder tokens
    mrope_pos: torch.Tensor | None = None
    image_embeds: torch.Tensor | None = None
    # raw tower inputs for engines that own the tower themselves (ZeRO engine: the ``vision`` unit): patches of every image of
    # the micro-batch in sequence order [N, patch_dim] and their (t, h, w) grids
    pixel_values: torch.Tensor | None = None
    grid_thw: object = None

    def rope_inputs(self, cfg: ModelConfig, device):
        """(positions, cos, sin) for the rope op; M-RoPE builds per-token tables once per micro-batch and indexes them by row."""
        cos, sin = rope_tables(cfg, device)
        if self.mrope_pos is None:
            return self.position_ids, cos, sin
        tabs = getattr(self, "_mrope_tabs", None)
        if tabs is None:
            from .vision import mrope_tables

            c, s = mrope_tables(cfg, self.mrope_pos, cos, sin)
            tabs = self._mrope_tabs = (torch.arange(c.shape[0], device=c.device), c, s)
        return tabs


# --------------------------------------------------------------------------------------------
# training / prefill forward
# --------------------------------------------------------------------------------------------
def embed_forward(cfg: ModelConfig, w: dict, inp: PackedInput) -> torch.Tensor:
    x = F.embedding(inp.input_ids, w["embed"])
    if cfg.family == "gpt2":
        x = x + F.embedding(inp.position_ids, w["pos_embed"])
    if inp.image_embeds is not None:
        from .vision import merge_image_embeds

        x = merge_image_embeds(cfg, x, inp.input_ids, inp.image_embeds)
    return x


XIELU_BETA, XIELU_EPS = 0.5, -1e-6


def xielu(x: torch.Tensor, ap: torch.Tensor, an: torch.Tensor) -> torch.Tensor:
    """xIELU with learned scalars (pre-softplus ``ap`` / ``an``), fp32 math: a_p x^2 + b x above zero, a_n (expm1(min(x, eps)) - x) + b x
    below (HF ``XIELUActivation._xielu_python``)."""
    xf = x.float()
    a_p, a_n = F.softplus(ap.float()), XIELU_BETA + F.softplus(an.float())
    y = torch.where(xf > 0, a_p * xf * xf + XIELU_BETA * xf, (torch.expm1(xf.clamp(max=XIELU_EPS)) - xf) * a_n + XIELU_BETA * xf)
    return y.to(x.dtype)


def xielu_mlp(w: dict, h: torch.Tensor, residual: torch.Tensor | None = None) -> torch.Tensor:
    """Apertus MLP: down(xIELU(up(h))) (+ residual)."""
    return ops.linear(xielu(ops.linear(h, w["wup"]), w["act_ap"], w["act_an"]), w["wd"], residual=residual)


def _mlp_forward(cfg: ModelConfig, w: dict, x: torch.Tensor, aux_out: list | None = None) -> torch.Tensor:
    """Second half of a llama-family block without adapters: x + MLP(norm(x)), dense SwiGLU or mixture of experts."""
    h = ops.rmsnorm(x, w["ln2"], cfg.rms_norm_eps)
    if "router" in w:
        return x + moe_mlp(cfg, w, h, aux_out)
    if "wup" in w:
        return xielu_mlp(w, h, x)
    return ops.linear(ops.linear_swiglu(h, w["wgu"]), w["wd"], residual=x)


def layer_forward(cfg: ModelConfig, w: dict, x: torch.Tensor, inp: PackedInput, kv_sink=None, aux_out: list | None = None) -> torch.Tensor:
    """One decoder block on packed activations x:[T,H].  ``kv_sink`` = (k_cache, v_cache, slot_mapping)
    makes this a prefill step that also fills the paged KV cache (rollout engine)."""
    nq, nkv, d = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    t = x.shape[0]
    if cfg.sliding_window and inp.max_seqlen > cfg.sliding_window:
        raise NotImplementedError(f"sequence of {inp.max_seqlen} tokens exceeds the model's attention window "
                                  f"({cfg.sliding_window}); windowed attention is not implemented, only the full-causal case")
    if cfg.family == "llama":
        lora = "_lora_scale" in w
        if lora:
            from . import lora as L
        h = ops.rmsnorm(x, w["ln1"], cfg.rms_norm_eps)
        if cfg.attn_type == "mla":
            from . import mla

            if lora:
                raise NotImplementedError("LoRA adapters are not implemented for MLA attention")
            x = ops.linear(mla.attention_forward(cfg, w, h, inp, kv_sink), w["wo"], residual=x)
            return x if w.get("_skip_mlp") else _mlp_forward(cfg, w, x, aux_out)
        qkv = ops.linear(h, w["wqkv"], bias=w.get("bqkv"))
        if lora and L.has_lora(w, "q_proj", "k_proj", "v_proj"):
            qkv = qkv + L.qkv_delta(cfg, w, h)
        qkv = qk_norm_(cfg, w, qkv)
        pos, cos, sin = inp.rope_inputs(cfg, x.device)
        if kv_sink is not None:
            qkv = rope_and_cache_(qkv, pos, cos, sin, nq, nkv, d, *kv_sink)
        else:
            qkv = ops.apply_rope(qkv, pos, cos, sin, nq, nkv, d)
        q = qkv[:, : nq * d].view(t, nq, d)
        k = qkv[:, nq * d:(nq + nkv) * d].view(t, nkv, d)
        v = qkv[:, (nq + nkv) * d:].view(t, nkv, d)
        if inp.kv_gather is not None and kv_sink is not None:
            from ..ops.attention import attention_varlen_prefix

            blk, off, cu_k, max_k = inp.kv_gather            # new tokens were appended to the cache just above
            a = attention_varlen_prefix(q, kv_sink[0][blk, :, off, :], kv_sink[1][blk, :, off, :], inp.cu_seqlens, cu_k,
                                        inp.max_seqlen, max_k)
        elif inp.sp_group is not None:
            from ..parallel.ulysses import ulysses_attention

            a = ulysses_attention(q, k, v, inp.cu_seqlens, inp.max_seqlen, inp.sp_group, inp.total_tokens, causal=True)
        else:
            a = attention_varlen(q, k, v, inp.cu_seqlens, inp.max_seqlen, causal=True)
        a = a.reshape(t, nq * d)
        if cfg.sandwich_norm:
            o = ops.linear(a, w["wo"])
            if lora and L.has_lora(w, "o_proj"):
                o = o + L.single_delta(w, a, "o_proj")
            x = x + ops.rmsnorm(o, w["ln_attn"], cfg.rms_norm_eps)
        else:
            x = ops.linear(a, w["wo"], residual=x)
            if lora and L.has_lora(w, "o_proj"):
                x = x + L.single_delta(w, a, "o_proj")
        if w.get("_skip_mlp"):              # callers that supply their own MLP (expert parallelism, parallel/expert.py)
            return x
        h = ops.rmsnorm(x, w["ln2"], cfg.rms_norm_eps)
        if "router" in w:
            y = moe_mlp(cfg, w, h, aux_out)
            return x + (ops.rmsnorm(y, w["ln_mlp"], cfg.rms_norm_eps) if cfg.sandwich_norm else y)
        if "wup" in w:
            if lora:
                raise NotImplementedError("LoRA adapters are not implemented for the xIELU MLP")
            return xielu_mlp(w, h, x)
        if lora and L.has_lora(w, "gate_proj", "up_proj"):
            a = ref.swiglu(ops.linear(h, w["wgu"]) + L.gate_up_delta(cfg, w, h))    # adapters act before the activation
        else:
            a = ops.linear_swiglu(h, w["wgu"])
        if cfg.sandwich_norm:
            y = ops.linear(a, w["wd"])
            if lora and L.has_lora(w, "down_proj"):
                y = y + L.single_delta(w, a, "down_proj")
            return x + ops.rmsnorm(y, w["ln_mlp"], cfg.rms_norm_eps)
        y = ops.linear(a, w["wd"], residual=x)
        if lora and L.has_lora(w, "down_proj"):
            y = y + L.single_delta(w, a, "down_proj")
        return y
    # ---- gpt2 block (library ops; plumbing / CPU config)
    h = F.layer_norm(x, (cfg.hidden_size,), w["ln1_w"], w["ln1_b"], cfg.rms_norm_eps)
    qkv = F.linear(h, w["wqkv"], w["bqkv"])
    q = qkv[:, : nq * d].reshape(t, nq, d)
    k = qkv[:, nq * d:2 * nq * d].reshape(t, nq, d)
    v = qkv[:, 2 * nq * d:].reshape(t, nq, d)
    if kv_sink is not None:
        k_cache, v_cache, slots = kv_sink[0], kv_sink[1], kv_sink[2]
        page = k_cache.shape[2]
        blk, off = torch.div(slots, page, rounding_mode="floor"), slots % page
        k_cache[blk, :, off] = k.to(k_cache.dtype)
        v_cache[blk, :, off] = v.to(v_cache.dtype)
    a = attention_varlen(q, k, v, inp.cu_seqlens, inp.max_seqlen, causal=True)
    x = x + F.linear(a.reshape(t, nq * d), w["wo"], w["bo"])
    h = F.layer_norm(x, (cfg.hidden_size,), w["ln2_w"], w["ln2_b"], cfg.rms_norm_eps)
    return x + F.linear(F.gelu(F.linear(h, w["wfc"], w["bfc"]), approximate="tanh"), w["wproj"], w["bproj"])


def final_norm(cfg: ModelConfig, w: dict, x: torch.Tensor) -> torch.Tensor:
    if cfg.family == "llama":
        return ops.rmsnorm(x, w["norm_w"], cfg.rms_norm_eps)
    return F.layer_norm(x, (cfg.hidden_size,), w["norm_w"], w["norm_b"], cfg.rms_norm_eps)


def lm_head_weight(cfg: ModelConfig, head_w: dict, embed_w: dict) -> torch.Tensor:
    return embed_w["embed"] if cfg.tie_word_embeddings else head_w["lm_head"]


def logits_forward(cfg: ModelConfig, head_w: dict, embed_w: dict, hidden: torch.Tensor) -> torch.Tensor:
    return ops.linear(hidden, lm_head_weight(cfg, head_w, embed_w).to(hidden.dtype))


def value_forward(cfg: ModelConfig, head_w: dict, hidden: torch.Tensor) -> torch.Tensor:
    v = F.linear(hidden, head_w["v_head"][:1].to(hidden.dtype)).squeeze(-1).float()
    return v + head_w["v_head_b"][0].float() if "v_head_b" in head_w else v


# --------------------------------------------------------------------------------------------
# decode forward (one new token per running sequence, paged KV)
# --------------------------------------------------------------------------------------------
@dataclass
class DecodeInput:
    input_ids: torch.Tensor       # [B]
    position_ids: torch.Tensor    # [B] int64 (index of the new token)
    slot_mapping: torch.Tensor    # [B] int64 flat slot (block*page + offset) for the new token
    block_table: torch.Tensor     # [B, max_blocks] int32
    seq_lens: torch.Tensor        # [B] int32 (length INCLUDING the new token)
    max_len: int = 0


def layer_decode(cfg: ModelConfig, w: dict, x: torch.Tensor, inp: DecodeInput, k_cache: torch.Tensor,
                 v_cache: torch.Tensor) -> torch.Tensor:
    nq, nkv, d = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    b = x.shape[0]
    scale = d ** -0.5
    if cfg.family == "llama" and cfg.attn_type == "mla":
        from . import mla

        h = ops.rmsnorm(x, w["ln1"], cfg.rms_norm_eps)
        x = ops.linear(mla.attention_decode(cfg, w, h, inp, k_cache), w["wo"], residual=x)
        h = ops.rmsnorm(x, w["ln2"], cfg.rms_norm_eps)
        if "router" in w:
            return x + (moe_mlp_static(cfg, w, h) if _MOE_STATIC_DECODE() else moe_mlp(cfg, w, h))
        if "wup" in w:
            return xielu_mlp(w, h, x)
        return ops.linear(ops.linear_swiglu(h, w["wgu"]), w["wd"], residual=x)
    if cfg.family == "llama":
        h = ops.rmsnorm(x, w["ln1"], cfg.rms_norm_eps)
        qkv = qk_norm_(cfg, w, ops.linear(h, w["wqkv"], bias=w.get("bqkv")))
        cos, sin = rope_tables(cfg, x.device)
        rope_and_cache_(qkv, inp.position_ids, cos, sin, nq, nkv, d, k_cache, v_cache, inp.slot_mapping,
                        k_cache.shape[2])
        a = ops.paged_decode_attention(qkv[:, : nq * d], k_cache, v_cache, inp.block_table, inp.seq_lens, nq, nkv, d,
                                       scale, max_len=inp.max_len)
        if cfg.sandwich_norm:
            x = x + ops.rmsnorm(ops.linear(a, w["wo"]), w["ln_attn"], cfg.rms_norm_eps)
        else:
            x = ops.linear(a, w["wo"], residual=x)
        h = ops.rmsnorm(x, w["ln2"], cfg.rms_norm_eps)
        if "router" in w:
            y = moe_mlp_static(cfg, w, h) if _MOE_STATIC_DECODE() else moe_mlp(cfg, w, h)
            return x + (ops.rmsnorm(y, w["ln_mlp"], cfg.rms_norm_eps) if cfg.sandwich_norm else y)
        if "wup" in w:
            return xielu_mlp(w, h, x)
        a = ops.linear_swiglu(h, w["wgu"])
        if cfg.sandwich_norm:
            return x + ops.rmsnorm(ops.linear(a, w["wd"]), w["ln_mlp"], cfg.rms_norm_eps)
        return ops.linear(a, w["wd"], residual=x)
    h = F.layer_norm(x, (cfg.hidden_size,), w["ln1_w"], w["ln1_b"], cfg.rms_norm_eps)
    qkv = F.linear(h, w["wqkv"], w["bqkv"])
    page = k_cache.shape[2]
    blk, off = torch.div(inp.slot_mapping, page, rounding_mode="floor"), inp.slot_mapping % page
    k_cache[blk, :, off] = qkv[:, nq * d:2 * nq * d].reshape(b, nq, d).to(k_cache.dtype)
    v_cache[blk, :, off] = qkv[:, 2 * nq * d:].reshape(b, nq, d).to(v_cache.dtype)
    a = ops.paged_decode_attention(qkv[:, : nq * d].contiguous(), k_cache, v_cache, inp.block_table, inp.seq_lens, nq,
                                   nq, d, scale, max_len=inp.max_len)
    x = x + F.linear(a, w["wo"], w["bo"])
    h = F.layer_norm(x, (cfg.hidden_size,), w["ln2_w"], w["ln2_b"], cfg.rms_norm_eps)
    return x + F.linear(F.gelu(F.linear(h, w["wfc"], w["bfc"]), approximate="tanh"), w["wproj"], w["bproj"])


def layer_decode_fused(cfg: ModelConfig, w: dict, x: torch.Tensor, delta: torch.Tensor | None, inp: DecodeInput,
                       k_cache: torch.Tensor, v_cache: torch.Tensor):
    """Decode block with the residual adds folded into the following RMSNorm (one kernel: x += delta; h = norm(x)).
    ``x`` is the residual stream (updated in place), ``delta`` the not-yet-added output of the previous block.
    Returns (x, new_delta).  Llama family, CUDA bf16."""
    from ..ops import ext

    C = ext()
    nq, nkv, d = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    h = ops.rmsnorm(x, w["ln1"], cfg.rms_norm_eps) if delta is None else C.add_rmsnorm_(x, delta, w["ln1"], cfg.rms_norm_eps)
    qkv = qk_norm_(cfg, w, ops.linear(h, w["wqkv"], bias=w.get("bqkv")))
    cos, sin = rope_tables(cfg, x.device)
    rope_and_cache_(qkv, inp.position_ids, cos, sin, nq, nkv, d, k_cache, v_cache, inp.slot_mapping, k_cache.shape[2])
    a = ops.paged_decode_attention(qkv[:, : nq * d], k_cache, v_cache, inp.block_table, inp.seq_lens, nq, nkv, d, d ** -0.5,
                                   max_len=inp.max_len)
    o = ops.linear(a, w["wo"])
    h = C.add_rmsnorm_(x, o, w["ln2"], cfg.rms_norm_eps)
    g = ops.linear_swiglu(h, w["wgu"])
    return x, ops.linear(g, w["wd"])


# --------------------------------------------------------------------------------------------
# a plain (unsharded) weight container — single-process inference, tests, HF import/export
# --------------------------------------------------------------------------------------------
class TransformerWeights:
    def __init__(self, cfg: ModelConfig, device="cpu", dtype=torch.float32, init: bool = True, seed: int = 0):
        self.cfg = cfg
        g = torch.Generator(device=device)
        g.manual_seed(seed)

        def make(specs):
            out = {}
            for n, shp in specs:
                t = torch.empty(shp, device=device, dtype=dtype)
                if init:
                    init_param_(n, t, cfg, g)
                out[n] = t
            return out

        self.embed = make(embed_param_specs(cfg))
        self.layers = [make(layer_param_specs(cfg, i)) for i in range(cfg.num_layers)]
        self.head = make(head_param_specs(cfg))

    def named_tensors(self):
        for n, t in self.embed.items():
            yield n, t
        for i, l in enumerate(self.layers):
            for n, t in l.items():
                yield f"layers.{i}.{n}", t
        for n, t in self.head.items():
            yield n, t

    def state_dict(self) -> dict[str, torch.Tensor]:
        return dict(self.named_tensors())

    def load_state_dict(self, sd: dict[str, torch.Tensor]) -> None:
        for n, t in self.named_tensors():
            t.copy_(sd[n])

    def requires_grad_(self, flag: bool = True):
        for _, t in self.named_tensors():
            t.requires_grad_(flag)
        return self


def forward_hidden(cfg: ModelConfig, weights: TransformerWeights, inp: PackedInput) -> torch.Tensor:
    x = embed_forward(cfg, weights.embed, inp)
    for i, w in enumerate(weights.layers):
        x = layer_forward(cfg, w, x, inp)
        if inp.image_embeds is not None and cfg.vision is not None:
            from .vision import deepstack_add

            x = deepstack_add(cfg, x, inp.input_ids, inp.image_embeds, i)          # Qwen3-VL DeepStack (no-op otherwise)
    return final_norm(cfg, weights.head, x)


def pack_padded(input_ids: torch.Tensor, attention_mask: torch.Tensor, position_ids: torch.Tensor | None = None):
    """[B,S] padded batch -> PackedInput + flat indices of the kept tokens (for re-padding)."""
    mask = attention_mask.bool()
    lens = mask.sum(-1).to(torch.int32)
    idx = mask.reshape(-1).nonzero(as_tuple=True)[0]
    if position_ids is None:
        position_ids = (mask.long().cumsum(-1) - 1).clamp_min(0)
    cu = torch.zeros(lens.numel() + 1, dtype=torch.int32, device=input_ids.device)
    cu[1:] = lens.cumsum(0)
    mrope = None
    if position_ids.dim() == 3:               # M-RoPE [B, 3, S] (vision-language batches): keep all three axes, packed [3, T]
        mrope = position_ids.permute(1, 0, 2).reshape(3, -1)[:, idx].long().contiguous()
        position_ids = position_ids[:, 0]
    inp = PackedInput(input_ids.reshape(-1)[idx], position_ids.reshape(-1)[idx].long(), cu,
                      int(lens.max().item()) if lens.numel() else 0, mrope_pos=mrope)
    return inp, idx


def unpack_to_padded(values: torch.Tensor, idx: torch.Tensor, batch: int, seqlen: int, fill: float = 0.0):
    out = values.new_full((batch * seqlen, *values.shape[1:]), fill)
    out = out.index_copy(0, idx, values)
    return out.view(batch, seqlen, *values.shape[1:])


# --------------------------------------------------------------------------------------------
# HF checkpoint name mapping (import / export)
# --------------------------------------------------------------------------------------------
def value_head_from_hf(cfg: ModelConfig, score_w: torch.Tensor | None, score_b: torch.Tensor | None) -> dict[str, torch.Tensor]:
    """``score.weight`` [1, H] (+ ``score.bias`` [1]) of a token-classification / reward-model checkpoint -> our padded value head.
    Empty when the checkpoint has no scoring head (a critic started from a language model keeps its fresh head)."""
    if score_w is None:
        return {}
    if score_w.shape[0] != 1:
        raise ValueError(f"value / reward heads are scalar: score.weight has {score_w.shape[0]} labels")
    w = score_w.new_zeros(8, cfg.hidden_size)
    w[0] = score_w[0]
    b = score_w.new_zeros(8)
    if score_b is not None:
        b[0] = score_b[0]
    return {"v_head": w, "v_head_b": b}


def _convert_mlp(cfg: ModelConfig, hf_sd: dict, out: dict, i: int, p: str) -> None:
    """MLP tensors of block ``i`` (HF prefix ``p``) into ``out``: dense SwiGLU (separate or fused gate_up) or router + experts
    (stacked, one module per expert, or Mixtral's block_sparse_moe naming) + shared experts."""
    if is_sparse_layer(cfg, i):
        fm = cfg.moe_intermediate_size
        if p + "mlp.gate.e_score_correction_bias" in hf_sd:
            out[f"layers.{i}.router_bias"] = hf_sd[p + "mlp.gate.e_score_correction_bias"]
        for sh in ("mlp.shared_expert.", "mlp.shared_experts."):          # Qwen2-MoE / DeepSeek-style naming
            if p + sh + "gate_proj.weight" in hf_sd:
                out[f"layers.{i}.swgu"] = ref.interleave_gate_up(hf_sd[p + sh + "gate_proj.weight"], hf_sd[p + sh + "up_proj.weight"])
                out[f"layers.{i}.swd"] = hf_sd[p + sh + "down_proj.weight"]
        if p + "mlp.shared_expert_gate.weight" in hf_sd:
            out[f"layers.{i}.sgate"] = hf_sd[p + "mlp.shared_expert_gate.weight"]
        if p + "block_sparse_moe.gate.weight" in hf_sd:      # Mixtral checkpoints on disk: block_sparse_moe.experts.{e}.w1 / w3 / w2
            q = p + "block_sparse_moe."
            out[f"layers.{i}.router"] = hf_sd[q + "gate.weight"]
            pairs = [(hf_sd[f"{q}experts.{e}.w1.weight"], hf_sd[f"{q}experts.{e}.w3.weight"]) for e in range(cfg.num_experts)]
            out[f"layers.{i}.ewd"] = torch.stack([hf_sd[f"{q}experts.{e}.w2.weight"] for e in range(cfg.num_experts)])
            out[f"layers.{i}.ewgu"] = torch.stack([ref.interleave_gate_up(g, u) for g, u in pairs])
            return
        out[f"layers.{i}.router"] = hf_sd[p + "mlp.gate.weight"]
        if p + "mlp.experts.gate_up_proj" in hf_sd:          # stacked layout [E, 2f, h] (gate rows then up rows)
            gu, dn = hf_sd[p + "mlp.experts.gate_up_proj"], hf_sd[p + "mlp.experts.down_proj"]
            pairs = [(gu[e, :fm], gu[e, fm:]) for e in range(cfg.num_experts)]
            out[f"layers.{i}.ewd"] = dn.clone()
        else:                                                # one module per expert
            pairs = [(hf_sd[f"{p}mlp.experts.{e}.gate_proj.weight"], hf_sd[f"{p}mlp.experts.{e}.up_proj.weight"]) for e in range(cfg.num_experts)]
            out[f"layers.{i}.ewd"] = torch.stack([hf_sd[f"{p}mlp.experts.{e}.down_proj.weight"] for e in range(cfg.num_experts)])
        out[f"layers.{i}.ewgu"] = torch.stack([ref.interleave_gate_up(g, u) for g, u in pairs])
        return
    if cfg.mlp_act == "xielu":                               # Apertus: no gate; the activation's two learned scalars live under act_fn
        out[f"layers.{i}.wup"], out[f"layers.{i}.wd"] = hf_sd[p + "mlp.up_proj.weight"], hf_sd[p + "mlp.down_proj.weight"]
        out[f"layers.{i}.act_ap"] = hf_sd[p + "mlp.act_fn.alpha_p"].reshape(1)
        out[f"layers.{i}.act_an"] = hf_sd[p + "mlp.act_fn.alpha_n"].reshape(1)
        return
    if p + "mlp.gate_up_proj.weight" in hf_sd:               # GLM: one fused projection, gate rows then up rows
        gu = hf_sd[p + "mlp.gate_up_proj.weight"]
        out[f"layers.{i}.wgu"] = ref.interleave_gate_up(gu[: gu.shape[0] // 2], gu[gu.shape[0] // 2:])
    else:
        out[f"layers.{i}.wgu"] = ref.interleave_gate_up(hf_sd[p + "mlp.gate_proj.weight"], hf_sd[p + "mlp.up_proj.weight"])
    out[f"layers.{i}.wd"] = hf_sd[p + "mlp.down_proj.weight"]


def convert_hf_state_dict(cfg: ModelConfig, hf_sd: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """Map HF Llama / Qwen2 / Qwen3 / GPT-2 parameter names to ours (fusing q/k/v and interleaving gate/up; GPT-2's Conv1D
    weights are stored [in, out] and get transposed)."""
    if cfg.family == "gpt2":
        g = {k.removeprefix("transformer."): v for k, v in hf_sd.items()}
        out = {"embed": g["wte.weight"], "pos_embed": g["wpe.weight"], "norm_w": g["ln_f.weight"], "norm_b": g["ln_f.bias"]}
        for i in range(cfg.num_layers):
            p, q = f"h.{i}.", f"layers.{i}."
            out[q + "ln1_w"], out[q + "ln1_b"] = g[p + "ln_1.weight"], g[p + "ln_1.bias"]
            out[q + "ln2_w"], out[q + "ln2_b"] = g[p + "ln_2.weight"], g[p + "ln_2.bias"]
            out[q + "wqkv"], out[q + "bqkv"] = g[p + "attn.c_attn.weight"].t().contiguous(), g[p + "attn.c_attn.bias"]
            out[q + "wo"], out[q + "bo"] = g[p + "attn.c_proj.weight"].t().contiguous(), g[p + "attn.c_proj.bias"]
            out[q + "wfc"], out[q + "bfc"] = g[p + "mlp.c_fc.weight"].t().contiguous(), g[p + "mlp.c_fc.bias"]
            out[q + "wproj"], out[q + "bproj"] = g[p + "mlp.c_proj.weight"].t().contiguous(), g[p + "mlp.c_proj.bias"]
        return out
    if any(k.startswith("model.language_model.") for k in hf_sd):        # VL checkpoints nest the text stack
        hf_sd = {k.replace("model.language_model.", "model.", 1): v for k, v in hf_sd.items()}
    out = {"embed": hf_sd["model.embed_tokens.weight"], "norm_w": hf_sd["model.norm.weight"]}
    if cfg.value_head:
        out.update(value_head_from_hf(cfg, hf_sd.get("score.weight"), hf_sd.get("score.bias")))
    elif not cfg.tie_word_embeddings:
        out["lm_head"] = hf_sd["lm_head.weight"]
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        apertus = p + "attention_layernorm.weight" in hf_sd
        out[f"layers.{i}.ln1"] = hf_sd[p + ("attention_layernorm.weight" if apertus else "input_layernorm.weight")]
        out[f"layers.{i}.ln2"] = hf_sd[p + ("feedforward_layernorm.weight" if apertus else "post_attention_layernorm.weight")]
        if cfg.attn_type == "mla":
            from . import mla

            out.update({f"layers.{i}.{k}": v for k, v in mla.from_hf(cfg, hf_sd, p).items()})
            _convert_mlp(cfg, hf_sd, out, i, p)
            continue
        perm = rope_feature_perm(cfg)
        fix = (lambda t, n: _permute_heads(t, perm, cfg.head_dim) if (perm is not None and n in "qk") else t)
        out[f"layers.{i}.wqkv"] = torch.cat([fix(hf_sd[p + f"self_attn.{n}_proj.weight"], n) for n in "qkv"], dim=0)
        if cfg.qkv_bias:
            out[f"layers.{i}.bqkv"] = torch.cat([fix(hf_sd[p + f"self_attn.{n}_proj.bias"], n) for n in "qkv"], dim=0)
        if cfg.qk_norm:
            qn, kn = hf_sd[p + "self_attn.q_norm.weight"], hf_sd[p + "self_attn.k_norm.weight"]
            out[f"layers.{i}.q_norm"], out[f"layers.{i}.k_norm"] = (qn, kn) if perm is None else (qn[perm], kn[perm])
        if cfg.sandwich_norm:
            out[f"layers.{i}.ln_attn"], out[f"layers.{i}.ln_mlp"] = hf_sd[p + "post_self_attn_layernorm.weight"], hf_sd[p + "post_mlp_layernorm.weight"]
        out[f"layers.{i}.wo"] = hf_sd[p + "self_attn.o_proj.weight"]
        _convert_mlp(cfg, hf_sd, out, i, p)
    return out


def _export_mlp(cfg: ModelConfig, sd: dict, out: dict, i: int, p: str, fused_gu: bool) -> None:
    """MLP tensors and the two block norms of block ``i`` in HF names (inverse of ``_convert_mlp``)."""
    if is_sparse_layer(cfg, i):                              # stacked HF layout
        gus = [ref.deinterleave_gate_up(sd[f"layers.{i}.ewgu"][e]) for e in range(cfg.num_experts)]
        out[p + "mlp.gate.weight"] = sd[f"layers.{i}.router"]
        out[p + "mlp.experts.gate_up_proj"] = torch.stack([torch.cat([g, u], dim=0) for g, u in gus])
        out[p + "mlp.experts.down_proj"] = sd[f"layers.{i}.ewd"]
        if f"layers.{i}.router_bias" in sd:
            out[p + "mlp.gate.e_score_correction_bias"] = sd[f"layers.{i}.router_bias"]
        if f"layers.{i}.swgu" in sd:
            sh = "mlp.shared_expert." if cfg.shared_expert_gate else "mlp.shared_experts."
            g, u = ref.deinterleave_gate_up(sd[f"layers.{i}.swgu"])
            out[p + sh + "gate_proj.weight"], out[p + sh + "up_proj.weight"], out[p + sh + "down_proj.weight"] = g, u, sd[f"layers.{i}.swd"]
        if f"layers.{i}.sgate" in sd:
            out[p + "mlp.shared_expert_gate.weight"] = sd[f"layers.{i}.sgate"]
    elif cfg.mlp_act == "xielu":
        out[p + "mlp.up_proj.weight"], out[p + "mlp.down_proj.weight"] = sd[f"layers.{i}.wup"], sd[f"layers.{i}.wd"]
        out[p + "mlp.act_fn.alpha_p"], out[p + "mlp.act_fn.alpha_n"] = sd[f"layers.{i}.act_ap"].reshape(1), sd[f"layers.{i}.act_an"].reshape(1)
        out[p + "mlp.act_fn.beta"] = sd[f"layers.{i}.act_ap"].new_tensor(XIELU_BETA)
        out[p + "mlp.act_fn.eps"] = sd[f"layers.{i}.act_ap"].new_tensor(XIELU_EPS)
        out[p + "attention_layernorm.weight"], out[p + "feedforward_layernorm.weight"] = sd[f"layers.{i}.ln1"], sd[f"layers.{i}.ln2"]
        return
    else:
        g, u = ref.deinterleave_gate_up(sd[f"layers.{i}.wgu"])
        if fused_gu:
            out[p + "mlp.gate_up_proj.weight"] = torch.cat([g, u], dim=0)
        else:
            out[p + "mlp.gate_proj.weight"], out[p + "mlp.up_proj.weight"] = g, u
        out[p + "mlp.down_proj.weight"] = sd[f"layers.{i}.wd"]
    out[p + "input_layernorm.weight"] = sd[f"layers.{i}.ln1"]
    out[p + "post_attention_layernorm.weight"] = sd[f"layers.{i}.ln2"]


def export_hf_state_dict(cfg: ModelConfig, sd: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    if cfg.family == "gpt2":
        out = {"transformer.wte.weight": sd["embed"], "transformer.wpe.weight": sd["pos_embed"], "transformer.ln_f.weight": sd["norm_w"],
               "transformer.ln_f.bias": sd["norm_b"], "lm_head.weight": sd["embed"]}
        for i in range(cfg.num_layers):
            p, q = f"transformer.h.{i}.", f"layers.{i}."
            out[p + "ln_1.weight"], out[p + "ln_1.bias"] = sd[q + "ln1_w"], sd[q + "ln1_b"]
            out[p + "ln_2.weight"], out[p + "ln_2.bias"] = sd[q + "ln2_w"], sd[q + "ln2_b"]
            out[p + "attn.c_attn.weight"], out[p + "attn.c_attn.bias"] = sd[q + "wqkv"].t().contiguous(), sd[q + "bqkv"]
            out[p + "attn.c_proj.weight"], out[p + "attn.c_proj.bias"] = sd[q + "wo"].t().contiguous(), sd[q + "bo"]
            out[p + "mlp.c_fc.weight"], out[p + "mlp.c_fc.bias"] = sd[q + "wfc"].t().contiguous(), sd[q + "bfc"]
            out[p + "mlp.c_proj.weight"], out[p + "mlp.c_proj.bias"] = sd[q + "wproj"].t().contiguous(), sd[q + "bproj"]
        return out
    out = {"model.embed_tokens.weight": sd["embed"], "model.norm.weight": sd["norm_w"]}
    if cfg.value_head:                                           # critics / reward models: HF ...ForTokenClassification, num_labels = 1
        out["score.weight"] = sd["v_head"][:1]
        out["score.bias"] = sd["v_head_b"][:1] if "v_head_b" in sd else sd["v_head"].new_zeros(1)
    else:
        out["lm_head.weight"] = sd["embed"] if cfg.tie_word_embeddings else sd["lm_head"]
    q, kv = cfg.q_size, cfg.kv_size
    perm = rope_feature_perm(cfg)
    back = (lambda t: t) if perm is None else (lambda t: _permute_heads(t, perm, cfg.head_dim, inverse=True))
    fused_gu = getattr(cfg, "hf_model_type", "") in ("glm", "glm4", "glm4v")
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        if cfg.attn_type == "mla":
            from . import mla

            out.update(mla.to_hf(cfg, sd, f"layers.{i}.", p))
            _export_mlp(cfg, sd, out, i, p, fused_gu)
            continue
        w = sd[f"layers.{i}.wqkv"]
        out[p + "self_attn.q_proj.weight"], out[p + "self_attn.k_proj.weight"], out[p + "self_attn.v_proj.weight"] = (
            back(w[:q]), back(w[q:q + kv]), w[q + kv:])
        if cfg.qkv_bias:
            b = sd[f"layers.{i}.bqkv"]
            out[p + "self_attn.q_proj.bias"], out[p + "self_attn.k_proj.bias"], out[p + "self_attn.v_proj.bias"] = (
                back(b[:q]), back(b[q:q + kv]), b[q + kv:])
        if cfg.qk_norm:
            out[p + "self_attn.q_norm.weight"], out[p + "self_attn.k_norm.weight"] = back(sd[f"layers.{i}.q_norm"]), back(sd[f"layers.{i}.k_norm"])
        if cfg.sandwich_norm:
            out[p + "post_self_attn_layernorm.weight"], out[p + "post_mlp_layernorm.weight"] = sd[f"layers.{i}.ln_attn"], sd[f"layers.{i}.ln_mlp"]
        out[p + "self_attn.o_proj.weight"] = sd[f"layers.{i}.wo"]
        _export_mlp(cfg, sd, out, i, p, fused_gu)
    return out
