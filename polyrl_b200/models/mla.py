"""Multi-head latent attention (MLA) — the attention of the DeepSeek-V2/V3 architecture (DeepSeek-V3 / R1, Moonlight, the text
stack of Kimi-VL); the reference reaches these models through verl's mcore registry (3rdparty/verl/verl/models/mcore/registry.py,
``DeepseekV3ForCausalLM``) and through HF ``AutoModel`` in the FSDP workers.

Layout chosen here (B200-first, not a translation of the HF module):

* ONE first-stage GEMM per block, ``wa`` = [query bottleneck (or the full query projection) | KV latent | shared rotary key
  features], instead of HF's separate ``q_a_proj`` / ``kv_a_proj_with_mqa``: the block input is read once.
* Rotary features are de-interleaved at load time (``rope_interleave`` checkpoints pair (2j, 2j+1)); q . k is invariant under a
  common permutation of the rotary features, so run time always uses the half-split rotation.
* Training / prefill ("expanded" form): latent -> per-head keys / values through ``wkv_up``, packed varlen attention with
  qk width ``head_dim`` = nope + rope; values are zero-padded to that width for the flash kernels and the pad is dropped after.
* Rollout ("absorbed" form): the paged cache stores ONLY the normalised latent and the rotated shared key features —
  ``kv_lora_rank + qk_rope_head_dim`` numbers per token and layer (576 for DeepSeek-V3 against 2 x 128 x 192 expanded: 85x
  smaller, which is what lets long rollouts of a 128-head model live in 180 GB next to the weights).  Decode folds the key
  up-projection into the query and the value up-projection into the output, so a step is multi-QUERY attention over the latent
  pool: scores = (q_nope W_uk) . c + q_rot . k_rot,  out = (softmax . c) W_uv.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .. import ops
from ..ops.attention import attention_varlen, attention_varlen_prefix
from .config import ModelConfig

_TABLES: dict = {}


def param_specs(cfg: ModelConfig) -> list[tuple[str, tuple[int, ...]]]:
    """Attention parameters of one MLA block (between ``ln1`` and ``ln2``)."""
    h, nh, r = cfg.hidden_size, cfg.num_heads, cfg.kv_lora_rank
    qk = nh * cfg.head_dim
    first = (cfg.q_lora_rank or qk) + r + cfg.qk_rope_head_dim
    specs = [("wa", (first, h))]
    if cfg.q_lora_rank:
        specs += [("q_a_ln", (cfg.q_lora_rank,)), ("wq_up", (qk, cfg.q_lora_rank))]
    specs += [("kv_a_ln", (r,)), ("wkv_up", (nh * (cfg.qk_nope_head_dim + cfg.v_head_dim), r)), ("wo", (h, cfg.o_in_size))]
    return specs


def softmax_scale(cfg: ModelConfig) -> float:
    """1 / sqrt(qk width), times YaRN's squared ``mscale_all_dim`` temperature when the checkpoint declares one (HF
    DeepseekV3Attention.__init__)."""
    s = cfg.head_dim ** -0.5
    rs = cfg.rope_scaling or {}
    if rs.get("rope_type", "default") != "default" and rs.get("mscale_all_dim"):
        f = float(rs.get("factor", 1.0))
        m = 1.0 if f <= 1.0 else 0.1 * float(rs["mscale_all_dim"]) * math.log(f) + 1.0
        s *= m * m
    return s


def rope_tables(cfg: ModelConfig, device) -> tuple[torch.Tensor, torch.Tensor]:
    """cos / sin [max_position, qk_rope_head_dim / 2] (fp32), long-context scaling applied."""
    from .transformer import scaled_inv_freq

    key = (cfg.qk_rope_head_dim, cfg.rope_theta, cfg.max_position_embeddings, str(device),
           tuple(sorted((k, str(v)) for k, v in (cfg.rope_scaling or {}).items())))
    if key not in _TABLES:
        inv, af = scaled_inv_freq(cfg)
        fr = torch.outer(torch.arange(cfg.max_position_embeddings, dtype=torch.float32), inv)
        _TABLES[key] = ((fr.cos() * af).to(device).contiguous(), (fr.sin() * af).to(device).contiguous())
    return _TABLES[key]


def _rotate(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """Half-split rotation of x:[T, n, r] with per-token tables cos / sin:[T, r/2] (fp32 math, input dtype out)."""
    x1, x2 = x.float().chunk(2, dim=-1)
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).to(x.dtype)


def project(cfg: ModelConfig, w: dict, h: torch.Tensor, pos: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Normed block input h:[T,H] -> (q [T, nh, nope + rope] with its rotary part rotated, cache entry [T, kv_lora_rank + rope] =
    RMS-normalised latent | rotated shared key features)."""
    t = h.shape[0]
    nh, dn, dr, r = cfg.num_heads, cfg.qk_nope_head_dim, cfg.qk_rope_head_dim, cfg.kv_lora_rank
    a = ops.linear(h, w["wa"])
    qw = cfg.q_lora_rank or nh * (dn + dr)
    q = a[:, :qw]
    if cfg.q_lora_rank:
        q = ops.linear(ops.rmsnorm(q.contiguous(), w["q_a_ln"], cfg.rms_norm_eps), w["wq_up"])
    q = q.reshape(t, nh, dn + dr)
    cos, sin = rope_tables(cfg, h.device)
    cos, sin = cos.index_select(0, pos), sin.index_select(0, pos)
    q = torch.cat([q[..., :dn], _rotate(q[..., dn:], cos, sin)], dim=-1)
    c = ops.rmsnorm(a[:, qw:qw + r].contiguous(), w["kv_a_ln"], cfg.rms_norm_eps)
    k_rot = _rotate(a[:, qw + r:].reshape(t, 1, dr), cos, sin).reshape(t, dr)
    return q, torch.cat([c, k_rot], dim=-1)


def expand_kv(cfg: ModelConfig, w: dict, entry: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Cache entries [T, r + rope] -> per-head keys [T, nh, nope + rope] and values [T, nh, v]."""
    t = entry.shape[0]
    nh, dn, dr, dv, r = cfg.num_heads, cfg.qk_nope_head_dim, cfg.qk_rope_head_dim, cfg.v_head_dim, cfg.kv_lora_rank
    kv = ops.linear(entry[:, :r].contiguous(), w["wkv_up"]).reshape(t, nh, dn + dv)
    k = torch.cat([kv[..., :dn], entry[:, r:].reshape(t, 1, dr).expand(t, nh, dr)], dim=-1)
    return k, kv[..., dn:]


def attention_forward(cfg: ModelConfig, w: dict, h: torch.Tensor, inp, kv_sink=None) -> torch.Tensor:
    """Attention of one block on packed rows (training, old-log-prob, engine prefill): h:[T,H] normed input -> [T, nh * v] (the
    input of ``wo``).  ``kv_sink`` = (latent pool [blocks, 1, page, r + rope], _, slot mapping): also appends this chunk's cache
    entries, and with ``inp.kv_gather`` attends over the cached prefix as well (chunked prefill / prefix cache).  Under Ulysses
    sequence parallelism (``inp.sp_group``) h holds this rank's token slice."""
    t = h.shape[0]
    dv = cfg.v_head_dim
    q, entry = project(cfg, w, h, inp.position_ids)
    scale = softmax_scale(cfg)
    if kv_sink is not None:
        pool, slots = kv_sink[0], kv_sink[2]
        page = pool.shape[2]
        pool[torch.div(slots, page, rounding_mode="floor"), 0, slots % page] = entry.to(pool.dtype)
        if inp.kv_gather is not None:
            blk, off, cu_k, max_k = inp.kv_gather
            k, v = expand_kv(cfg, w, pool[blk, 0, off].to(h.dtype))
            v = F.pad(v, (0, cfg.head_dim - dv))
            return attention_varlen_prefix(q, k, v, inp.cu_seqlens, cu_k, inp.max_seqlen, max_k, scale=scale)[..., :dv].reshape(t, -1)
    k, v = expand_kv(cfg, w, entry)
    v = F.pad(v, (0, cfg.head_dim - dv))
    if inp.sp_group is not None:
        # Ulysses SP: the projections ran on this rank's token slice; the expanded per-head q / k / v swap tokens for heads around
        # the attention (reference capability: 3rdparty/verl/verl/models/transformers/kimi_vl.py, Ulysses over the DeepSeek-V3 MLA)
        from ..parallel.ulysses import ulysses_attention

        a = ulysses_attention(q, k, v, inp.cu_seqlens, inp.max_seqlen, inp.sp_group, inp.total_tokens, causal=True, scale=scale)
        return a[..., :dv].reshape(t, -1)
    return attention_varlen(q, k, v, inp.cu_seqlens, inp.max_seqlen, causal=True, scale=scale)[..., :dv].reshape(t, -1)


def paged_latent_attention(q: torch.Tensor, pool: torch.Tensor, block_table: torch.Tensor, seq_lens: torch.Tensor, r: int,
                           scale: float) -> torch.Tensor:
    """Multi-query attention over the paged latent pool, static shapes (CUDA-graph capturable).  q:[B, nh, r + rope] (absorbed
    queries), pool:[blocks, 1, page, r + rope], block_table:[B, max_blocks], seq_lens:[B] -> [B, nh, r] (softmax-weighted latents)."""
    b = q.shape[0]
    page = pool.shape[2]
    ent = pool[block_table.long(), 0].reshape(b, block_table.shape[1] * page, pool.shape[-1])          # [B, L, r + rope]
    s = torch.bmm(q, ent.transpose(1, 2)).float() * scale                                              # tensor-core GEMM in the cache dtype, fp32 softmax
    dead = torch.arange(ent.shape[1], device=q.device)[None, :] >= seq_lens.long().clamp(min=1)[:, None]      # idle slots: finite rows
    p = torch.softmax(s.masked_fill(dead[:, None, :], float("-inf")), dim=-1)
    return torch.bmm(p.to(ent.dtype), ent[..., :r])


def attention_decode(cfg: ModelConfig, w: dict, h: torch.Tensor, inp, pool: torch.Tensor) -> torch.Tensor:
    """One new token per sequence in the absorbed form: h:[B,H] -> [B, nh * v]; appends the token's cache entry first."""
    b = h.shape[0]
    nh, dn, dv, r = cfg.num_heads, cfg.qk_nope_head_dim, cfg.v_head_dim, cfg.kv_lora_rank
    q, entry = project(cfg, w, h, inp.position_ids)
    page = pool.shape[2]
    pool[torch.div(inp.slot_mapping, page, rounding_mode="floor"), 0, inp.slot_mapping % page] = entry.to(pool.dtype)
    up = w["wkv_up"].to(h.dtype).view(nh, dn + dv, r)
    q_lat = torch.einsum("bhd,hdc->bhc", q[..., :dn], up[:, :dn])                     # fold W_uk into the query
    o_lat = paged_latent_attention(torch.cat([q_lat, q[..., dn:]], dim=-1), pool, inp.block_table, inp.seq_lens, r, softmax_scale(cfg))
    return torch.einsum("bhc,hdc->bhd", o_lat, up[:, dn:]).reshape(b, nh * dv)       # fold W_uv into the output


# --------------------------------------------------------------------------------------------
# HF DeepseekV3 checkpoint naming
# --------------------------------------------------------------------------------------------
def _deinterleave(cfg: ModelConfig, t: torch.Tensor, inverse: bool = False) -> torch.Tensor:
    """Rows [..., rope, in] of a rotary projection: HF ``rope_interleave`` order (2j, 2j+1 pairs) <-> half-split order."""
    if not cfg.rope_interleaved:
        return t
    dr = cfg.qk_rope_head_dim
    j = torch.arange(dr // 2)
    perm = torch.cat([2 * j, 2 * j + 1])                   # ours[i] = hf[perm[i]]
    idx = torch.argsort(perm) if inverse else perm
    return t.index_select(-2, idx.to(t.device))


def from_hf(cfg: ModelConfig, hf_sd: dict, p: str) -> dict[str, torch.Tensor]:
    """Attention tensors of block prefix ``p`` ('model.layers.i.') in our names."""
    nh, dn, dr = cfg.num_heads, cfg.qk_nope_head_dim, cfg.qk_rope_head_dim
    a = p + "self_attn."

    def q_rows(wq: torch.Tensor) -> torch.Tensor:          # [nh * (nope + rope), in]: de-interleave every head's rotary rows
        wq = wq.reshape(nh, dn + dr, -1)
        return torch.cat([wq[:, :dn], _deinterleave(cfg, wq[:, dn:])], dim=1).reshape(nh * (dn + dr), -1)

    kva = hf_sd[a + "kv_a_proj_with_mqa.weight"]
    kva = torch.cat([kva[: cfg.kv_lora_rank], _deinterleave(cfg, kva[cfg.kv_lora_rank:])], dim=0)
    out = {"kv_a_ln": hf_sd[a + "kv_a_layernorm.weight"], "wkv_up": hf_sd[a + "kv_b_proj.weight"], "wo": hf_sd[a + "o_proj.weight"]}
    if cfg.q_lora_rank:
        out.update(wa=torch.cat([hf_sd[a + "q_a_proj.weight"], kva], dim=0), q_a_ln=hf_sd[a + "q_a_layernorm.weight"],
                   wq_up=q_rows(hf_sd[a + "q_b_proj.weight"]))
    else:
        out["wa"] = torch.cat([q_rows(hf_sd[a + "q_proj.weight"]), kva], dim=0)
    return out


def to_hf(cfg: ModelConfig, sd: dict, q: str, p: str) -> dict[str, torch.Tensor]:
    """Inverse of ``from_hf``: our block tensors (prefix ``q`` = 'layers.i.') -> HF names under ``p``."""
    nh, dn, dr, r = cfg.num_heads, cfg.qk_nope_head_dim, cfg.qk_rope_head_dim, cfg.kv_lora_rank
    a = p + "self_attn."

    def q_rows(wq: torch.Tensor) -> torch.Tensor:
        wq = wq.reshape(nh, dn + dr, -1)
        return torch.cat([wq[:, :dn], _deinterleave(cfg, wq[:, dn:], inverse=True)], dim=1).reshape(nh * (dn + dr), -1)

    wa = sd[q + "wa"]
    qw = cfg.q_lora_rank or nh * (dn + dr)
    out = {a + "kv_a_proj_with_mqa.weight": torch.cat([wa[qw:qw + r], _deinterleave(cfg, wa[qw + r:], inverse=True)], dim=0),
           a + "kv_a_layernorm.weight": sd[q + "kv_a_ln"], a + "kv_b_proj.weight": sd[q + "wkv_up"], a + "o_proj.weight": sd[q + "wo"]}
    if cfg.q_lora_rank:
        out.update({a + "q_a_proj.weight": wa[:qw], a + "q_a_layernorm.weight": sd[q + "q_a_ln"], a + "q_b_proj.weight": q_rows(sd[q + "wq_up"])})
    else:
        out[a + "q_proj.weight"] = q_rows(wa[:qw])
    return out
