"""Packed (varlen) causal self-attention for training / prefill.

Backends, in order: the own tcgen05 / TMEM flash-attention kernels (csrc/attn_fwd.cu, csrc/attn_bwd.cu: head_dim 128, bf16,
GQA, bottom-right aligned causal mask for prefill behind a cached prefix) on CUDA; the flash-attn varlen library (what the
reference calls, 3rdparty/verl/verl/models/transformers/monkey_patch.py:108) for other head sizes / dtypes or when
``POLYRL_ATTN_BACKEND=flash``; PyTorch SDPA over a padded [B, S] view otherwise (CPU tests).  Decode-time attention is NOT
here: it is the hand-written paged kernel in csrc/decode_attn.cu.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

import os

_FLASH = None


def _own_ok(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> bool:
    return (q.is_cuda and q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and v.dtype == torch.bfloat16
            and q.shape[-1] == 128 and k.shape[-1] == 128 and q.shape[1] % k.shape[1] == 0
            and os.environ.get("POLYRL_ATTN_BACKEND", "own") == "own" and os.environ.get("POLYRL_GEMM_BACKEND", "tcgen05") != "cublas")


def _rows(t: torch.Tensor) -> torch.Tensor:
    """[T, heads, 128] view whose heads are contiguous inside a token row (row pitch free): what the kernels' TMA maps need."""
    if t.stride(2) != 1 or t.stride(1) != t.shape[2] or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t


class _OwnAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, cu_q, cu_k, scale, causal):
        from . import ext

        q, k, v = _rows(q), _rows(k), _rows(v)
        need = q.requires_grad or k.requires_grad or v.requires_grad
        outs = ext().attn_fwd(q, k, v, cu_q, cu_k, float(scale), bool(causal), bool(need))
        if need:
            ctx.save_for_backward(q, k, v, outs[0], outs[1], cu_q, cu_k)
            ctx.scale, ctx.causal = float(scale), bool(causal)
        return outs[0]

    @staticmethod
    def backward(ctx, do):
        from . import ext

        q, k, v, o, lse, cu_q, cu_k = ctx.saved_tensors
        dq, dk, dv = ext().attn_bwd(q, k, v, o, _rows(do), lse, cu_q, cu_k, ctx.scale, ctx.causal)
        return dq, dk, dv, None, None, None, None


def _own_attention(q, k, v, cu_q, cu_k, scale, causal=True):
    cu_q = cu_q if cu_q.dtype == torch.int32 else cu_q.to(torch.int32)
    cu_k = cu_k if cu_k.dtype == torch.int32 else cu_k.to(torch.int32)
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
        return _OwnAttnFn.apply(q, k, v, cu_q.contiguous(), cu_k.contiguous(), scale, causal)
    from . import ext

    return ext().attn_fwd(_rows(q), _rows(k), _rows(v), cu_q.contiguous(), cu_k.contiguous(), float(scale), bool(causal), False)[0]


def _flash():
    global _FLASH
    if _FLASH is None:
        try:
            from flash_attn import flash_attn_varlen_func

            _FLASH = flash_attn_varlen_func
        except Exception:  # pragma: no cover
            _FLASH = False
    return _FLASH


def _sdpa_varlen(q, k, v, cu_seqlens, max_seqlen, causal, scale):
    """q:[T,nq,d] k,v:[T,nkv,d] -> [T,nq,d] via a padded batch (supports autograd)."""
    t, nq, d = q.shape
    nkv = k.shape[1]
    b = cu_seqlens.numel() - 1
    lens = (cu_seqlens[1:] - cu_seqlens[:-1]).long()
    s = int(max_seqlen)
    seq_id = torch.repeat_interleave(torch.arange(b, device=q.device), lens, output_size=t)
    pos = torch.arange(t, device=q.device) - cu_seqlens[:-1].long()[seq_id]
    flat = seq_id * s + pos

    def pad(x):
        out = x.new_zeros(b * s, x.shape[1], x.shape[2])
        out = out.index_copy(0, flat, x)
        return out.view(b, s, x.shape[1], x.shape[2]).transpose(1, 2)      # [B, h, S, d]

    qp, kp, vp = pad(q), pad(k), pad(v)
    if nkv != nq:
        kp = kp.repeat_interleave(nq // nkv, dim=1)
        vp = vp.repeat_interleave(nq // nkv, dim=1)
    key_ok = torch.arange(s, device=q.device)[None, :] < lens[:, None]          # [B, S]
    mask = key_ok[:, None, None, :]
    if causal:
        mask = mask & torch.ones(s, s, dtype=torch.bool, device=q.device).tril()[None, None]
    o = F.scaled_dot_product_attention(qp, kp, vp, attn_mask=mask, scale=scale)
    return o.transpose(1, 2).reshape(b * s, nq, d).index_select(0, flat)


def attention_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int,
                     causal: bool = True, scale: float | None = None) -> torch.Tensor:
    """q:[T,nq,d], k/v:[T,nkv,d] packed over sequences delimited by ``cu_seqlens`` (int32 [B+1])."""
    if scale is None:
        scale = q.shape[-1] ** -0.5
    if _own_ok(q, k, v):
        return _own_attention(q, k, v, cu_seqlens, cu_seqlens, scale, causal)
    fa = _flash() if q.is_cuda and q.dtype in (torch.bfloat16, torch.float16) else False
    if fa:
        return fa(q, k, v, cu_seqlens, cu_seqlens, int(max_seqlen), int(max_seqlen), 0.0, softmax_scale=scale,
                  causal=causal)
    return _sdpa_varlen(q, k, v, cu_seqlens, max_seqlen, causal, scale)


def attention_varlen_prefix(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_q: torch.Tensor, cu_k: torch.Tensor, max_q: int,
                            max_k: int, scale: float | None = None) -> torch.Tensor:
    """Prefill behind a cached prefix: sequence i has ``lq_i`` NEW query tokens and ``lk_i >= lq_i`` keys (cached prefix first,
    then the new tokens); causal mask aligned bottom-right.  q:[Tq,nq,d], k/v:[Tk,nkv,d]."""
    if scale is None:
        scale = q.shape[-1] ** -0.5
    if _own_ok(q, k, v):
        return _own_attention(q, k, v, cu_q, cu_k, scale, True)
    fa = _flash() if q.is_cuda and q.dtype in (torch.bfloat16, torch.float16) else False
    if fa:
        return fa(q, k, v, cu_q, cu_k, int(max_q), int(max_k), 0.0, softmax_scale=scale, causal=True)
    nq, nkv = q.shape[1], k.shape[1]
    out = torch.empty_like(q)
    cq, ck = cu_q.tolist(), cu_k.tolist()
    for i in range(len(cq) - 1):
        qs, ks, vs = q[cq[i]:cq[i + 1]], k[ck[i]:ck[i + 1]], v[ck[i]:ck[i + 1]]
        lq, lk = qs.shape[0], ks.shape[0]
        if nkv != nq:
            ks, vs = ks.repeat_interleave(nq // nkv, dim=1), vs.repeat_interleave(nq // nkv, dim=1)
        mask = torch.arange(lk, device=q.device)[None, :] <= (torch.arange(lq, device=q.device)[:, None] + (lk - lq))
        o = F.scaled_dot_product_attention(qs.transpose(0, 1), ks.transpose(0, 1), vs.transpose(0, 1), attn_mask=mask, scale=scale)
        out[cq[i]:cq[i + 1]] = o.transpose(0, 1)
    return out
