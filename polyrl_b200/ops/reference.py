"""Plain-PyTorch reference implementations of every fused op (fp32 math).

Used (a) as the CPU execution path so the whole framework runs under gloo without a GPU, and (b) as
the numerical oracle for the sm_100a kernels in ``tests/``.  Math follows SURVEY Appendix E
(reference: 3rdparty/verl/verl/utils/kernel/kernels.py:242-266,343-371,1237-1245).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def split_gate_up(gu: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Undo the block-interleaved [128 gate | 128 up] column layout of the fused gate/up projection."""
    *lead, n2 = gu.shape
    v = gu.reshape(*lead, n2 // 256, 2, 128)
    return v[..., 0, :].reshape(*lead, n2 // 2), v[..., 1, :].reshape(*lead, n2 // 2)


def interleave_gate_up(w_gate: torch.Tensor, w_up: torch.Tensor) -> torch.Tensor:
    """Pack separate gate / up weights [F, H] into the interleaved [2F, H] layout (F % 128 == 0)."""
    f, h = w_gate.shape
    assert f % 128 == 0, "intermediate size must be a multiple of 128"
    g = w_gate.reshape(f // 128, 1, 128, h)
    u = w_up.reshape(f // 128, 1, 128, h)
    return torch.cat([g, u], dim=1).reshape(2 * f, h)


def deinterleave_gate_up(w_gu: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    f2, h = w_gu.shape
    v = w_gu.reshape(f2 // 256, 2, 128, h)
    return v[:, 0].reshape(f2 // 2, h), v[:, 1].reshape(f2 // 2, h)


def swiglu(gu: torch.Tensor) -> torch.Tensor:
    g, u = split_gate_up(gu.float())
    return (F.silu(g) * u).to(gu.dtype)


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (xf * rstd * w.float()).to(x.dtype)


def rope_tables(max_pos: int, d: int, theta: float, device=None) -> tuple[torch.Tensor, torch.Tensor]:
    inv = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float32, device=device) / d))
    t = torch.arange(max_pos, dtype=torch.float32, device=device)
    fr = torch.outer(t, inv)
    return fr.cos().contiguous(), fr.sin().contiguous()


def rope(qkv: torch.Tensor, pos: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, nq: int, nkv: int, d: int,
         sign: float = 1.0) -> torch.Tensor:
    """Rotate-half RoPE on the q and k heads of packed qkv rows [T, (nq+2nkv)*d]; v untouched."""
    t = qkv.shape[0]
    half = d // 2
    out = qkv.clone()
    rot = out[:, : (nq + nkv) * d].reshape(t, nq + nkv, d).float()
    c = cos[pos].unsqueeze(1)
    s = sin[pos].unsqueeze(1) * sign
    x1, x2 = rot[..., :half], rot[..., half:]
    r = torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)
    out[:, : (nq + nkv) * d] = r.reshape(t, -1).to(qkv.dtype)
    return out


def linear_cross_entropy(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor,
                         temperature: float = 1.0) -> tuple[torch.Tensor, torch.Tensor]:
    """log-prob of ``labels`` and entropy of softmax(hidden @ weight.T / temperature), fp32."""
    z = (hidden.float() @ weight.float().t()) / temperature
    lse = torch.logsumexp(z, dim=-1)
    lp = z.gather(-1, labels.unsqueeze(-1)).squeeze(-1) - lse
    p = torch.softmax(z, dim=-1)
    ent = lse - (p * z).sum(-1)
    return lp, ent


def paged_decode_attention(q, k_cache, v_cache, block_table, seq_lens, nq, nkv, d, sm_scale):
    """q:[B,nq*d]; caches [num_blocks, nkv, page, d]; returns [B, nq*d]."""
    b = q.shape[0]
    page = k_cache.shape[2]
    out = torch.empty_like(q)
    g = nq // nkv
    for i in range(b):
        n = int(seq_lens[i])
        npg = (n + page - 1) // page
        blks = block_table[i, :npg].long()
        k = k_cache[blks].permute(1, 0, 2, 3).reshape(nkv, npg * page, d)[:, :n].float()
        v = v_cache[blks].permute(1, 0, 2, 3).reshape(nkv, npg * page, d)[:, :n].float()
        qi = q[i].reshape(nkv, g, d).float()
        s = torch.einsum("hgd,htd->hgt", qi, k) * sm_scale
        p = torch.softmax(s, dim=-1)
        out[i] = torch.einsum("hgt,htd->hgd", p, v).reshape(-1).to(q.dtype)
    return out


def adamw_step(master, m, v, grad, lr, beta1, beta2, eps, wd, step, clip_coef=1.0):
    g = grad * clip_coef
    master.mul_(1 - lr * wd)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v / bc2).sqrt().add_(eps)
    master.addcdiv_(m, denom, value=-lr / bc1)
    return master
