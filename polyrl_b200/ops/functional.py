"""Autograd-aware wrappers over the sm_100a kernels (CUDA) and the PyTorch references (CPU)."""
from __future__ import annotations

import os

import torch

from . import reference as ref


def _C():
    from . import ext

    return ext()


def _cublas() -> bool:
    return os.environ.get("POLYRL_GEMM_BACKEND", "tcgen05") == "cublas"


def _is_kernel_tensor(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.bfloat16


def _g2_operand(t: torch.Tensor) -> torch.Tensor:
    """Operand of the CTA-pair GEMM family: unit inner stride, row pitch a multiple of 8 elements, 16-byte aligned base."""
    if t.stride(-1) != 1 or t.stride(0) % 8 != 0 or t.stride(0) < t.shape[-1] or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t


def _g2_ok(*dims: int) -> bool:
    return all(d % 8 == 0 for d in dims)


# ---- gradient sink: weight gradients written by the wgrad GEMM straight into caller-provided storage ----
# The ZeRO engine registers, per weight of the unit whose backward is about to run, a view of the symmetric reduce-scatter slot
# (parallel/zero3.py::_UnitFn.backward).  The wgrad GEMM's TMA store then lands in the buffer the peers will read: no staging
# copy of the unit's gradients (2 x 16 GB of HBM traffic per micro-batch for an 8B model).
_GRAD_SINK: dict[int, torch.Tensor] = {}


class grad_sink:
    def __init__(self, mapping: dict[int, torch.Tensor]):
        self.mapping = mapping

    def __enter__(self):
        _GRAD_SINK.update(self.mapping)
        return self

    def __exit__(self, *exc):
        for k in self.mapping:
            _GRAD_SINK.pop(k, None)
        return False


def _sink_for(w: torch.Tensor) -> torch.Tensor | None:
    if not _GRAD_SINK:
        return None
    t = _GRAD_SINK.get(w.data_ptr())
    return t if t is not None and t.shape == w.shape and t.dtype == w.dtype else None


def mm_nn(dy: torch.Tensor, w: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """dgrad: ``dy[M, N] @ w[N, K]`` -> [M, K].  ``w`` is consumed as an MN-major tcgen05 operand (no transpose copy)."""
    if _cublas() or not (_is_kernel_tensor(dy) and w.dtype == torch.bfloat16 and _g2_ok(dy.shape[1], w.shape[1])):
        return torch.matmul(dy, w, out=out) if out is not None else dy @ w
    return _C().gemm2_bf16(_g2_operand(dy), _g2_operand(w), False, True, None, out, 0)


def mm_tn(dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """wgrad: ``dy[M, N].T @ x[M, K]`` -> [N, K].  Both operands are consumed MN-major straight from their row-major storage."""
    if _cublas() or not (_is_kernel_tensor(dy) and x.dtype == torch.bfloat16 and _g2_ok(dy.shape[1], x.shape[1])):
        return torch.mm(dy.t(), x, out=out) if out is not None else dy.t() @ x
    return _C().gemm2_bf16(_g2_operand(dy), _g2_operand(x), True, True, None, out, 0)


# ---- decode-shape ("skinny") GEMM: swap-AB + split-K tcgen05 kernel for <= 256 rows, inference only ----
_SKINNY_WS: dict = {}
_SKINNY_MAX_ROWS = int(os.environ.get("POLYRL_SKINNY_MAX_ROWS", "256"))


def _skinny_buffers(device):
    key = (device.type, device.index)
    if key not in _SKINNY_WS:
        ws = torch.empty(24 * 1024 * 1024, device=device, dtype=torch.float32)       # 96 MB of split-K partials
        ctr = torch.zeros(4096, device=device, dtype=torch.int32)
        _SKINNY_WS[key] = (ws, ctr)
    return _SKINNY_WS[key]


def _decode_policy() -> str:
    """Decode-shape (<= 256 rows, inference) GEMM backend.  ``auto`` follows the in-graph measurements in
    profiles/r1_decode_gemm_in_graph_v0.jsonl: library GEMM for plain projections, own fused tile kernel for gate/up+SwiGLU;
    ``skinny`` forces the swap-AB split-K kernel, ``tile`` the 128-row-tile kernel, ``cublas`` the library everywhere."""
    return os.environ.get("POLYRL_DECODE_GEMM", "auto")


def _is_decode_shape(x2: torch.Tensor, w: torch.Tensor) -> bool:
    return (not torch.is_grad_enabled() or not (x2.requires_grad or w.requires_grad)) and x2.shape[0] <= _SKINNY_MAX_ROWS


def _use_skinny(x2: torch.Tensor, w: torch.Tensor) -> bool:
    if _cublas() or _decode_policy() != "skinny":
        return False
    return _is_decode_shape(x2, w) and x2.shape[1] % 16 == 0 and w.shape[0] >= 128


def skinny_linear(x2: torch.Tensor, w: torch.Tensor, bias=None, residual=None, mode: int = 0, splits: int = 0) -> torch.Tensor:
    """Own decode-shape GEMM: the round-2 stream GEMM (cluster split-K + DSMEM reduce-scatter, csrc/decode_gemm.cu); the round-1
    swap-AB kernel remains behind ``POLYRL_DECODE_GEMM_IMPL=skinny`` (it still backs the fused TP all-reduce epilogue)."""
    if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0:
        x2 = x2.contiguous()
    if residual is not None and not residual.is_contiguous():
        residual = residual.contiguous()
    n_out = w.shape[0] // 2 if mode == 1 else w.shape[0]
    if os.environ.get("POLYRL_DECODE_GEMM_IMPL", "stream") == "stream" and n_out % 8 == 0:
        return _C().stream_gemm(x2, w, residual, bias, None, None, mode, splits, None, -1)
    ws, ctr = _skinny_buffers(x2.device)
    return _C().skinny_gemm(x2, w, residual, bias, None, None, ws, ctr, mode, splits)


# --------------------------------------------------------------------------------------------
# fp8 (e4m3, per-output-row scale) inference weights: rollout-side only (BASELINE config "fp8 rollout, bf16 update")
# --------------------------------------------------------------------------------------------
class Fp8Weight:
    """``q`` [N, K] float8_e4m3fn + ``scale`` [N] fp32 (w ~= q * scale[:, None]).  Activations are quantised per token
    on the fly; the GEMM is the tcgen05 kind::f8f6f4 kernel with both scale vectors applied in the epilogue."""
    __slots__ = ("q", "scale")

    def __init__(self, q: torch.Tensor, scale: torch.Tensor):
        self.q, self.scale = q, scale

    @property
    def shape(self):
        return self.q.shape

    @property
    def dtype(self):
        return self.q.dtype

    def dequant(self, dtype=torch.float32) -> torch.Tensor:
        return (self.q.float() * self.scale[:, None]).to(dtype)


def quantize_fp8_rowwise(x: torch.Tensor):
    if x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2:
        q, s = _C().quant_fp8_rowwise(x.contiguous())
        return q, s
    xf = x.float()
    s = xf.abs().amax(-1).clamp_min(1e-12) / 448.0
    return (xf / s[:, None]).to(torch.float8_e4m3fn), s


def quantize_weight_fp8(w: torch.Tensor) -> Fp8Weight:
    return Fp8Weight(*quantize_fp8_rowwise(w))


def _linear_fp8(x: torch.Tensor, w: Fp8Weight, bias=None, residual=None, swiglu: bool = False) -> torch.Tensor:
    if x.is_cuda and x.dtype == torch.bfloat16:
        xq, xs = quantize_fp8_rowwise(x)
        n = w.q.shape[0]
        if x.shape[0] <= _SKINNY_MAX_ROWS and n >= 128 and (not swiglu or n % 256 == 0):
            # stream GEMM: fp8 operands, fused scales / residual / SwiGLU, cluster split-K
            r = residual.contiguous() if residual is not None else None
            if (n // 2 if swiglu else n) % 8 == 0 and os.environ.get("POLYRL_DECODE_GEMM_IMPL", "stream") == "stream":
                return _C().stream_gemm(xq, w.q, r, bias, xs, w.scale, 1 if swiglu else 0, 0, None, 0)
            wsb, ctr = _skinny_buffers(x.device)
            return _C().skinny_gemm(xq, w.q, r, bias, xs, w.scale, wsb, ctr, 1 if swiglu else 0, 0)
        y = _C().gemm_fp8(xq, w.q, xs, w.scale, 0)
    else:                                                       # CPU emulation: same quantisation, fp32 matmul
        xq, xs = quantize_fp8_rowwise(x)
        y = ((xq.float() * xs[:, None]) @ w.dequant().t()).to(x.dtype)
    if swiglu:
        return ref.swiglu(y)
    if bias is not None:
        y = y + bias.to(y.dtype)
    return y + residual if residual is not None else y


# --------------------------------------------------------------------------------------------
# linear (+ residual) : tcgen05 GEMM forward with fused residual epilogue
# --------------------------------------------------------------------------------------------
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, residual):
        x2 = x.reshape(-1, x.shape[-1])
        r2 = residual.reshape(-1, w.shape[0]) if residual is not None else None
        if _cublas():
            y = x2 @ w.t()
            if r2 is not None:
                y = y + r2
        else:
            if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0:
                x2 = x2.contiguous()
            if r2 is not None and not r2.is_contiguous():
                r2 = r2.contiguous()
            y = _C().gemm_bf16(x2, w, r2)
        ctx.save_for_backward(x2, w)
        ctx.has_res = residual is not None
        ctx.x_shape = x.shape
        return y if x.dim() == 2 else y.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        dx = dw = dres = None
        if ctx.needs_input_grad[0]:
            dx = mm_nn(dy2, w).reshape(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = mm_tn(dy2, x2, out=_sink_for(w))
        if ctx.has_res and ctx.needs_input_grad[2]:
            dres = dy
        return dx, dw, dres


def linear(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None,
           residual: torch.Tensor | None = None) -> torch.Tensor:
    """y = x @ w.T (+ bias) (+ residual).  w is [out, in] (nn.Linear layout)."""
    if isinstance(w, Fp8Weight):
        return _linear_fp8(x, w, bias, residual)
    if _is_kernel_tensor(x) and w.dtype == torch.bfloat16 and x.dim() == 2 and _use_skinny(x, w):
        return skinny_linear(x, w, bias, residual, 0)
    if _is_kernel_tensor(x) and w.dtype == torch.bfloat16 and x.dim() == 2 and _decode_policy() in ("auto", "cublas") \
            and _is_decode_shape(x, w):
        y = torch.nn.functional.linear(x, w, bias)
        return y.add_(residual) if residual is not None else y
    if _is_kernel_tensor(x) and w.dtype == torch.bfloat16 and w.shape[0] % 8 == 0 and w.shape[1] % 8 == 0:
        y = _LinearFn.apply(x, w, residual if bias is None else None)
        if bias is not None:
            y = y + bias
            if residual is not None:
                y = y + residual
        return y
    if x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16:
        # odd shapes (e.g. vocab not multiple of 8): library GEMM
        y = torch.nn.functional.linear(x, w, bias)
        return y + residual if residual is not None else y
    y = torch.nn.functional.linear(x, w.to(x.dtype), None if bias is None else bias.to(x.dtype))
    return y + residual if residual is not None else y


# --------------------------------------------------------------------------------------------
# gate/up projection with SwiGLU fused into the GEMM epilogue
# --------------------------------------------------------------------------------------------
class _LinearSwigluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wgu):
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1 or x2.stride(0) % 8 != 0:
            x2 = x2.contiguous()
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        outs = _C().gemm_swiglu(x2, wgu, bool(need))
        h = outs[0]
        if need:
            ctx.save_for_backward(x2, wgu, outs[1])
        ctx.x_shape = x.shape
        return h if x.dim() == 2 else h.reshape(*x.shape[:-1], wgu.shape[0] // 2)

    @staticmethod
    def backward(ctx, dh):
        x2, wgu, gu = ctx.saved_tensors
        dh2 = dh.reshape(-1, dh.shape[-1]).contiguous()
        dgu = _C().swiglu_bwd(gu, dh2)
        dx = mm_nn(dgu, wgu).reshape(ctx.x_shape) if ctx.needs_input_grad[0] else None
        dw = mm_tn(dgu, x2, out=_sink_for(wgu)) if ctx.needs_input_grad[1] else None
        return dx, dw


def linear_swiglu(x: torch.Tensor, wgu: torch.Tensor) -> torch.Tensor:
    """silu(x @ Wg.T) * (x @ Wu.T) with W_gu stored block-interleaved ([128 gate | 128 up] rows)."""
    if isinstance(wgu, Fp8Weight):
        return _linear_fp8(x, wgu, swiglu=True)
    if _is_kernel_tensor(x) and wgu.dtype == torch.bfloat16 and x.dim() == 2 and _use_skinny(x, wgu):
        return skinny_linear(x, wgu, None, None, 1)
    if _is_kernel_tensor(x) and wgu.dtype == torch.bfloat16 and not _cublas():
        return _LinearSwigluFn.apply(x, wgu)
    gu = torch.nn.functional.linear(x, wgu.to(x.dtype))
    if x.is_cuda and x.dtype == torch.bfloat16:
        # library arm on GPUs: what an HF decoder layer does — act_fn(gate) * up in bf16, three eager kernels
        g, u = ref.split_gate_up(gu)
        return torch.nn.functional.silu(g) * u
    return ref.swiglu(gu)


# --------------------------------------------------------------------------------------------
# RMSNorm
# --------------------------------------------------------------------------------------------
class _RmsNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        xc = x.contiguous()
        y, rstd = _C().rmsnorm_fwd(xc, w, eps)
        ctx.save_for_backward(xc, w, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, rstd = ctx.saved_tensors
        dx, dw = _C().rmsnorm_bwd(x, w, dy.contiguous(), rstd)
        return dx, dw.to(w.dtype), None


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    if _is_kernel_tensor(x) and w.dtype == torch.bfloat16 and x.shape[-1] % 8 == 0 and x.shape[-1] <= 8192:
        return _RmsNormFn.apply(x, w, eps)
    return ref.rmsnorm(x, w, eps)


# --------------------------------------------------------------------------------------------
# RoPE (in place on packed qkv), optional paged-KV append
# --------------------------------------------------------------------------------------------
class _RopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, pos, cos, sin, nq, nkv, d):
        ctx.mark_dirty(qkv)
        _C().rope_(qkv, pos, cos, sin, nq, nkv, d, 1.0, None, None, None, 16)
        ctx.save_for_backward(pos, cos, sin)
        ctx.dims = (nq, nkv, d)
        return qkv

    @staticmethod
    def backward(ctx, dy):
        pos, cos, sin = ctx.saved_tensors
        nq, nkv, d = ctx.dims
        g = dy.contiguous().clone()
        _C().rope_(g, pos, cos, sin, nq, nkv, d, -1.0, None, None, None, 16)
        return g, None, None, None, None, None, None


def apply_rope(qkv: torch.Tensor, pos: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, nq: int, nkv: int,
               d: int) -> torch.Tensor:
    """Rotate q/k heads of packed rows [T, (nq+2*nkv)*d].  CUDA path works in place on ``qkv``."""
    if _is_kernel_tensor(qkv):
        if not qkv.requires_grad and not torch.is_grad_enabled():
            _C().rope_(qkv, pos, cos, sin, nq, nkv, d, 1.0, None, None, None, 16)
            return qkv
        return _RopeFn.apply(qkv, pos, cos, sin, nq, nkv, d)
    return ref.rope(qkv, pos, cos, sin, nq, nkv, d)


def rope_and_cache_(qkv, pos, cos, sin, nq, nkv, d, k_cache, v_cache, slot_mapping, page: int = 16):
    """Inference: rotate in place and append rotated k / v to the paged cache at ``slot_mapping``."""
    if _is_kernel_tensor(qkv):
        _C().rope_(qkv, pos, cos, sin, nq, nkv, d, 1.0, k_cache, v_cache, slot_mapping, page)
        return qkv
    out = ref.rope(qkv, pos, cos, sin, nq, nkv, d)
    t = out.shape[0]
    k = out[:, nq * d:(nq + nkv) * d].reshape(t, nkv, d)
    v = out[:, (nq + nkv) * d:].reshape(t, nkv, d)
    blk = torch.div(slot_mapping, page, rounding_mode="floor")
    off = slot_mapping % page
    k_cache[blk, :, off] = k.to(k_cache.dtype)
    v_cache[blk, :, off] = v.to(v_cache.dtype)
    qkv.copy_(out)
    return qkv


# --------------------------------------------------------------------------------------------
# fused LM head: log-prob + entropy without materialising logits
# --------------------------------------------------------------------------------------------
_CE_CHUNK = int(os.environ.get("POLYRL_CE_BWD_CHUNK", "16384"))


class _LinearCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight, labels, inv_tau):
        h2 = hidden.reshape(-1, hidden.shape[-1])
        if not h2.is_contiguous():
            h2 = h2.contiguous()
        lab = labels.reshape(-1).contiguous()
        lp, ent, lse, ez = _C().linear_ce_fwd(h2, weight, lab, inv_tau)
        ctx.save_for_backward(h2, weight, lab, lse, ez)
        ctx.inv_tau = inv_tau
        ctx.h_shape = hidden.shape
        return lp.reshape(labels.shape), ent.reshape(labels.shape)

    @staticmethod
    def backward(ctx, g_lp, g_ent):
        h2, w, lab, lse, ez = ctx.saved_tensors
        C = _C()
        g_lp = g_lp.reshape(-1).float().contiguous() if g_lp is not None else None
        g_ent = g_ent.reshape(-1).float().contiguous() if g_ent is not None else None
        V = w.shape[0]
        need_h, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        own = not _cublas() and h2.shape[1] % 8 == 0
        dh = (torch.empty if own else torch.zeros)(h2.shape, device=h2.device, dtype=torch.float32) if need_h else None
        dw = torch.empty_like(w) if need_w else None
        for v0 in range(0, V, _CE_CHUNK):
            v1 = min(V, v0 + _CE_CHUNK)
            ws = w[v0:v1]
            dz = C.linear_ce_dz(h2, ws, lab, lse, ez, g_lp, g_ent, ctx.inv_tau, v0)
            if need_h:
                if own:      # dH (fp32) += dz @ W_slice inside the GEMM epilogue
                    C.gemm2_accum_f32_(dh, dz, ws, False, True, 0.0 if v0 == 0 else 1.0)
                else:
                    dh.add_(dz @ ws)
            if need_w:
                mm_tn(dz, h2, out=dw[v0:v1])
        dh_out = dh.to(h2.dtype).reshape(ctx.h_shape) if need_h else None
        return dh_out, dw, None, None


def _chunked_torch_ce(hidden, weight, labels, temperature, chunk: int = 2048):
    """Library baseline: materialise [chunk, V] logits with cuBLAS, eager softmax statistics, recompute per chunk in backward
    (the reference's non-fused path: verl/utils/experimental/torch_functional.py:75-216 ``FusedLinearForPPO`` chunks of tokens)."""
    from torch.utils.checkpoint import checkpoint

    h2, lab = hidden.reshape(-1, hidden.shape[-1]), labels.reshape(-1)

    def piece(h, w, l):
        z = (h @ w.t()).float() / temperature
        lse = torch.logsumexp(z, -1)
        lp = z.gather(-1, l.unsqueeze(-1)).squeeze(-1) - lse
        ent = lse - (torch.softmax(z, -1) * z).sum(-1)
        return lp, ent

    lps, ents = [], []
    for s in range(0, h2.shape[0], chunk):
        if torch.is_grad_enabled() and (h2.requires_grad or weight.requires_grad):
            lp, ent = checkpoint(piece, h2[s:s + chunk], weight, lab[s:s + chunk], use_reentrant=False)
        else:
            lp, ent = piece(h2[s:s + chunk], weight, lab[s:s + chunk])
        lps.append(lp)
        ents.append(ent)
    return torch.cat(lps).reshape(labels.shape), torch.cat(ents).reshape(labels.shape)


def linear_cross_entropy(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor,
                         temperature: float = 1.0) -> tuple[torch.Tensor, torch.Tensor]:
    """(log p(labels), entropy) of softmax(hidden @ weight.T / temperature); fp32 outputs, shape of labels."""
    if os.environ.get("POLYRL_CE_BACKEND", "fused") == "torch" and hidden.is_cuda:
        return _chunked_torch_ce(hidden, weight, labels, temperature)
    if _is_kernel_tensor(hidden) and weight.dtype == torch.bfloat16 and weight.shape[0] % 8 == 0:
        return _LinearCEFn.apply(hidden, weight, labels, 1.0 / float(temperature))
    lp, ent = ref.linear_cross_entropy(hidden.reshape(-1, hidden.shape[-1]), weight, labels.reshape(-1), temperature)
    return lp.reshape(labels.shape), ent.reshape(labels.shape)


class _LinearCETPFn(torch.autograd.Function):
    """Vocab-parallel fused linear + cross-entropy: every TP rank holds ``weight[V/tp, H]``.  Local online-softmax
    stats come from the same tcgen05 kernel; ONE all-gather of [T, 3] (lse, E[z], label logit) replaces the three
    all-reduces of the reference's TP mode (verl/utils/kernel/kernels.py:642-694), and the backward all-reduces dH."""

    @staticmethod
    def forward(ctx, hidden, w_shard, labels, inv_tau, group, vocab_start):
        import torch.distributed as dist

        h2 = hidden.reshape(-1, hidden.shape[-1])
        if not h2.is_contiguous():
            h2 = h2.contiguous()
        vs = w_shard.shape[0]
        lab_local = (labels.reshape(-1) - vocab_start).contiguous()
        in_shard = (lab_local >= 0) & (lab_local < vs)
        kernel = _is_kernel_tensor(h2) and w_shard.dtype == torch.bfloat16 and vs % 8 == 0
        if kernel:
            lp_l, _, lse_l, ez_l = _C().linear_ce_fwd(h2, w_shard, lab_local, inv_tau)
            zlab = torch.where(in_shard, lp_l + lse_l, torch.zeros_like(lp_l))
        else:
            z = (h2.float() @ w_shard.float().t()) * inv_tau
            lse_l = torch.logsumexp(z, -1)
            ez_l = (torch.softmax(z, -1) * z).sum(-1)
            zlab = torch.where(in_shard, z.gather(-1, lab_local.clamp(0, vs - 1).unsqueeze(-1)).squeeze(-1), torch.zeros_like(lse_l))
        tp = dist.get_world_size(group)
        stats = torch.stack([lse_l, ez_l, zlab], dim=-1).contiguous()                  # [T, 3]
        allst = [torch.empty_like(stats) for _ in range(tp)]
        dist.all_gather(allst, stats, group=group)
        allst = torch.stack(allst)                                                      # [tp, T, 3]
        lse = torch.logsumexp(allst[..., 0], dim=0)
        ez = (torch.exp(allst[..., 0] - lse) * allst[..., 1]).sum(0)
        lp = allst[..., 2].sum(0) - lse
        ent = lse - ez
        ctx.save_for_backward(h2, w_shard, lab_local, lse.contiguous(), ez.contiguous())
        ctx.inv_tau, ctx.group, ctx.kernel, ctx.h_shape = inv_tau, group, kernel, hidden.shape
        return lp.reshape(labels.shape), ent.reshape(labels.shape)

    @staticmethod
    def backward(ctx, g_lp, g_ent):
        import torch.distributed as dist

        h2, w, lab, lse, ez = ctx.saved_tensors
        g_lp = g_lp.reshape(-1).float().contiguous() if g_lp is not None else torch.zeros_like(lse)
        g_ent = g_ent.reshape(-1).float().contiguous() if g_ent is not None else torch.zeros_like(lse)
        vs = w.shape[0]
        need_h, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dh = torch.zeros_like(h2, dtype=torch.float32)
        dw = torch.empty_like(w) if need_w else None
        for v0 in range(0, vs, _CE_CHUNK):
            v1 = min(vs, v0 + _CE_CHUNK)
            ws = w[v0:v1]
            if ctx.kernel:
                dz = _C().linear_ce_dz(h2, ws, lab, lse, ez, g_lp, g_ent, ctx.inv_tau, v0)
            else:
                z = (h2.float() @ ws.float().t()) * ctx.inv_tau
                p = torch.exp(z - lse[:, None])
                onehot = (torch.arange(v0, v1, device=z.device)[None] == lab[:, None]).float()
                dz = ctx.inv_tau * (g_lp[:, None] * (onehot - p) - g_ent[:, None] * p * (z - ez[:, None]))
                dz = dz.to(h2.dtype)
            if need_h:
                if ctx.kernel and not _cublas() and h2.shape[1] % 8 == 0:
                    _C().gemm2_accum_f32_(dh, dz, ws, False, True, 1.0)
                else:
                    dh.add_((dz @ ws).float())
            if need_w:
                mm_tn(dz, h2, out=dw[v0:v1])
        if need_h:
            dist.all_reduce(dh, group=ctx.group)
        return (dh.to(h2.dtype).reshape(ctx.h_shape) if need_h else None), dw, None, None, None, None


def linear_cross_entropy_tp(hidden: torch.Tensor, weight_shard: torch.Tensor, labels: torch.Tensor, temperature: float,
                            group, vocab_start: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Vocab-parallel variant of ``linear_cross_entropy``: ``weight_shard`` = rows [vocab_start, vocab_start + V/tp)."""
    return _LinearCETPFn.apply(hidden, weight_shard, labels, 1.0 / float(temperature), group, int(vocab_start))


# --------------------------------------------------------------------------------------------
# rollout-side ops (no autograd)
# --------------------------------------------------------------------------------------------
def _auto_splits(num_seqs: int, nkv: int, max_len: int) -> int:
    ctas = num_seqs * nkv
    if ctas >= 4 * 148:
        return 1
    want = (4 * 148 + ctas - 1) // ctas
    by_len = max(1, max_len // 256)
    return max(1, min(want, by_len, 32))


@torch.no_grad()
def paged_decode_attention(q, k_cache, v_cache, block_table, seq_lens, nq, nkv, d, sm_scale, splits: int = 0,
                           max_len: int = 0, out=None):
    if _is_kernel_tensor(q):
        if splits <= 0:
            splits = _auto_splits(q.shape[0], nkv, max_len if max_len > 0 else 1024)
        return _C().paged_decode(q, k_cache, v_cache, block_table, seq_lens, nq, nkv, d, sm_scale, splits, out)
    return ref.paged_decode_attention(q, k_cache, v_cache, block_table, seq_lens, nq, nkv, d, sm_scale)


@torch.no_grad()
def sample_tokens(logits: torch.Tensor, temperature: torch.Tensor | float = 1.0, seeds: torch.Tensor | None = None,
                  greedy: bool = False, top_k: int = -1, top_p: float = 1.0):
    """Sample one token per row and return (token ids int64, log-prob of the sampled token fp32)."""
    b, v = logits.shape
    if not torch.is_tensor(temperature):
        temperature = torch.full((b,), float(temperature), device=logits.device, dtype=torch.float32)
    simple = top_k <= 0 and top_p >= 1.0
    if logits.is_cuda and logits.dtype in (torch.bfloat16, torch.float32):
        inv_t = torch.where(temperature > 0, 1.0 / temperature.clamp_min(1e-6), torch.zeros_like(temperature))
        if seeds is None:
            seeds = torch.randint(0, 2 ** 62, (b,), device=logits.device, dtype=torch.int64)
        if simple:
            return tuple(_C().sample(logits, inv_t.float().contiguous(), seeds, bool(greedy)))
        # top-k / top-p: radix-select kernel (no sort); log-prob under the renormalised kept set
        return tuple(_C().sample_topk_topp(logits, inv_t.float().contiguous(), seeds, bool(greedy), int(top_k), float(top_p)))
    z = logits.float()
    temperature = temperature.to(z.device)
    row_greedy = torch.ones_like(temperature, dtype=torch.bool) if greedy else temperature <= 0     # per row, like the kernel
    is_greedy = bool(row_greedy.all())
    z = z / torch.where(row_greedy, torch.ones_like(temperature), temperature).unsqueeze(-1)
    if top_k > 0:
        kth = torch.topk(z, min(top_k, v), dim=-1).values[..., -1:]
        z = z.masked_fill(z < kth, float("-inf"))
    if top_p < 1.0:
        sz, si = torch.sort(z, dim=-1, descending=True)
        cp = torch.softmax(sz, dim=-1).cumsum(-1)
        drop = cp - torch.softmax(sz, dim=-1) >= top_p
        sz = sz.masked_fill(drop, float("-inf"))
        z = torch.full_like(z, float("-inf")).scatter(-1, si, sz)
    logp = torch.log_softmax(z, dim=-1)
    if is_greedy:
        tok = z.argmax(-1)
    else:
        gen = None
        if seeds is not None:
            gen = torch.Generator(device=z.device)
            gen.manual_seed(int(seeds.reshape(-1)[0].item()) & 0x7FFFFFFF)
        tok = torch.multinomial(logp.exp(), 1, generator=gen).squeeze(-1)
        tok = torch.where(row_greedy, z.argmax(-1), tok)
    return tok, logp.gather(-1, tok.unsqueeze(-1)).squeeze(-1)
