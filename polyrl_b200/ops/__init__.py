"""Fused ops: sm_100a kernels on CUDA tensors, plain-PyTorch reference on CPU tensors.

The CUDA path never silently falls back: if a CUDA tensor reaches an op and the in-tree extension
``polyrl_b200/_C.so`` is missing or fails to load, we raise.
"""
from __future__ import annotations

import importlib
import os

_EXT = None
_EXT_ERR: Exception | None = None
LAUNCHES: dict[str, int] = {}


class _CountingExt:
    """Thin proxy over the compiled module that counts calls per kernel entry point, so benchmarks can
    report how many of OUR kernels ran (CUDA-graph replays add the count recorded at capture)."""

    def __init__(self, mod):
        self._mod = mod
        self._cache: dict = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._mod, name)
            if not callable(raw) or isinstance(raw, type):       # attributes and host-side classes (KvPageManager) are not launches
                return raw

            def fn(*a, _raw=raw, _n=name, **kw):
                LAUNCHES[_n] = LAUNCHES.get(_n, 0) + 1
                return _raw(*a, **kw)

            self._cache[name] = fn
        return fn


def total_launches() -> int:
    return sum(v for k, v in LAUNCHES.items() if not k.startswith(("sym_", "ipc_", "tensor_from", "enable_peer")))


def ext():
    """Return the compiled extension module, building nothing; raises with a clear message if absent."""
    global _EXT, _EXT_ERR
    if _EXT is None and _EXT_ERR is None:
        try:
            import torch  # noqa: F401  (must be imported before the extension)

            _EXT = _CountingExt(importlib.import_module("polyrl_b200._C"))
        except Exception as e:  # pragma: no cover - depends on build state
            _EXT_ERR = e
    if _EXT is None:
        raise RuntimeError(
            "polyrl_b200._C (sm_100a extension) is not available: run `python -m polyrl_b200.ops.build` "
            f"(or __graft_entry__.build()). Import error: {_EXT_ERR!r}"
        )
    return _EXT


def have_ext() -> bool:
    try:
        ext()
        return True
    except RuntimeError:
        return False


def use_cublas_gemm() -> bool:
    """Baseline switch: route plain GEMMs through cuBLAS (torch.matmul) instead of the tcgen05 kernel."""
    return os.environ.get("POLYRL_GEMM_BACKEND", "tcgen05") == "cublas"


from .functional import (  # noqa: E402
    linear,
    linear_swiglu,
    rmsnorm,
    apply_rope,
    Fp8Weight,
    quantize_weight_fp8,
    linear_cross_entropy,
    linear_cross_entropy_tp,
    paged_decode_attention,
    sample_tokens,
)

__all__ = [
    "ext",
    "LAUNCHES",
    "total_launches",
    "have_ext",
    "linear",
    "linear_swiglu",
    "rmsnorm",
    "apply_rope",
    "Fp8Weight",
    "quantize_weight_fp8",
    "linear_cross_entropy",
    "linear_cross_entropy_tp",
    "paged_decode_attention",
    "sample_tokens",
]
