"""Precision names <-> torch dtypes, in every spelling configs use ("bf16", "bfloat16", "16-mixed", ``torch.bfloat16`` ...).
Capability parity: 3rdparty/verl/verl/utils/torch_dtypes.py (``PrecisionType``)."""
from __future__ import annotations

import torch

_HALF = ("16", "fp16", "float16", "half", "16-mixed", "16-true", torch.float16)
_BF16 = ("bf16", "bfloat16", "bf16-mixed", "bf16-true", torch.bfloat16)
_FP32 = ("32", "fp32", "float32", "float", "32-true", torch.float32)
_FP64 = ("64", "fp64", "float64", "double", "64-true", torch.float64)


class PrecisionType:
    @staticmethod
    def _norm(p):
        return p if isinstance(p, torch.dtype) else str(p).lower().removeprefix("torch.")

    @classmethod
    def supported(cls, precision) -> bool:
        return cls._norm(precision) in (*_HALF, *_BF16, *_FP32, *_FP64)

    @classmethod
    def is_fp16(cls, precision) -> bool:
        return cls._norm(precision) in _HALF

    @classmethod
    def is_bf16(cls, precision) -> bool:
        return cls._norm(precision) in _BF16

    @classmethod
    def is_fp32(cls, precision) -> bool:
        return cls._norm(precision) in _FP32

    @classmethod
    def to_dtype(cls, precision) -> torch.dtype:
        p = cls._norm(precision)
        for names, dt in ((_HALF, torch.float16), (_BF16, torch.bfloat16), (_FP32, torch.float32), (_FP64, torch.float64)):
            if p in names:
                return dt
        raise ValueError(f"unknown precision {precision!r}")

    @classmethod
    def to_str(cls, precision) -> str:
        return {torch.float16: "fp16", torch.bfloat16: "bf16", torch.float32: "fp32", torch.float64: "fp64"}[cls.to_dtype(precision)]
