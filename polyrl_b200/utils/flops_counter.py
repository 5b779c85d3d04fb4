"""Model-FLOPs estimator and MFU (reference: 3rdparty/verl/verl/utils/flops_counter.py:41-396; device peak table
:78-97 with B200 = 2.25e15 dense bf16).  ``estimate_flops(batch_seqlens, delta_time)`` -> (achieved TFLOP/s, peak TFLOP/s)."""
from __future__ import annotations

import json
from pathlib import Path

import torch

from ..models.config import ModelConfig

DEVICE_PEAK_BF16 = {"B200": 2.25e15, "GB200": 2.5e15, "B300": 2.25e15, "H100": 989e12, "H800": 989e12, "H200": 989e12,
                    "A100": 312e12, "A800": 312e12, "L40": 181.05e12, "L20": 119.5e12, "H20": 148e12, "910B": 354e12}


def get_device_flops(unit: str = "T", measured: bool = False) -> float:
    """Nominal dense bf16 peak of the current device; ``measured=True`` prefers MEASURED_PEAKS.json (cuBLAS burst)."""
    div = {"B": 1e9, "K": 1e3, "M": 1e6, "G": 1e9, "T": 1e12, "P": 1e15}[unit]
    if measured:
        p = Path(__file__).resolve().parents[2] / "MEASURED_PEAKS.json"
        if p.exists():
            return json.loads(p.read_text())["bf16_tflops"] * 1e12 / div
    name = torch.cuda.get_device_name() if torch.cuda.is_available() else "CPU"
    for k, v in DEVICE_PEAK_BF16.items():
        if k in name:
            return v / div
    return float("inf")


class FlopsCounter:
    """Dense decoder FLOPs: 6·N_dense·tokens (+ 12·L·H·Σ s² attention) for fwd+bwd; forward-only = 1/3."""

    def __init__(self, cfg: ModelConfig):
        self.cfg = cfg

    def _dense_params(self) -> int:
        c = self.cfg
        h, f, v = c.hidden_size, c.intermediate_size, c.vocab_size
        if c.num_experts:                                 # MoE: only the routed experts (+ the router) do work per token
            mlp = c.num_experts_per_tok * 3 * h * c.moe_intermediate_size + c.num_experts * h
        else:
            mlp = 3 * h * f if c.family == "llama" else 2 * h * f
        per_layer = h * (c.q_size + 2 * c.kv_size) + c.q_size * h + mlp
        return per_layer * c.num_layers + v * h           # lm_head included, input embedding lookup excluded

    def estimate_flops(self, batch_seqlens: list[int], delta_time: float, forward_only: bool = False) -> tuple[float, float]:
        c = self.cfg
        tokens = sum(batch_seqlens)
        dense = 6 * self._dense_params() * tokens
        attn = 12 * c.num_layers * c.q_size * sum(s * s for s in batch_seqlens)
        total = dense + attn
        if forward_only:
            total /= 3
        return total / max(delta_time, 1e-9) / 1e12, get_device_flops("T")
