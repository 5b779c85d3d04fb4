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
        dense_mlp = 3 * h * f if (c.family == "llama" and c.mlp_act == "swiglu") else 2 * h * f      # gated vs plain / xIELU MLP
        attn = c.attn_num_params() - ((c.kv_lora_rank + c.q_lora_rank) if c.attn_type == "mla" else 0)      # projections only, not the latent norm gains
        if c.num_experts:                                 # MoE: only the routed (+ shared) experts and the router do work per token
            sparse_mlp = (c.num_experts_per_tok * 3 * h * c.moe_intermediate_size + c.num_experts * h
                          + 3 * h * c.shared_expert_intermediate_size + (h if c.shared_expert_gate else 0))
            n_sparse = c.num_sparse_layers()
            return attn * c.num_layers + sparse_mlp * n_sparse + dense_mlp * (c.num_layers - n_sparse) + v * h
        return (attn + dense_mlp) * c.num_layers + v * h   # lm_head included, input embedding lookup excluded

    def _attn_flops(self, batch_seqlens: list[int]) -> float:
        """QK^T + PV, fwd+bwd, counted without the causal discount (the reference's convention).  Sliding-window layers
        (``sliding_window`` + ``layer_types`` / ``max_window_layers``) attend to at most ``window`` keys per query."""
        c = self.cfg
        full = sum(s * s for s in batch_seqlens)
        win = getattr(c, "sliding_window", None)
        if c.attn_type == "mla":                # QK^T over nope + rope features, PV over v_head_dim features
            return 12.0 * c.num_layers * c.num_heads * 0.5 * (c.head_dim + c.v_head_dim) * full
        if not win:
            return 12.0 * c.num_layers * c.q_size * full
        windowed = sum(s * min(s, int(win)) for s in batch_seqlens)
        types = getattr(c, "layer_types", None)
        if types:
            n_win = sum(1 for t in types if "sliding" in str(t))
        else:
            mwl = getattr(c, "max_window_layers", None)          # HF Qwen2: layers with index >= max_window_layers slide
            n_win = c.num_layers if mwl is None else max(0, c.num_layers - int(mwl))
        return 12.0 * c.q_size * ((c.num_layers - n_win) * full + n_win * windowed)

    def _vision_flops(self, images_seqlens: list[int]) -> float:
        """ViT tower (Qwen2-VL / 2.5-VL layout, models/vision.py), fwd+bwd: ``images_seqlens`` = patches per image (before the
        spatial merge).  Dense 6·N_vit·patches + 12·depth·dim·Σ p² attention (full attention inside an image) + the merger MLP."""
        v = self.cfg.vision
        if not v or not images_seqlens:
            return 0.0
        dim, depth = int(v["embed_dim"]), int(v["depth"])
        inter = int(v.get("intermediate_size") or dim * float(v.get("mlp_ratio", 4)))
        glm = v.get("arch") == "glm4v"
        if glm:                                   # block MLPs are as wide as the tower output; gated
            inter = self.cfg.hidden_size
        mlp = (3 if v.get("arch") in ("qwen2_5", "glm4v") else 2) * dim * inter
        per_block = 4 * dim * dim + mlp
        patches = sum(images_seqlens)
        in_ch = int(v.get("in_channels", 3)) * int(v.get("temporal_patch_size", 2)) * int(v["patch_size"]) ** 2
        merge = int(v.get("spatial_merge_size", 2)) ** 2
        merger = (dim * merge) * (dim * merge) + (dim * merge) * self.cfg.hidden_size          # per merged token
        if glm:                                   # strided-conv cell merge + projection + gated MLP of width vision.intermediate_size
            o = self.cfg.hidden_size
            merger = (dim * merge) * o + o * o + 3 * o * int(v["intermediate_size"])
        dense = 6.0 * (per_block * depth + in_ch * dim) * patches + 6.0 * merger * (patches / merge)
        return dense + 12.0 * depth * dim * sum(p * p for p in images_seqlens)

    def estimate_flops(self, batch_seqlens: list[int], delta_time: float, forward_only: bool = False,
                       images_seqlens: list[int] | None = None) -> tuple[float, float]:
        """(achieved TFLOP/s, device peak TFLOP/s).  Families: llama / qwen2 / qwen3 dense (gated MLP), gpt2 (plain MLP), MoE
        (routed experts only), sliding-window variants, vision-language models (pass ``images_seqlens``)."""
        tokens = sum(batch_seqlens)
        total = 6.0 * self._dense_params() * tokens + self._attn_flops(batch_seqlens) + self._vision_flops(images_seqlens or [])
        if forward_only:
            total /= 3
        return total / max(delta_time, 1e-9) / 1e12, get_device_flops("T")
