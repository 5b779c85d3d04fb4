"""Merge per-rank ZeRO-3 checkpoint shards into a single state dict (ours or HuggingFace naming):
    python -m polyrl_b200.utils.model_merger --ckpt checkpoints/global_step_10/actor --out merged/ [--hf]
Reference capability: ``python -m verl.model_merger`` (3rdparty/verl/verl/model_merger/)."""
from __future__ import annotations

import argparse
import math
import re
from pathlib import Path

import torch

from ..models import transformer as tf
from ..models.config import ModelConfig


def merge(ckpt_dir: str) -> tuple[ModelConfig, dict[str, torch.Tensor]]:
    d = Path(ckpt_dir)
    cfg = ModelConfig.load(d / "config.json")
    files = sorted(d.glob("model_world_size_*_rank_*.pt"), key=lambda p: int(re.search(r"rank_(\d+)", p.name).group(1)))
    if not files:
        raise FileNotFoundError(f"no model shards in {d}")
    world = int(re.search(r"world_size_(\d+)", files[0].name).group(1))
    if len(files) != world:
        raise ValueError(f"expected {world} shards, found {len(files)}")
    shards = [torch.load(f, map_location="cpu", weights_only=False) for f in files]
    units = [("embed", tf.embed_param_specs(cfg))] + [(f"layers.{i}", tf.layer_param_specs(cfg, i)) for i in range(cfg.num_layers)] + \
            [("head", tf.head_param_specs(cfg))]
    if cfg.vision is not None and "vision" in shards[0]:          # vision-language checkpoints carry the tower as one more unit
        from ..models.vision import vision_unit_specs

        units.append(("vision", vision_unit_specs(cfg)))
    out: dict[str, torch.Tensor] = {}
    for uname, specs in units:
        so = 0
        for pname, shp in specs:
            numel = math.prod(shp)
            chunk = ((numel + world - 1) // world + 7) // 8 * 8
            full = torch.cat([s[uname]["shard"][so:so + chunk] for s in shards])[:numel].view(shp)
            out[pname if uname in ("embed", "head") else f"{uname}.{pname}"] = full
            so += chunk
    return cfg, out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--hf", action="store_true", help="export HuggingFace parameter names into pytorch_model.bin")
    ap.add_argument("--hf-dir", action="store_true", help="write a complete HuggingFace directory (config.json + sharded safetensors; "
                    "Llama / Qwen2 / Qwen3 / Qwen3-MoE / Qwen2-VL / Qwen2.5-VL / GPT-2) that from_pretrained can load")
    a = ap.parse_args()
    cfg, sd = merge(a.ckpt)
    if a.hf_dir:
        from ..models.hf_io import save_hf

        names = save_hf(sd, cfg, a.out)
        print(f"merged {len(sd)} tensors -> {a.out} ({len(names)} safetensors shard(s))")
        return
    if a.hf:
        if cfg.vision is not None:
            from ..models.vision import export_hf_vision_state_dict

            text = {("model.language_model." + k[6:] if k.startswith("model.") else k): v
                    for k, v in tf.export_hf_state_dict(cfg, {k: v for k, v in sd.items() if not k.startswith("vision.")}).items()}
            sd = {**text, **export_hf_vision_state_dict(cfg, sd)}
        else:
            sd = tf.export_hf_state_dict(cfg, sd)
    Path(a.out).mkdir(parents=True, exist_ok=True)
    torch.save({k: v.contiguous() for k, v in sd.items()}, Path(a.out) / "pytorch_model.bin")
    cfg.save(Path(a.out) / "config.json")
    print(f"merged {len(sd)} tensors -> {a.out}")


if __name__ == "__main__":
    main()
