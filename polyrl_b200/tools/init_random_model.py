"""Write a randomly initialised HF-format checkpoint of a preset / config (optionally shrunk), e.g. for dry runs of the trainer and
engine servers without downloading weights (reference: scripts/init_random_model.py).

    python -m polyrl_b200.tools.init_random_model --model llama3-8b --out /tmp/llama-2l --override num_layers=2 hidden_size=1024
"""
from __future__ import annotations

import argparse
import sys

import torch


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model", required=True, help="preset name, config.json or HF directory (architecture only)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--override", nargs="*", default=[], help="ModelConfig fields, e.g. num_layers=2 hidden_size=512")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--dtype", default="bfloat16", choices=["bfloat16", "float32", "float16"])
    a = ap.parse_args(argv)
    from ..models import TransformerWeights, get_model_config
    from ..models.hf_io import save_hf

    ov = {}
    for kv in a.override:
        k, _, v = kv.partition("=")
        ov[k] = (v.lower() == "true") if v.lower() in ("true", "false") else (int(v) if v.lstrip("-").isdigit() else (float(v) if v.replace(".", "", 1).replace("e-", "", 1).isdigit() else v))
    cfg = get_model_config(a.model, **ov)
    w = TransformerWeights(cfg, device="cpu", dtype=torch.float32, init=True, seed=a.seed)
    sd = w.state_dict()
    if cfg.vision is not None:
        from ..models.vision import VisionWeights

        sd.update(VisionWeights(cfg, seed=a.seed + 1).state_dict())
    files = save_hf(sd, cfg, a.out, dtype=getattr(torch, a.dtype))
    print(f"[init_random_model] {cfg.num_params() / 1e6:.1f} M parameters -> {a.out} ({len(files)} file(s))")
    return 0


if __name__ == "__main__":
    sys.exit(main())
