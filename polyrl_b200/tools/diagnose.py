"""Environment report for bug reports and cluster bring-up (reference: scripts/diagnose.py): interpreter, torch / CUDA / NCCL
versions, GPUs with their peer-access matrix and NVLS (multicast) support, the state of the in-tree extension and of the C++
manager binary, and the optional dependencies the framework can use.

    python -m polyrl_b200.tools.diagnose [--json]
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import platform
import shutil
import subprocess
import sys


def collect() -> dict:
    info: dict = {"python": sys.version.split()[0], "platform": platform.platform(), "cpus": os.cpu_count()}
    try:
        import torch

        info["torch"] = torch.__version__
        info["cuda_runtime"] = torch.version.cuda
        info["cuda_available"] = torch.cuda.is_available()
        try:
            info["nccl"] = ".".join(map(str, torch.cuda.nccl.version()))
        except Exception:                         # noqa: BLE001
            info["nccl"] = None
        gpus = []
        if torch.cuda.is_available():
            n = torch.cuda.device_count()
            for i in range(n):
                p = torch.cuda.get_device_properties(i)
                gpus.append({"index": i, "name": p.name, "sm": f"{p.major}.{p.minor}", "sms": p.multi_processor_count,
                             "hbm_gb": round(p.total_memory / 2 ** 30, 1)})
            info["peer_access"] = [[i == j or torch.cuda.can_device_access_peer(i, j) for j in range(n)] for i in range(n)]
        info["gpus"] = gpus
    except Exception as e:                        # noqa: BLE001
        info["torch_error"] = repr(e)
    try:
        from ..ops import LAUNCHES, ext  # noqa: F401

        C = ext()
        info["extension"] = {"loaded": True, "path": getattr(C._mod, "__file__", None)}
        if info.get("cuda_available"):
            try:
                info["extension"]["nvls_supported"] = bool(C.nvls_supported(0))
            except Exception as e:                # noqa: BLE001
                info["extension"]["nvls_supported"] = f"error: {e!r}"
    except Exception as e:                        # noqa: BLE001
        info["extension"] = {"loaded": False, "error": str(e)[:300]}
    try:
        from ..manager.build import build_manager

        info["manager_binary"] = str(build_manager())
    except Exception as e:                        # noqa: BLE001
        info["manager_binary"] = f"unavailable: {e!r}"[:300]
    info["tools"] = {t: shutil.which(t) for t in ("nvcc", "g++", "ncu", "nvidia-smi", "compute-sanitizer")}
    if info["tools"]["nvcc"]:
        try:
            info["nvcc"] = subprocess.run(["nvcc", "--version"], capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[-1]
        except Exception:                         # noqa: BLE001
            pass
    opt = {}
    for m in ("transformers", "safetensors", "pyarrow", "yaml", "cv2", "PIL", "tensorboard", "wandb", "mlflow", "ray", "vllm", "flash_attn"):
        try:
            opt[m] = getattr(importlib.import_module(m), "__version__", "ok")
        except Exception:                         # noqa: BLE001
            opt[m] = None
    info["optional_modules"] = opt
    info["env"] = {k: os.environ[k] for k in sorted(os.environ) if k.startswith(("POLYRL_", "NCCL_", "CUDA_VISIBLE", "MASTER_", "WORLD_SIZE", "RANK"))}
    return info


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args(argv)
    info = collect()
    if a.json:
        print(json.dumps(info, indent=1, default=str))
        return 0
    for k, v in info.items():
        if isinstance(v, dict):
            print(f"{k}:")
            for kk, vv in v.items():
                print(f"  {kk}: {vv}")
        elif isinstance(v, list) and v and isinstance(v[0], dict):
            print(f"{k}:")
            for row in v:
                print("  " + ", ".join(f"{a_}={b_}" for a_, b_ in row.items()))
        else:
            print(f"{k}: {v}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
