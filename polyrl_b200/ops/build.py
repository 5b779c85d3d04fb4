"""In-tree build of the sm_100a extension ``polyrl_b200/_C.so``.

Kernel ``.cu`` files are torch-free and compiled directly with ``nvcc -gencode arch=compute_100a,
code=sm_100a -lineinfo``; only ``bindings.cpp`` sees torch headers.  The resulting ``.so`` lives in the
package directory (git-ignored, but shipped to the GPU box by ``gpurun``), never in a JIT cache.

Usage:  python -m polyrl_b200.ops.build [--force] [--verbose]
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG = Path(__file__).resolve().parent.parent
CSRC = PKG / "csrc"
BUILD = CSRC / "build"
OUT = PKG / "_C.so"

CU_SOURCES = [
    "gemm_tcgen05.cu",
    "gemm2.cu",
    "skinny_gemm.cu",
    "decode_gemm.cu",
    "vmm.cu",
    "linear_ce.cu",
    "ag_gemm.cu",
    "elementwise.cu",
    "optim.cu",
    "decode_attn.cu",
    "attn_fwd.cu",
    "attn_bwd.cu",
    "sampling.cu",
    "p2p.cu",
    "nvls.cu",
]
HEADERS = ["common.cuh", "gemm_sm100.cuh", "gemm2_sm100.cuh"]
HOST_HEADERS = ["kv_manager.hpp", "scheduler.hpp"]          # included by bindings.cpp only
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math" if False else "-DPRL_NO_FAST_MATH",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


LAST_BUILD: dict = {}          # what the last build() call did (clean / incremental / up-to-date), printed by __graft_entry__.build()


def _nvcc() -> str:
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    return str(Path(cuda_home) / "bin" / "nvcc")


def _newer(src: Path, dst: Path, extra: list[Path]) -> bool:
    if not dst.exists():
        return True
    t = dst.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, *extra] if p.exists())


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {' '.join(cmd[:4])} ...")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False) -> Path:
    import torch  # noqa: F401  (headers + libs)
    from torch.utils import cpp_extension as ce

    BUILD.mkdir(parents=True, exist_ok=True)
    hdrs = [CSRC / h for h in HEADERS]
    jobs: list[list[str]] = []
    objs: list[Path] = []
    for name in CU_SOURCES:
        src = CSRC / name
        obj = BUILD / (src.stem + ".o")
        objs.append(obj)
        if force or _newer(src, obj, hdrs):
            jobs.append([_nvcc(), *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)])
    bsrc = CSRC / "bindings.cpp"
    bobj = BUILD / "bindings.o"
    objs.append(bobj)
    if force or _newer(bsrc, bobj, [CSRC / h for h in HOST_HEADERS]):
        inc = []
        for p in ce.include_paths("cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(True):
            inc += ["-isystem", p]
        inc += ["-isystem", sysconfig.get_paths()["include"]]
        jobs.append([
            "g++", "-O2", "-std=c++17", "-fPIC", "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", *inc, "-c", str(bsrc), "-o", str(bobj),
        ])
    import time

    t0 = time.time()
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs))
    global LAST_BUILD
    LAST_BUILD = {"build_mode": "clean" if len(jobs) == len(objs) else ("incremental" if jobs else "up-to-date"),
                  "objects_compiled": len(jobs), "objects_total": len(objs), "compile_s": round(time.time() - t0, 1),
                  "arch": "compute_100a/sm_100a", "nvcc_flags": NVCC_FLAGS}
    if jobs or force or not OUT.exists():
        lib_dirs = ce.library_paths("cuda") if "device_type" in ce.library_paths.__code__.co_varnames else ce.library_paths(True)
        tmp_out = OUT.with_suffix(".so.tmp")           # link next to the target, then rename: readers never see a half-written .so
        link = ["g++", "-shared", "-o", str(tmp_out), *map(str, objs)]
        for d in lib_dirs:
            link += [f"-L{d}", f"-Wl,-rpath,{d}"]
        link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
        _run(link, verbose)
        tmp_out.replace(OUT)
    return OUT


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
    import json

    print(json.dumps(LAST_BUILD))


if __name__ == "__main__":
    main()
