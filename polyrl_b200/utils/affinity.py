"""Pin a rank's host threads to the CPUs next to its GPU.

One process per GPU on a two-socket B200 box: the decode loop's launches, the pinned-buffer H2D / D2H copies of every step and the
NCCL proxy threads all run on the host; on the far socket each of them crosses the inter-socket link.  ``set_numa_affinity`` asks NVML
for the GPU's ideal CPU set and applies it to the calling process (threads created afterwards inherit it).  Opt-in:
``POLYRL_NUMA_AFFINITY=1`` (or ``trainer.numa_affinity``-style callers passing ``force=True``); a missing NVML, a container without
the permission, or a CPU mask that would leave no CPU are reported and ignored — it never stops a run.
Capability parity: 3rdparty/verl/verl/utils/distributed.py:26-51 (``set_numa_affinity`` through pynvml)."""
from __future__ import annotations

import os


def gpu_cpu_set(index: int) -> list[int]:
    """CPU ids NVML considers local to GPU ``index`` (the device's ideal affinity mask), empty when NVML cannot tell."""
    import pynvml

    pynvml.nvmlInit()
    try:
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        return [w * 64 + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
    finally:
        pynvml.nvmlShutdown()


def _physical_index(local_rank: int) -> int:
    """NVML enumerates ALL GPUs of the box; CUDA_VISIBLE_DEVICES may have re-ordered / hidden some for this process."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    ids = [v.strip() for v in vis.split(",") if v.strip()]
    if ids and local_rank < len(ids) and ids[local_rank].isdigit():
        return int(ids[local_rank])
    return local_rank


def set_numa_affinity(local_rank: int | None = None, force: bool = False) -> list[int] | None:
    """Bind this process to its GPU's CPU set; returns the set applied, or None when nothing was changed."""
    if not force and os.environ.get("POLYRL_NUMA_AFFINITY", "0") != "1":
        return None
    if not hasattr(os, "sched_setaffinity"):
        return None
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", 0))
    try:
        cpus = gpu_cpu_set(_physical_index(local_rank))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]              # never ask for CPUs the container does not own
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except ImportError:
        print("[affinity] pynvml not available; NUMA affinity not set", flush=True)
    except Exception as e:  # noqa: BLE001 — NVML errors, EPERM: report and carry on
        print(f"[affinity] NUMA affinity not set: {e}", flush=True)
    return None
