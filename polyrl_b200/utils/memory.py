"""GPU memory utilities (reference: 3rdparty/verl/verl/utils/memory_utils.py:31 ``aggressive_empty_cache``,
verl/utils/device.py:89 ``set_expandable_segments``, the memory-snapshot sampler at memory_utils.py:240 and
``log_gpu_memory_usage`` in verl/utils/debug)."""
from __future__ import annotations

import gc
import os
import threading
import time
from pathlib import Path

import torch


def gpu_memory_stats(device=None) -> dict[str, float]:
    if not torch.cuda.is_available():
        return {"allocated_gb": 0.0, "reserved_gb": 0.0, "free_gb": 0.0, "total_gb": 0.0}
    free, total = torch.cuda.mem_get_info(device)
    g = float(1 << 30)
    return {"allocated_gb": torch.cuda.memory_allocated(device) / g, "reserved_gb": torch.cuda.memory_reserved(device) / g,
            "max_allocated_gb": torch.cuda.max_memory_allocated(device) / g, "free_gb": free / g, "total_gb": total / g}


def log_gpu_memory_usage(head: str, logger=None, rank: int | None = 0, device=None) -> dict[str, float]:
    st = gpu_memory_stats(device)
    cur = int(os.environ.get("RANK", "0"))
    if rank is None or cur == rank:
        msg = f"{head}: " + ", ".join(f"{k}={v:.2f}" for k, v in st.items())
        (logger.info if logger is not None else print)(msg)
    return st


def aggressive_empty_cache(force_sync: bool = True, max_retries: int = 3, device=None) -> dict[str, float]:
    """gc + empty_cache until the reserved pool stops shrinking; returns the memory stats afterwards."""
    if not torch.cuda.is_available():
        gc.collect()
        return gpu_memory_stats()
    last = None
    for _ in range(max_retries):
        gc.collect()
        if force_sync:
            torch.cuda.synchronize(device)
        torch.cuda.empty_cache()
        cur = torch.cuda.memory_reserved(device)
        if last is not None and last - cur < (1 << 20):
            break
        last = cur
    return gpu_memory_stats(device)


def set_expandable_segments(enable: bool) -> None:
    """Toggle the caching allocator's expandable segments (less fragmentation when the KV pool and the trainer's
    activations alternate on the same GPU)."""
    if torch.cuda.is_available():
        torch.cuda.memory._set_allocator_settings(f"expandable_segments:{bool(enable)}")


class MemorySnapshotSampler:
    """Background sampler: records allocator stats every ``interval_s`` and can dump a torch memory snapshot
    (``torch.cuda.memory._dump_snapshot``) for offline visualisation."""

    def __init__(self, out_dir: str, interval_s: float = 1.0, record_history: bool = False, device=None):
        self.out_dir, self.interval_s, self.device = Path(out_dir), interval_s, device
        self.record_history = record_history and torch.cuda.is_available()
        self.samples: list[dict] = []
        self._stop = threading.Event()
        self._thread: threading.Thread | None = None

    def start(self) -> "MemorySnapshotSampler":
        self.out_dir.mkdir(parents=True, exist_ok=True)
        if self.record_history:
            torch.cuda.memory._record_memory_history(max_entries=100000)
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def _run(self) -> None:
        while not self._stop.is_set():
            self.samples.append({"t": time.time(), **gpu_memory_stats(self.device)})
            self._stop.wait(self.interval_s)

    def dump(self, tag: str = "snapshot") -> Path:
        import json

        p = self.out_dir / f"{tag}_rank{os.environ.get('RANK', '0')}.json"
        p.write_text(json.dumps(self.samples))
        if self.record_history:
            torch.cuda.memory._dump_snapshot(str(self.out_dir / f"{tag}_rank{os.environ.get('RANK', '0')}.pickle"))
        return p

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=5)
        if self.record_history:
            torch.cuda.memory._record_memory_history(enabled=None)
