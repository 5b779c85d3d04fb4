"""Profiling hooks: torch.profiler / CUDA-profiler-API (nsys capture-range) / CUDA memory-history
snapshots, NVTX-annotated timers.  Capability parity with verl's profiler package (SURVEY §5.1:
3rdparty/verl/verl/utils/profiler/{profile.py:26-371,nvtx_profile.py:85-200,performance.py:155-195})."""
from __future__ import annotations

import contextlib
import functools
import os
import time
from pathlib import Path

import torch


@contextlib.contextmanager
def marked_timer(name: str, timing_raw: dict, color: str | None = None):
    """Accumulate wall seconds into ``timing_raw[name]`` and emit an NVTX range when CUDA is present."""
    nvtx = torch.cuda.is_available()
    if nvtx:
        torch.cuda.nvtx.range_push(name)
    t0 = time.perf_counter()
    try:
        yield
    finally:
        timing_raw[name] = timing_raw.get(name, 0.0) + time.perf_counter() - t0
        if nvtx:
            torch.cuda.nvtx.range_pop()


simple_timer = marked_timer


def annotate(name: str | None = None):
    def deco(fn):
        label = name or fn.__qualname__

        @functools.wraps(fn)
        def wrapped(*a, **kw):
            if torch.cuda.is_available():
                with torch.cuda.nvtx.range(label):
                    return fn(*a, **kw)
            return fn(*a, **kw)

        return wrapped

    return deco


class DistProfiler:
    """tool: 'nsys' (cudaProfilerStart/Stop, to be run under `nsys --capture-range=cudaProfilerApi`),
    'torch' (torch.profiler chrome traces), 'torch_memory' (allocator history snapshot)."""

    def __init__(self, rank: int, cfg):
        self.rank = rank
        self.tool = getattr(cfg, "tool", None)
        self.save_path = getattr(cfg, "save_path", "outputs/profile")
        self._prof = None

    def start(self, **kw) -> None:
        if self.tool == "nsys" and torch.cuda.is_available():
            torch.cuda.profiler.start()
        elif self.tool == "torch":
            acts = [torch.profiler.ProfilerActivity.CPU]
            if torch.cuda.is_available():
                acts.append(torch.profiler.ProfilerActivity.CUDA)
            self._prof = torch.profiler.profile(activities=acts, record_shapes=True)
            self._prof.__enter__()
        elif self.tool == "torch_memory" and torch.cuda.is_available():
            torch.cuda.memory._record_memory_history(max_entries=100000)

    def stop(self) -> None:
        if self.tool == "nsys" and torch.cuda.is_available():
            torch.cuda.profiler.stop()
        elif self.tool == "torch" and self._prof is not None:
            self._prof.__exit__(None, None, None)
            Path(self.save_path).mkdir(parents=True, exist_ok=True)
            self._prof.export_chrome_trace(str(Path(self.save_path) / f"trace_rank{self.rank}_{int(time.time())}.json"))
            self._prof = None
        elif self.tool == "torch_memory" and torch.cuda.is_available():
            dump_memory_snapshot(self.save_path, "profile", None, self.rank)
            torch.cuda.memory._record_memory_history(enabled=None)


def dump_memory_snapshot(save_path: str, tag: str, sub_dir: str | None, rank: int) -> str | None:
    if not torch.cuda.is_available():
        return None
    d = Path(save_path) / (sub_dir or "")
    d.mkdir(parents=True, exist_ok=True)
    f = d / f"memory_snapshot_{tag}_rank{rank}.pickle"
    torch.cuda.memory._dump_snapshot(str(f))
    return str(f)


def log_gpu_memory_usage(head: str, rank: int = 0) -> str:
    if not torch.cuda.is_available():
        return f"{head}: no CUDA"
    a, r = torch.cuda.memory_allocated() / 2 ** 30, torch.cuda.memory_reserved() / 2 ** 30
    free, total = torch.cuda.mem_get_info()
    msg = f"{head}: allocated {a:.2f} GB, reserved {r:.2f} GB, free {free / 2**30:.2f}/{total / 2**30:.2f} GB"
    if int(os.environ.get("RANK", 0)) == rank:
        print(msg, flush=True)
    return msg
