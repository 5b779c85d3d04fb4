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
    """Per-rank profiler of one worker role (reference: verl/utils/profiler/profile.py:26-371, nvtx_profile.py:85-200, config.py).

    ``global_cfg`` (``global_profiler.*``) picks the tool and where traces go; the optional ``role_cfg`` (a role's ``profiler.*``:
    ``enable``, ``all_ranks``, ``ranks``, ``tool_config``) decides whether THIS rank of THIS role records at all — e.g. profile
    rank 0 of the actor and nothing of the critic.  Without a role config every rank records (the trainer decides the steps).

    tools
      * ``nsys``  — ``cudaProfilerStart/Stop`` for runs under ``nsys --capture-range=cudaProfilerApi``.  ``tool_config.nsys.discrete``
        = True turns every ``@DistProfiler.annotate``-d worker method into its OWN capture range (one report per phase: generate /
        log-prob / update ...) instead of one range per training step.
      * ``torch`` — chrome traces; ``tool_config.torch.step_start / step_end`` restrict recording to a window of the micro-steps
        reported through ``step()`` (the actor calls it once per micro-batch), so a long update produces a small trace.
      * ``torch_memory`` — allocator history snapshot (``trace_alloc_max_entries``, ``stack_depth``)."""

    def __init__(self, rank: int, global_cfg, role_cfg=None, role: str = "worker"):
        self.rank, self.role = rank, role
        self.tool = getattr(global_cfg, "tool", None)
        self.save_path = getattr(global_cfg, "save_path", "outputs/profile")
        tool_cfg = dict(getattr(global_cfg, "global_tool_config", None) or {})
        if role_cfg is not None:
            for k, v in (getattr(role_cfg, "tool_config", None) or {}).items():
                tool_cfg[k] = {**(tool_cfg.get(k) or {}), **(v or {})}
            self.this_rank = bool(role_cfg.enable) and (bool(role_cfg.all_ranks) or rank in [int(r) for r in (role_cfg.ranks or [])])
        else:
            self.this_rank = True
        self.tool_cfg = tool_cfg
        self.discrete = bool((tool_cfg.get("nsys") or {}).get("discrete", False))
        tc = tool_cfg.get("torch") or {}
        self.step_start, self.step_end = int(tc.get("step_start", -1)), int(tc.get("step_end", -1))
        self._prof = None
        self._active = False
        self._micro_step = 0
        self.ranges_opened = 0                 # discrete nsys ranges / torch traces produced (tests, logs)

    # ---- helpers
    def _enabled(self) -> bool:
        return bool(self.tool) and self.this_rank

    def _torch_begin(self) -> None:
        acts = [torch.profiler.ProfilerActivity.CPU]
        if torch.cuda.is_available():
            acts.append(torch.profiler.ProfilerActivity.CUDA)
        self._prof = torch.profiler.profile(activities=acts, record_shapes=True)
        self._prof.__enter__()

    def _torch_end(self) -> None:
        if self._prof is None:
            return
        self._prof.__exit__(None, None, None)
        Path(self.save_path).mkdir(parents=True, exist_ok=True)
        self._prof.export_chrome_trace(str(Path(self.save_path) / f"trace_{self.role}_rank{self.rank}_{int(time.time() * 1e3)}.json"))
        self._prof = None
        self.ranges_opened += 1

    # ---- step-level control (called by the trainer around the profiled steps)
    def start(self, **kw) -> None:
        if not self._enabled():
            return
        self._active = True
        self._micro_step = 0
        if self.tool == "nsys":
            if not self.discrete and torch.cuda.is_available():
                torch.cuda.profiler.start()
        elif self.tool == "torch":
            if self.step_start < 0:            # no window: record the whole step
                self._torch_begin()
        elif self.tool == "torch_memory" and torch.cuda.is_available():
            mc = self.tool_cfg.get("torch_memory") or {}
            torch.cuda.memory._record_memory_history(max_entries=int(mc.get("trace_alloc_max_entries", 100000)),
                                                     stacks="all" if int(mc.get("stack_depth", 32)) > 0 else "python")

    def step(self) -> None:
        """One micro-step of the profiled phase has finished (torch tool with a step window)."""
        if not (self._active and self.tool == "torch" and self.step_start >= 0):
            return
        if self._micro_step == self.step_start and self._prof is None:
            self._torch_begin()
        self._micro_step += 1
        if self._prof is not None and self.step_end >= 0 and self._micro_step >= self.step_end:
            self._torch_end()

    def stop(self) -> None:
        if not (self._enabled() and self._active):
            return
        self._active = False
        if self.tool == "nsys":
            if not self.discrete and torch.cuda.is_available():
                torch.cuda.profiler.stop()
        elif self.tool == "torch":
            self._torch_end()
        elif self.tool == "torch_memory" and torch.cuda.is_available():
            dump_memory_snapshot(self.save_path, "profile", None, self.rank)
            torch.cuda.memory._record_memory_history(enabled=None)

    # ---- phase-level ranges
    @contextlib.contextmanager
    def range(self, message: str):
        """NVTX range; in discrete nsys mode (and while a profiled step is active) also its own capture range."""
        cuda = torch.cuda.is_available()
        own_capture = self._active and self._enabled() and self.tool == "nsys" and self.discrete
        if own_capture:
            self.ranges_opened += 1
            if cuda:
                torch.cuda.profiler.start()
        if cuda:
            torch.cuda.nvtx.range_push(f"{self.role}.{message}")
        try:
            yield
        finally:
            if cuda:
                torch.cuda.nvtx.range_pop()
            if own_capture and cuda:
                torch.cuda.profiler.stop()

    @staticmethod
    def annotate(message: str | None = None):
        """Decorator for worker methods: wraps the call in ``self.profiler.range(...)`` when the worker has a profiler."""

        def deco(fn):
            label = message or fn.__name__

            @functools.wraps(fn)
            def wrapped(self, *a, **kw):
                prof = getattr(self, "profiler", None)
                if prof is None:
                    return fn(self, *a, **kw)
                with prof.range(label):
                    return fn(self, *a, **kw)

            return wrapped

        return deco


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
