"""Per-sample rollout tracing: record the inputs / outputs / latency of rollout-side operations (generation turns, tool calls,
reward calls) as JSONL spans grouped by (step, sample index, rollout n).

Reference capability: 3rdparty/verl/verl/utils/rollout_trace.py:23-237 (``RolloutTraceConfig`` singleton with weave / mlflow
back-ends, ``rollout_trace_attr`` context, ``rollout_trace_op`` decorator).  Offline image: the back-end here is a local JSONL
file (``backend="file"``); ``weave`` / ``mlflow`` are accepted and import-gated.
"""
from __future__ import annotations

import contextlib
import contextvars
import functools
import json
import os
import threading
import time
from pathlib import Path
from typing import Any

_attrs: contextvars.ContextVar[dict] = contextvars.ContextVar("rollout_trace_attrs", default={})


class RolloutTraceConfig:
    _instance: "RolloutTraceConfig | None" = None

    def __init__(self):
        self.backend: str | None = None
        self.project = self.experiment = ""
        self.token2text = False
        self.path: Path | None = None
        self._lock = threading.Lock()

    @classmethod
    def get_instance(cls) -> "RolloutTraceConfig":
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    @classmethod
    def init(cls, project_name: str, experiment_name: str, backend: str | None, token2text: bool = False, out_dir: str = "outputs/traces"):
        c = cls.get_instance()
        c.backend, c.project, c.experiment, c.token2text = backend, project_name, experiment_name, token2text
        if backend == "file":
            c.path = Path(out_dir) / project_name / f"{experiment_name}.rank{os.environ.get('RANK', '0')}.jsonl"
            c.path.parent.mkdir(parents=True, exist_ok=True)
        elif backend in ("weave", "mlflow"):
            __import__(backend)            # fail loudly when the tracer is requested but not installed
        elif backend is not None:
            raise ValueError(f"unknown rollout trace backend {backend!r}")
        return c

    @classmethod
    def reset(cls) -> None:
        cls._instance = None

    def emit(self, span: dict) -> None:
        if self.backend != "file" or self.path is None:
            return
        with self._lock, open(self.path, "a") as f:
            f.write(json.dumps(span, default=str) + "\n")


@contextlib.contextmanager
def rollout_trace_attr(step: int | None = None, sample_index: int | None = None, rollout_n: int | None = None, validate: bool = False,
                       name: str = "rollout"):
    """Attach (step, sample, n) attributes to every traced op executed inside the block."""
    tok = _attrs.set({**_attrs.get(), "step": step, "sample_index": sample_index, "rollout_n": rollout_n, "validate": validate,
                      "trace": name})
    try:
        yield
    finally:
        _attrs.reset(tok)


def _short(v: Any, limit: int = 2000):
    if isinstance(v, (int, float, bool, type(None))):
        return v
    if isinstance(v, (list, tuple)) and len(v) <= 4096 and all(isinstance(x, (int, float)) for x in v):
        return list(v)
    s = repr(v)
    return s if len(s) <= limit else s[:limit] + "..."


def rollout_trace_op(fn):
    """Decorator: one span per call (arguments, result, wall time) when tracing is enabled; zero overhead otherwise."""
    @functools.wraps(fn)
    def wrapped(*a, **kw):
        cfg = RolloutTraceConfig.get_instance()
        if cfg.backend is None:
            return fn(*a, **kw)
        t0 = time.time()
        err = None
        try:
            out = fn(*a, **kw)
            return out
        except Exception as e:
            err, out = repr(e), None
            raise
        finally:
            cfg.emit({"op": fn.__qualname__, "t_start": t0, "dur_s": time.time() - t0, **_attrs.get(),
                      "inputs": {"args": [_short(x) for x in a[1:] if not callable(x)], "kwargs": {k: _short(v) for k, v in kw.items()}},
                      "output": _short(out), "error": err})
    return wrapped
