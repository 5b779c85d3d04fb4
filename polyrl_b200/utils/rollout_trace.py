"""Per-sample rollout tracing: record the inputs / outputs / latency of rollout-side operations (generation turns, tool calls,
reward calls) as JSONL spans grouped by (step, sample index, rollout n).

Reference capability: 3rdparty/verl/verl/utils/rollout_trace.py:23-237 (``RolloutTraceConfig`` singleton with weave / mlflow
back-ends, ``rollout_trace_attr`` context, ``rollout_trace_op`` decorator for sync and async methods, ``token2text``).
Back-ends: ``file`` (local JSONL, needs nothing), ``weave`` (one call per traced op under the attribute scope) and ``mlflow``
(one span per traced op nested in a span per sample, attributes as trace tags); the two services import their client when selected.
With ``token2text`` a result that carries ``prompt_ids`` / ``response_ids`` is recorded with the decoded texts next to the ids
(the traced method's object must have a ``tokenizer``).
"""
from __future__ import annotations

import contextlib
import contextvars
import functools
import inspect
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
        self.client = None                      # weave client / the mlflow module
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
        elif backend == "weave":               # fails loudly when the tracer is requested but not installed
            import weave

            c.client = weave.init(project_name)
        elif backend == "mlflow":
            import mlflow

            if hasattr(getattr(mlflow, "config", None), "enable_async_logging"):
                mlflow.config.enable_async_logging()
            mlflow.set_tracking_uri(os.environ.get("MLFLOW_TRACKING_URI", "sqlite:////tmp/mlruns.db"))
            mlflow.set_experiment(project_name)
            c.client = mlflow
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
    cfg = RolloutTraceConfig.get_instance()
    attrs = {"step": step, "sample_index": sample_index, "rollout_n": rollout_n, "validate": validate, "trace": name}
    tok = _attrs.set({**_attrs.get(), **attrs})
    try:
        tags = {k: v for k, v in attrs.items() if v is not None and k != "trace"}
        tags["experiment_name"] = cfg.experiment
        if cfg.backend == "weave":
            import weave

            with weave.attributes(tags):
                yield
        elif cfg.backend == "mlflow":
            with cfg.client.start_span(name=name) as span:
                for k, v in tags.items():
                    cfg.client.set_trace_tag(span.trace_id, str(k), str(v))
                yield
        else:
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


def _named_inputs(fn, a, kw) -> dict:
    """Arguments by parameter name (``self`` dropped), falling back to positions for callables without a usable signature."""
    try:
        b = inspect.signature(fn).bind(*a, **kw)
        b.apply_defaults()
        return {k: v for k, v in b.arguments.items() if k != "self" and not callable(v)}
    except (TypeError, ValueError):
        return {"args": [x for x in a[1:] if not callable(x)], **kw}


def _with_text(owner, result):
    """``token2text``: decoded prompt / response next to the ids of a result object (or dict) that carries them."""
    tok = getattr(owner, "tokenizer", None)
    get = (lambda k: result.get(k)) if isinstance(result, dict) else (lambda k: getattr(result, k, None))
    if tok is None or not hasattr(tok, "decode") or (get("prompt_ids") is None and get("response_ids") is None):
        return result
    out = dict(result) if isinstance(result, dict) else dict(vars(result))
    for ids_key, text_key in (("prompt_ids", "prompt_text"), ("response_ids", "response_text")):
        if out.get(ids_key) is not None:
            out[text_key] = tok.decode(out[ids_key])
    return out


class _Span:
    """One traced call on the configured backend: ``begin`` before the call, ``end`` with its result or exception."""

    def __init__(self, cfg: RolloutTraceConfig, fn, a, kw):
        self.cfg, self.fn, self.owner = cfg, fn, (a[0] if a else None)
        self.inputs = _named_inputs(fn, a, kw)
        self.t0 = time.time()
        self.call = self.ctx = self.span = None
        if cfg.backend == "weave":
            from weave.trace.context import call_context

            self.call = cfg.client.create_call(op=fn.__qualname__, inputs=self.inputs, attributes={**call_context.call_attributes.get()})
        elif cfg.backend == "mlflow":
            self.ctx = cfg.client.start_span(name=fn.__qualname__)
            self.span = self.ctx.__enter__()
            self.span.set_inputs(self.inputs)

    def end(self, out=None, exc: BaseException | None = None) -> None:
        cfg = self.cfg
        shown = _with_text(self.owner, out) if (cfg.token2text and exc is None) else out
        if cfg.backend == "weave":
            if exc is not None:
                cfg.client.finish_call(self.call, exception=exc)
            else:
                cfg.client.finish_call(self.call, output=shown)
        elif cfg.backend == "mlflow":
            if exc is None:
                self.span.set_outputs(shown)
            self.ctx.__exit__(type(exc) if exc else None, exc, exc.__traceback__ if exc else None)
        else:
            cfg.emit({"op": self.fn.__qualname__, "t_start": self.t0, "dur_s": time.time() - self.t0, **_attrs.get(),
                      "inputs": {k: _short(v) for k, v in self.inputs.items()}, "output": _short(shown), "error": repr(exc) if exc else None})


def rollout_trace_op(fn):
    """Decorator for sync and async methods: one span per call (named arguments, result, wall time) when tracing is enabled; a plain
    call otherwise."""
    if inspect.iscoroutinefunction(fn):
        @functools.wraps(fn)
        async def awrapped(*a, **kw):
            cfg = RolloutTraceConfig.get_instance()
            if cfg.backend is None:
                return await fn(*a, **kw)
            span = _Span(cfg, fn, a, kw)
            try:
                out = await fn(*a, **kw)
            except Exception as e:
                span.end(exc=e)
                raise
            span.end(out)
            return out
        return awrapped

    @functools.wraps(fn)
    def wrapped(*a, **kw):
        cfg = RolloutTraceConfig.get_instance()
        if cfg.backend is None:
            return fn(*a, **kw)
        span = _Span(cfg, fn, a, kw)
        try:
            out = fn(*a, **kw)
        except Exception as e:
            span.end(exc=e)
            raise
        span.end(out)
        return out
    return wrapped
