"""Small Python helpers the trainers, reward functions and user extensions lean on: a wall-clock limit for arbitrary callables, a
scoped environment override, dict merging with conflict detection, metric accumulation, attribute-style views of nested dicts, enums
that are extended at run time.  Capability parity: 3rdparty/verl/verl/utils/py_functional.py (``timeout_limit``, ``temp_env_var``,
``union_two_dict``, ``append_to_dict``, ``NestedNamespace``, ``DynamicEnum``, ``convert_to_regular_types``)."""
from __future__ import annotations

import contextlib
import functools
import multiprocessing as mp
import os
import signal
import threading
from types import SimpleNamespace
from typing import Any, Callable


def _run_in_child(conn, fn, args, kwargs):
    try:
        conn.send((True, fn(*args, **kwargs)))
    except BaseException as e:  # noqa: BLE001 — everything is reported to the parent
        conn.send((False, e))
    finally:
        conn.close()


def timeout_limit(seconds: float, use_signals: bool = False) -> Callable:
    """Decorator: raise ``TimeoutError`` when the call runs longer than ``seconds``.

    ``use_signals=True`` arms ``SIGALRM`` (cheap, in-process; only legal in the main thread, silently falls back otherwise).  The default
    runs the call in a forked child and kills it at the deadline, which also stops C extensions that never return to the interpreter
    (maths graders, sympy simplification) — arguments and result must be picklable."""

    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            if use_signals and threading.current_thread() is threading.main_thread() and hasattr(signal, "SIGALRM"):
                def on_alarm(signum, frame):
                    raise TimeoutError(f"{fn.__name__} timed out after {seconds} s")

                old = signal.signal(signal.SIGALRM, on_alarm)
                signal.setitimer(signal.ITIMER_REAL, seconds)
                try:
                    return fn(*args, **kwargs)
                finally:
                    signal.setitimer(signal.ITIMER_REAL, 0)
                    signal.signal(signal.SIGALRM, old)
            ctx = mp.get_context("fork" if "fork" in mp.get_all_start_methods() else "spawn")
            parent, child = ctx.Pipe(duplex=False)
            proc = ctx.Process(target=_run_in_child, args=(child, fn, args, kwargs), daemon=True)
            proc.start()
            child.close()
            try:
                if not parent.poll(seconds):
                    raise TimeoutError(f"{fn.__name__} timed out after {seconds} s")
                ok, val = parent.recv()
            except EOFError:
                raise RuntimeError(f"{fn.__name__} died without a result (exit code {proc.exitcode})") from None
            finally:
                if proc.is_alive():
                    proc.kill()
                proc.join(1.0)
                parent.close()
            if ok:
                return val
            raise val
        return wrapped
    return deco


@contextlib.contextmanager
def temp_env_var(key: str, value: str | None):
    """Set (or, with ``None``, unset) an environment variable for the duration of the block; the previous state comes back even when
    the block raises."""
    had, old = key in os.environ, os.environ.get(key)
    if value is None:
        os.environ.pop(key, None)
    else:
        os.environ[key] = str(value)
    try:
        yield
    finally:
        if had:
            os.environ[key] = old
        else:
            os.environ.pop(key, None)


def union_two_dict(a: dict, b: dict) -> dict:
    """a | b where a key present in both must carry the same value (metrics / meta-info merging must never silently overwrite)."""
    out = dict(a)
    for k, v in b.items():
        if k in out and out[k] is not v:
            same = out[k] == v
            if not (bool(same) if isinstance(same, bool) else bool(getattr(same, "all", lambda: same)())):
                raise ValueError(f"key {k!r} has different values in the two dicts")
        out[k] = v
    return out


def append_to_dict(acc: dict, new: dict, prefix: str = "") -> dict:
    """acc[prefix + k].append(v) for every item of ``new`` (per-micro-batch metrics -> lists)."""
    for k, v in new.items():
        acc.setdefault(prefix + k, []).append(v)
    return acc


class NestedNamespace(SimpleNamespace):
    """Attribute access over nested dicts (``ns.actor.optim.lr``); lists of dicts become lists of namespaces."""

    def __init__(self, d: dict | None = None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            setattr(self, k, self._wrap(v))

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def to_dict(self) -> dict:
        return convert_to_regular_types(self)


def convert_to_regular_types(x: Any) -> Any:
    """Namespaces / mapping-like config nodes / tuples -> plain dicts and lists (what ``json.dumps`` and loggers accept)."""
    if isinstance(x, SimpleNamespace):
        return {k: convert_to_regular_types(v) for k, v in vars(x).items()}
    if hasattr(x, "items") and callable(x.items):
        return {k: convert_to_regular_types(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)) or (hasattr(x, "__iter__") and type(x).__name__ == "ListConfig"):
        return [convert_to_regular_types(v) for v in x]
    return x


class _DynamicEnumMeta(type):
    def __iter__(cls):
        return iter(cls._members.values())

    def __contains__(cls, item):
        return item in cls._members or item in cls._members.values()

    def __getitem__(cls, name):
        return cls._members[name]

    def __len__(cls):
        return len(cls._members)


class DynamicEnum(metaclass=_DynamicEnumMeta):
    """An enum whose members are registered at run time (``Sub.register("NAME")``): plug-ins add advantage estimators / loss modes
    without editing the core.  Subclass it once per enum; members are singletons with ``name`` and ``value``."""
    _members: dict = {}
    _next = 0

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        cls._members, cls._next = {}, 0

    def __init__(self, name: str, value: int):
        self.name, self.value = name, value

    def __repr__(self):
        return f"<{type(self).__name__}.{self.name}: {self.value}>"

    @classmethod
    def register(cls, name: str):
        key = name.upper()
        if key in cls._members:
            raise ValueError(f"{cls.__name__}.{key} is already registered")
        m = cls(key, cls._next)
        cls._members[key] = m
        setattr(cls, key, m)
        cls._next += 1
        return m

    @classmethod
    def remove(cls, name: str):
        m = cls._members.pop(name.upper())
        delattr(cls, m.name)
        return m

    @classmethod
    def from_name(cls, name: str):
        return cls._members.get(name.upper())
