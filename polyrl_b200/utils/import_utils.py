"""Optional-dependency probes and user-extension loading.  Capability parity: 3rdparty/verl/verl/utils/import_utils.py
(``is_*_available``, ``import_external_libs`` for ``model.external_lib``, ``load_extern_type`` for custom reward / dataset / agent-loop
classes named in the config, ``deprecated``)."""
from __future__ import annotations

import functools
import importlib
import importlib.util
import sys
import warnings
from functools import cache


@cache
def is_available(module: str) -> bool:
    try:
        return importlib.util.find_spec(module) is not None
    except (ImportError, ValueError):
        return False


def is_vllm_available() -> bool:
    return is_available("vllm")


def is_sglang_available() -> bool:
    return is_available("sglang")


def is_megatron_core_available() -> bool:
    return is_available("megatron.core") if is_available("megatron") else False


def is_nvtx_available() -> bool:
    return is_available("nvtx")


def is_trl_available() -> bool:
    return is_available("trl")


def import_external_libs(libs) -> list:
    """Import the module(s) named by ``model.external_lib`` (a name, a comma separated string or a list) before the model is built, so
    user code can register architectures, reward functions or tools as an import side effect."""
    if not libs:
        return []
    names = [x.strip() for x in libs.split(",")] if isinstance(libs, str) else list(libs)
    return [importlib.import_module(n) for n in names if n]


def load_extern_type(path: str | None, name: str, default_module=None):
    """Attribute ``name`` of the python FILE ``path`` (or of ``pkg://dotted.module``; or of ``default_module`` when ``path`` is empty):
    how custom datasets, samplers, reward functions and agent loops named in the config are resolved."""
    if not path:
        if default_module is None:
            raise ValueError(f"no file given for {name!r} and no default module")
        mod = default_module
    elif path.startswith("pkg://"):
        mod = importlib.import_module(path[len("pkg://"):])
    else:
        key = f"polyrl_extern_{abs(hash(path))}"
        mod = sys.modules.get(key)
        if mod is None:
            spec = importlib.util.spec_from_file_location(key, path)
            if spec is None or spec.loader is None:
                raise FileNotFoundError(f"cannot load python file {path!r}")
            mod = importlib.util.module_from_spec(spec)
            sys.modules[key] = mod
            try:
                spec.loader.exec_module(mod)
            except BaseException:
                sys.modules.pop(key, None)
                raise
    if not hasattr(mod, name):
        raise AttributeError(f"{name!r} not found in {path or getattr(mod, '__name__', mod)!r}")
    return getattr(mod, name)


def deprecated(replacement: str = ""):
    """Decorator for functions / classes kept for compatibility: one ``FutureWarning`` per call site naming the replacement."""

    def deco(obj):
        msg = f"{getattr(obj, '__qualname__', obj)} is deprecated" + (f"; use {replacement}" if replacement else "")
        if isinstance(obj, type):
            init = obj.__init__

            @functools.wraps(init)
            def new_init(self, *a, **k):
                warnings.warn(msg, FutureWarning, stacklevel=2)
                init(self, *a, **k)

            obj.__init__ = new_init
            return obj

        @functools.wraps(obj)
        def wrapped(*a, **k):
            warnings.warn(msg, FutureWarning, stacklevel=2)
            return obj(*a, **k)
        return wrapped
    return deco
