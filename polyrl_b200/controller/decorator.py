"""Single-controller dispatch / collect protocol.

A worker method decorated with ``@register(dispatch_mode=..., execute_mode=..., blocking=...)`` declares how
the driver-side ``WorkerGroup`` splits its arguments across workers and merges their results.
Capability parity: 3rdparty/verl/verl/single_controller/base/decorator.py:39-455 (modes :39-64, auto
padding :88-114, nd-mesh lazy dispatch :211-315, ``register`` :409-455).
"""
from __future__ import annotations

import functools
import inspect
from enum import Enum
from typing import Callable

from ..protocol import DataProto, DataProtoFuture, pad_dataproto_to_divisor, unpad_dataproto

MAGIC_ATTR = "__polyrl_dispatch__"


class Dispatch(str, Enum):
    RANK_ZERO = "rank_zero"
    ONE_TO_ALL = "one_to_all"
    ALL_TO_ALL = "all_to_all"
    DP_COMPUTE = "dp_compute"
    DP_COMPUTE_PROTO = "dp_compute_proto"
    DP_COMPUTE_PROTO_WITH_FUNC = "dp_compute_proto_with_func"
    DP_COMPUTE_METRIC = "dp_compute_metric"
    DIRECT_ROLLOUT_METHOD = "direct_rollout_method"


class Execute(str, Enum):
    ALL = "all"
    RANK_ZERO = "rank_zero"


def _split_args(world: int, args, kwargs, auto_pad: bool = False):
    """Chunk every DataProto / DataProtoFuture argument into ``world`` parts; scalars are replicated."""
    pad = 0

    def chunk(v):
        nonlocal pad
        if isinstance(v, DataProto):
            if auto_pad and len(v) % world != 0:
                v, pad = pad_dataproto_to_divisor(v, world)
            return v.chunk(world)
        if isinstance(v, DataProtoFuture):
            return v.chunk(world)
        return [v] * world

    s_args = [chunk(a) for a in args]
    s_kwargs = {k: chunk(v) for k, v in kwargs.items()}
    per_worker = [([a[i] for a in s_args], {k: v[i] for k, v in s_kwargs.items()}) for i in range(world)]
    return per_worker, pad


def _concat(outputs: list):
    outs = [o for o in outputs if o is not None]
    if not outs:
        return None
    if isinstance(outs[0], DataProto):
        return DataProto.concat(outs)
    if hasattr(outs[0], "result") or isinstance(outs[0], DataProtoFuture):
        return DataProtoFuture.concat(outs)
    return outs


# Each mode: dispatch_fn(worker_group, *args, **kwargs) -> list[(args, kwargs)] per worker,
#            collect_fn(worker_group, outputs) -> merged
def dispatch_one_to_all(wg, *args, **kwargs):
    return [(list(args), dict(kwargs)) for _ in range(wg.world_size)]


def collect_all(wg, outputs):
    return outputs


def dispatch_all_to_all(wg, *args, **kwargs):
    return [(list(args), dict(kwargs)) for _ in range(wg.world_size)]


def dispatch_rank_zero(wg, *args, **kwargs):
    return [(list(args), dict(kwargs))]


def collect_rank_zero(wg, outputs):
    return outputs[0]


def dispatch_dp_compute(wg, *args, **kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, (list, tuple)) and len(a) != wg.world_size:
            raise ValueError("DP_COMPUTE expects per-worker lists of length world_size")
    return [([a[i] for a in args], {k: v[i] for k, v in kwargs.items()}) for i in range(wg.world_size)]


def dispatch_dp_compute_proto(wg, *args, **kwargs):
    per_worker, pad = _split_args(wg.world_size, args, kwargs, auto_pad=True)
    wg._last_pad = pad
    return per_worker


def collect_dp_compute_proto(wg, outputs):
    out = _concat(outputs)
    pad = getattr(wg, "_last_pad", 0)
    if pad and isinstance(out, DataProto):
        out = unpad_dataproto(out, pad)
    wg._last_pad = 0
    return out


def dispatch_dp_compute_proto_with_func(wg, *args, **kwargs):
    func, rest = args[0], args[1:]
    per_worker, pad = _split_args(wg.world_size, rest, kwargs, auto_pad=True)
    wg._last_pad = pad
    return [([func] + a, k) for a, k in per_worker]


def collect_dp_metric(wg, outputs):
    return [o for o in outputs if o is not None]


def make_nd_compute_dataproto_dispatch_fn(mesh_name: str):
    """Lazy nd-mesh dispatch: the first call queries every worker's ``(dp_rank, dp_size, is_collect)`` for
    ``mesh_name`` (registered by the worker via ``_register_dispatch_collect_info``); data is chunked by
    dp_size, replicated inside each model-parallel group, and collected from ``is_collect`` ranks only."""

    def dispatch(wg, *args, **kwargs):
        info = wg.get_mesh_info(mesh_name)
        dp_size = max(i["dp_size"] for i in info)
        per_dp, pad = _split_args(dp_size, args, kwargs, auto_pad=True)
        wg._last_pad = pad
        return [per_dp[i["dp_rank"]] for i in info]

    def collect(wg, outputs):
        info = wg.get_mesh_info(mesh_name)
        keep = sorted(((i["dp_rank"], o) for i, o in zip(info, outputs) if i["is_collect"]), key=lambda t: t[0])
        out = _concat([o for _, o in keep])
        pad = getattr(wg, "_last_pad", 0)
        if pad and isinstance(out, DataProto):
            out = unpad_dataproto(out, pad)
        wg._last_pad = 0
        return out

    return {"dispatch_fn": dispatch, "collect_fn": collect}


_PREDEFINED = {
    Dispatch.ONE_TO_ALL: {"dispatch_fn": dispatch_one_to_all, "collect_fn": collect_all},
    Dispatch.ALL_TO_ALL: {"dispatch_fn": dispatch_all_to_all, "collect_fn": collect_all},
    Dispatch.RANK_ZERO: {"dispatch_fn": dispatch_rank_zero, "collect_fn": collect_rank_zero},
    Dispatch.DP_COMPUTE: {"dispatch_fn": dispatch_dp_compute, "collect_fn": collect_all},
    Dispatch.DP_COMPUTE_PROTO: {"dispatch_fn": dispatch_dp_compute_proto, "collect_fn": collect_dp_compute_proto},
    Dispatch.DP_COMPUTE_PROTO_WITH_FUNC: {"dispatch_fn": dispatch_dp_compute_proto_with_func, "collect_fn": collect_dp_compute_proto},
    Dispatch.DP_COMPUTE_METRIC: {"dispatch_fn": dispatch_dp_compute_proto, "collect_fn": collect_dp_metric},
    Dispatch.DIRECT_ROLLOUT_METHOD: {"dispatch_fn": dispatch_one_to_all, "collect_fn": collect_all},
}


def get_predefined_dispatch_fn(mode) -> dict:
    if isinstance(mode, dict):
        if not {"dispatch_fn", "collect_fn"} <= set(mode):
            raise ValueError("custom dispatch mode needs dispatch_fn and collect_fn")
        return mode
    return _PREDEFINED[Dispatch(mode)]


def _materialize_futures(args, kwargs):
    args = [a.get() if isinstance(a, DataProtoFuture) else a for a in args]
    kwargs = {k: (v.get() if isinstance(v, DataProtoFuture) else v) for k, v in kwargs.items()}
    return args, kwargs


def register(dispatch_mode=Dispatch.ALL_TO_ALL, execute_mode=Execute.ALL, blocking: bool = True,
             materialize_futures: bool = True) -> Callable:
    """Attach dispatch metadata to a worker method (sync or async)."""
    get_predefined_dispatch_fn(dispatch_mode)          # validate eagerly
    Execute(execute_mode)

    def deco(fn):
        if inspect.iscoroutinefunction(fn):
            @functools.wraps(fn)
            async def inner(*args, **kwargs):
                if materialize_futures:
                    args, kwargs = _materialize_futures(args, kwargs)
                return await fn(*args, **kwargs)
        else:
            @functools.wraps(fn)
            def inner(*args, **kwargs):
                if materialize_futures:
                    args, kwargs = _materialize_futures(args, kwargs)
                return fn(*args, **kwargs)
        setattr(inner, MAGIC_ATTR, {"dispatch_mode": dispatch_mode, "execute_mode": Execute(execute_mode), "blocking": blocking})
        return inner

    return deco


def dispatch_attr(fn) -> dict | None:
    return getattr(fn, MAGIC_ATTR, None)
