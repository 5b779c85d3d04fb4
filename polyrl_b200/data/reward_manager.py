"""Reward managers: turn a rollout batch into token-level scores (the scalar outcome reward sits on the last
valid response token).  ``naive`` scores sample by sample, ``batch`` hands the whole batch to the scorer,
``dapo`` adds the over-long soft penalty, ``prime`` scores in a thread pool; registry + ``load_reward_manager``
with custom scorer files.  Reference: 3rdparty/verl/verl/workers/reward_manager/{naive,batch,dapo,prime}.py,
rlboost/verl_stream/trainer/ppo/reward.py:95-174."""
from __future__ import annotations

import importlib.util
from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor
from typing import Callable

import torch

from ..protocol import DataProto
from .reward_score import default_compute_score

_REGISTRY: dict[str, type] = {}


def register(name: str):
    def deco(cls):
        _REGISTRY[name] = cls
        return cls

    return deco


def get_reward_manager_cls(name: str):
    if name not in _REGISTRY:
        raise ValueError(f"unknown reward manager {name!r}; have {sorted(_REGISTRY)}")
    return _REGISTRY[name]


def _decode(tokenizer, ids) -> str:
    if tokenizer is None:
        return " ".join(str(int(t)) for t in ids)
    return tokenizer.decode(ids, skip_special_tokens=True)


@register("naive")
class NaiveRewardManager:
    def __init__(self, tokenizer=None, num_examine: int = 0, compute_score: Callable | None = None,
                 reward_fn_key: str = "data_source", **kwargs):
        self.tokenizer, self.num_examine = tokenizer, num_examine
        self.compute_score = compute_score or default_compute_score
        self.reward_fn_key = reward_fn_key
        self.kwargs = kwargs

    def _items(self, data: DataProto):
        b = data.batch
        r = b["responses"].shape[-1]
        resp_mask = b["response_mask"] if "response_mask" in b else b["attention_mask"][:, -r:]
        lens = resp_mask.sum(-1).tolist()
        resp = b["responses"].cpu()
        nt = data.non_tensor_batch
        for i in range(len(data)):
            ids = resp[i, : int(lens[i])].tolist()
            extra = dict(nt["extra_info"][i]) if "extra_info" in nt and nt["extra_info"][i] is not None else {}
            extra["response_ids"] = ids
            yield i, int(lens[i]), ids, nt[self.reward_fn_key][i] if self.reward_fn_key in nt else "unknown", \
                nt["reward_model"][i]["ground_truth"], extra

    def _score_one(self, item) -> tuple[int, int, float, dict]:
        i, ln, ids, source, gt, extra = item
        res = self.compute_score(data_source=source, solution_str=_decode(self.tokenizer, ids), ground_truth=gt, extra_info=extra)
        if isinstance(res, dict):
            return i, ln, float(res["score"]), res
        return i, ln, float(res), {}

    def __call__(self, data: DataProto, return_dict: bool = False):
        if "rm_scores" in data.batch:
            out = data.batch["rm_scores"]
            return {"reward_tensor": out, "reward_extra_info": {}} if return_dict else out
        r = data.batch["responses"].shape[-1]
        out = torch.zeros(len(data), r, dtype=torch.float32)
        extra_info: dict[str, list] = defaultdict(list)
        printed: dict[str, int] = {}
        for item in self._items(data):
            i, ln, score, ex = self._score_one(item)
            score = self._post(score, ln, r)
            if ln > 0:
                out[i, ln - 1] = score
            for k, v in ex.items():
                extra_info[k].append(v)
            src = item[3]
            if printed.get(src, 0) < self.num_examine:
                printed[src] = printed.get(src, 0) + 1
                print(f"[reward] source={src} score={score} gt={item[4]} response={_decode(self.tokenizer, item[2])[:200]!r}")
        out = out.to(data.batch["responses"].device)
        return {"reward_tensor": out, "reward_extra_info": dict(extra_info)} if return_dict else out

    def _post(self, score: float, length: int, max_len: int) -> float:
        return score


@register("batch")
class BatchRewardManager(NaiveRewardManager):
    """``compute_score`` receives lists (data_sources, solution_strs, ground_truths, extra_infos) and returns a list."""

    def __call__(self, data: DataProto, return_dict: bool = False):
        items = list(self._items(data))
        scores = self.compute_score(data_sources=[it[3] for it in items],
                                    solution_strs=[_decode(self.tokenizer, it[2]) for it in items],
                                    ground_truths=[it[4] for it in items], extra_infos=[it[5] for it in items])
        r = data.batch["responses"].shape[-1]
        out = torch.zeros(len(data), r, dtype=torch.float32)
        for it, s in zip(items, scores):
            if it[1] > 0:
                out[it[0], it[1] - 1] = float(s["score"] if isinstance(s, dict) else s)
        out = out.to(data.batch["responses"].device)
        return {"reward_tensor": out, "reward_extra_info": {}} if return_dict else out


@register("dapo")
class DAPORewardManager(NaiveRewardManager):
    """Adds DAPO's soft over-long penalty: linear from 0 to -penalty_factor over the last ``overlong_buffer`` tokens."""

    def __init__(self, *a, max_resp_len: int | None = None, overlong_buffer_len: int = 0, penalty_factor: float = 1.0, **kw):
        super().__init__(*a, **kw)
        self.max_resp_len, self.buf, self.pf = max_resp_len, overlong_buffer_len, penalty_factor

    def _post(self, score: float, length: int, max_len: int) -> float:
        if self.buf <= 0:
            return score
        limit = (self.max_resp_len or max_len) - self.buf
        exceed = length - limit
        return score + min(-exceed / self.buf * self.pf, 0.0) if exceed > 0 else score


@register("prime")
class PrimeRewardManager(NaiveRewardManager):
    """Scores in a thread pool (scorers such as code sandboxes are latency bound)."""

    def __call__(self, data: DataProto, return_dict: bool = False):
        items = list(self._items(data))
        with ThreadPoolExecutor(max_workers=min(32, max(1, len(items)))) as ex:
            res = list(ex.map(self._score_one, items))
        r = data.batch["responses"].shape[-1]
        out = torch.zeros(len(data), r, dtype=torch.float32)
        for i, ln, score, _ in res:
            if ln > 0:
                out[i, ln - 1] = score
        out = out.to(data.batch["responses"].device)
        return {"reward_tensor": out, "reward_extra_info": {}} if return_dict else out


def get_custom_reward_fn(path: str | None, name: str = "compute_score") -> Callable | None:
    if not path:
        return None
    spec = importlib.util.spec_from_file_location("custom_reward_module", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not hasattr(mod, name):
        raise AttributeError(f"reward function {name!r} not found in {path}")
    return getattr(mod, name)


def load_reward_manager(config, tokenizer=None, num_examine: int = 0, **kwargs):
    name = config.reward_model.reward_manager
    fn = get_custom_reward_fn(config.custom_reward_function.path, config.custom_reward_function.name)
    rm = config.reward_model
    if getattr(rm, "use_reward_loop", False) or getattr(rm, "genrm_endpoints", None):
        from .reward_loop import GenerativeRewardModel, RewardLoopManager

        if rm.genrm_endpoints and fn is None:          # model-based judge instead of a rule-based scorer
            fn = GenerativeRewardModel(list(rm.genrm_endpoints), tokenizer, max_new_tokens=rm.genrm_max_new_tokens).compute_score
        return RewardLoopManager(name if name in ("naive", "dapo") else "naive", tokenizer, max_concurrency=rm.reward_loop_concurrency,
                                 compute_score=fn, reward_fn_key=config.data.reward_fn_key, **kwargs)
    ob = getattr(rm, "overlong_buffer", None)
    if name == "dapo" and ob is not None and ob.enable:
        kwargs.setdefault("max_resp_len", config.data.max_response_length)
        kwargs.setdefault("overlong_buffer_len", ob.len)
        kwargs.setdefault("penalty_factor", ob.penalty_factor)
    return get_reward_manager_cls(name)(tokenizer=tokenizer, num_examine=num_examine, compute_score=fn,
                                        reward_fn_key=config.data.reward_fn_key, **kwargs)


def compute_reward(data: DataProto, reward_fn) -> tuple[torch.Tensor, dict]:
    res = reward_fn(data, return_dict=True)
    return res["reward_tensor"], res.get("reward_extra_info", {})


def compute_reward_async(data: DataProto, reward_fn, executor: ThreadPoolExecutor | None = None):
    """Run the (CPU-bound) reward computation concurrently with GPU work; returns a future."""
    ex = executor or ThreadPoolExecutor(max_workers=1)
    return ex.submit(compute_reward, data, reward_fn)
