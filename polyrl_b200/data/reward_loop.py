"""Asynchronous reward loop: score samples one by one, concurrently, as they become available — rule-based scorers in a thread
pool, model-based ("generative") rewards through rollout engine servers behind a round-robin router.  Counterpart of the reference's
experimental reward stack (3rdparty/verl/verl/experimental/reward/reward_loop/{base,naive,dapo,registry}.py — per-sample
``async run_single``; reward_model.py + router/naive_router.py — reward-model servers behind a router).

    loop = RewardLoopManager("dapo", tokenizer, max_resp_len=..., overlong_buffer_len=...)
    scores = loop.compute(batch)                 # -> {"reward_tensor": [b, R], "reward_extra_info": {...}}

    rm = GenerativeRewardModel(["127.0.0.1:30000"], tokenizer, template=..., parse=...)   # GenRM scoring over HTTP
    loop = RewardLoopManager("naive", tokenizer, compute_score=rm.compute_score)
"""
from __future__ import annotations

import asyncio
import http.client
import inspect
import itertools
import json
from collections import defaultdict
from typing import Callable

import torch

from ..protocol import DataProto
from .reward_manager import _decode
from .reward_score import default_compute_score

_LOOPS: dict[str, type] = {}


def register_reward_loop(name: str):
    def deco(cls):
        _LOOPS[name] = cls
        return cls

    return deco


def get_reward_loop_cls(name: str):
    if name not in _LOOPS:
        raise ValueError(f"unknown reward loop {name!r}; have {sorted(_LOOPS)}")
    return _LOOPS[name]


class RewardLoopBase:
    """One sample in, one ``{"reward_score": float, "reward_extra_info": {...}}`` out.  ``compute_score`` may be a plain function
    (run in the default executor so slow scorers overlap) or a coroutine function (awaited directly, e.g. an HTTP reward model)."""

    def __init__(self, tokenizer=None, compute_score: Callable | None = None, reward_fn_key: str = "data_source", **kwargs):
        self.tokenizer, self.reward_fn_key = tokenizer, reward_fn_key
        self.compute_score = compute_score or default_compute_score
        self.is_async = inspect.iscoroutinefunction(self.compute_score)

    async def _score(self, **kw):
        if self.is_async:
            return await self.compute_score(**kw)
        return await asyncio.get_running_loop().run_in_executor(None, lambda: self.compute_score(**kw))

    async def run_single(self, sample: dict) -> dict:
        raise NotImplementedError


@register_reward_loop("naive")
class NaiveRewardLoop(RewardLoopBase):
    async def run_single(self, sample: dict) -> dict:
        res = await self._score(data_source=sample["data_source"], solution_str=_decode(self.tokenizer, sample["response_ids"]),
                                ground_truth=sample["ground_truth"], extra_info=sample["extra_info"])
        if isinstance(res, dict):
            return {"reward_score": float(res["score"]), "reward_extra_info": {k: v for k, v in res.items()}}
        return {"reward_score": float(res), "reward_extra_info": {}}


@register_reward_loop("dapo")
class DAPORewardLoop(NaiveRewardLoop):
    """Adds DAPO's soft over-long penalty: linear from 0 to ``-penalty_factor`` over the last ``overlong_buffer_len`` tokens."""

    def __init__(self, *a, max_resp_len: int | None = None, overlong_buffer_len: int = 0, penalty_factor: float = 1.0, **kw):
        super().__init__(*a, **kw)
        self.max_resp_len, self.buf, self.pf = max_resp_len, overlong_buffer_len, penalty_factor

    async def run_single(self, sample: dict) -> dict:
        out = await super().run_single(sample)
        if self.buf > 0 and self.max_resp_len:
            expected = self.max_resp_len - self.buf
            pen = min(-(len(sample["response_ids"]) - expected) / self.buf * self.pf, 0.0)
            out["reward_score"] += pen
            out["reward_extra_info"]["overlong_reward"] = pen
            out["reward_extra_info"]["overlong"] = pen < 0
        return out


class RewardLoopManager:
    """Scores a batch (or a stream of samples) with bounded concurrency; drop-in for a reward manager (``__call__``)."""

    def __init__(self, name: str = "naive", tokenizer=None, max_concurrency: int = 64, **kwargs):
        self.loop_impl: RewardLoopBase = get_reward_loop_cls(name)(tokenizer, **kwargs)
        self.max_concurrency = max_concurrency

    @staticmethod
    def samples(data: DataProto, reward_fn_key: str = "data_source"):
        b, nt = data.batch, data.non_tensor_batch
        r = b["responses"].shape[-1]
        mask = b["response_mask"] if "response_mask" in b else b["attention_mask"][:, -r:]
        lens = mask.sum(-1).tolist()
        resp = b["responses"].cpu()
        for i in range(len(data)):
            ids = resp[i, : int(lens[i])].tolist()
            extra = dict(nt["extra_info"][i]) if "extra_info" in nt and nt["extra_info"][i] is not None else {}
            yield {"index": i, "response_ids": ids, "data_source": nt[reward_fn_key][i] if reward_fn_key in nt else "unknown",
                   "ground_truth": nt["reward_model"][i]["ground_truth"], "extra_info": extra}

    async def compute_async(self, samples: list[dict]) -> list[dict]:
        sem = asyncio.Semaphore(self.max_concurrency)

        async def one(s):
            async with sem:
                return await self.loop_impl.run_single(s)

        return await asyncio.gather(*(one(s) for s in samples))

    def compute(self, data: DataProto) -> dict:
        samples = list(self.samples(data, self.loop_impl.reward_fn_key))
        results = asyncio.run(self.compute_async(samples))
        r = data.batch["responses"].shape[-1]
        out = torch.zeros(len(data), r, dtype=torch.float32)
        extra: dict[str, list] = defaultdict(list)
        for s, res in zip(samples, results):
            n = len(s["response_ids"])
            if n > 0:
                out[s["index"], n - 1] = res["reward_score"]
            for k, v in res["reward_extra_info"].items():
                extra[k].append(v)
        return {"reward_tensor": out.to(data.batch["responses"].device), "reward_extra_info": dict(extra)}

    def __call__(self, data: DataProto, return_dict: bool = False):
        res = self.compute(data)
        return res if return_dict else res["reward_tensor"]


class GenerativeRewardModel:
    """Model-based rewards from rollout engine servers (``rollout/server.py`` or any SGLang-compatible ``/generate``): the judge
    prompt is built from ``template(question, response, ground_truth)``, generated greedily, and ``parse(text) -> float`` turns the
    judge's text into a score.  Requests go round-robin over ``endpoints`` (the reference's naive router)."""

    def __init__(self, endpoints: list[str], tokenizer, template: Callable[[str, str, str], str] | None = None,
                 parse: Callable[[str], float] | None = None, max_new_tokens: int = 64, timeout: float = 600.0):
        if not endpoints:
            raise ValueError("at least one reward-model endpoint is required")
        self.endpoints, self.tok = list(endpoints), tokenizer
        self._rr = itertools.cycle(range(len(self.endpoints)))
        self.template = template or (lambda q, r, gt: f"Question: {q}\nReference answer: {gt}\nResponse: {r}\nIs the response correct? Answer yes or no.\n")
        self.parse = parse or (lambda text: 1.0 if text.strip().lower().startswith("yes") else 0.0)
        self.max_new_tokens, self.timeout = max_new_tokens, timeout

    def _post(self, endpoint: str, body: dict) -> dict:
        host, port = endpoint.rsplit(":", 1)
        conn = http.client.HTTPConnection(host, int(port), timeout=self.timeout)
        try:
            conn.request("POST", "/generate", body=json.dumps(body), headers={"Content-Type": "application/json"})
            r = conn.getresponse()
            data = r.read().decode()
            if r.status != 200:
                raise RuntimeError(f"reward model {endpoint} returned HTTP {r.status}: {data[:200]}")
            return json.loads(data)
        finally:
            conn.close()

    def judge(self, question: str, response: str, ground_truth: str) -> tuple[float, str]:
        ids = self.tok.encode(self.template(question, response, ground_truth), add_special_tokens=False)
        out = self._post(self.endpoints[next(self._rr)], {"input_ids": list(map(int, ids)),
                                                           "sampling_params": {"max_new_tokens": self.max_new_tokens, "temperature": 0.0}})
        toks = [int(t[1]) for t in out.get("meta_info", {}).get("output_token_logprobs", [])]
        text = self.tok.decode(toks, skip_special_tokens=True)
        return float(self.parse(text)), text

    async def compute_score(self, data_source, solution_str, ground_truth, extra_info=None):
        q = (extra_info or {}).get("question", "")
        score, text = await asyncio.get_running_loop().run_in_executor(None, self.judge, q, solution_str, str(ground_truth))
        return {"score": score, "judge_text": text}
