"""Rollout back-end registry.

``native`` is the product (paged-KV continuous-batching engine, rollout/engine.py); ``vllm-lib`` wraps the vLLM library
(rollout/vllm_adapter.py) as a comparison baseline.  ``naive`` / ``hf`` are
library-op baselines kept for debugging and as a numerics oracle for the engine: they re-run the full packed
forward over prompt+generated tokens each step (no KV cache), i.e. the role of the reference's
``NaiveRollout`` / ``HFRollout`` (3rdparty/verl/verl/workers/rollout/naive/naive_rollout.py:36,
hf_rollout.py:39) and its registry (rollout/base.py:80-102).
"""
from __future__ import annotations

from typing import Callable

import torch

from ..models import transformer as tf
from ..models.config import ModelConfig
from .engine import Request, RolloutEngine, SamplingParams

_ROLLOUTS: dict[str, Callable] = {}


def register_rollout(name: str):
    def deco(cls):
        _ROLLOUTS[name] = cls
        return cls
    return deco


def get_rollout_cls(name: str):
    name = {"sglang": "native", "vllm": "native", "sglang-disaggregated": "native-disaggregated"}.get(name, name)
    if name not in _ROLLOUTS:
        raise ValueError(f"unknown rollout {name!r}; registered: {sorted(_ROLLOUTS)}")
    return _ROLLOUTS[name]


register_rollout("native")(RolloutEngine)
register_rollout("native-disaggregated")(RolloutEngine)


@register_rollout("naive")
class NaiveRollout:
    """Cache-less batched sampling loop with the same ``generate`` / ``set_weights`` surface as ``RolloutEngine``."""

    def __init__(self, cfg: ModelConfig, device, max_model_len: int = 2048, seed: int = 0, **_):
        self.cfg, self.device, self.max_model_len = cfg, torch.device(device), max_model_len
        self.gen = torch.Generator(device="cpu").manual_seed(seed)
        self.embed_w = self.layer_w = self.head_w = None
        self.weight_version = 0

    def set_weights(self, embed_w, layer_w, head_w, version: int | None = None) -> None:
        self.embed_w, self.layer_w, self.head_w = embed_w, layer_w, head_w
        self.weight_version = self.weight_version + 1 if version is None else version

    def resume_memory(self, tags=()):
        pass

    def release_memory(self, tags=()):
        pass

    @torch.no_grad()
    def _next_logits(self, seqs: list[list[int]]) -> torch.Tensor:
        lens = [len(s) for s in seqs]
        ids = torch.tensor([t for s in seqs for t in s], dtype=torch.long, device=self.device)
        pos = torch.cat([torch.arange(n, device=self.device) for n in lens])
        cu = torch.tensor([0, *torch.tensor(lens).cumsum(0).tolist()], dtype=torch.int32, device=self.device)
        inp = tf.PackedInput(ids, pos, cu, max(lens))
        x = tf.embed_forward(self.cfg, self.embed_w, inp)
        for w in self.layer_w:
            x = tf.layer_forward(self.cfg, w, x, inp)
        last = x[(cu[1:] - 1).long()]
        h = tf.final_norm(self.cfg, self.head_w, last)
        return tf.logits_forward(self.cfg, self.head_w, self.embed_w, h).float()

    @torch.no_grad()
    def generate(self, prompts: list[list[int]], params: SamplingParams | dict | None = None, on_finish=None):
        p = params if isinstance(params, SamplingParams) else SamplingParams.from_dict(params)
        groups = [[Request(f"n{i}_{j}", list(pr), p) for j in range(p.n)] for i, pr in enumerate(prompts)]
        live = [r for g in groups for r in g]
        while live:
            logits = self._next_logits([r.prompt_ids + r.output_ids for r in live])
            if p.temperature > 0:
                lp = torch.log_softmax(logits / p.temperature, -1)
                tok = torch.multinomial(lp.exp().cpu(), 1, generator=self.gen).squeeze(1).to(lp.device)
            else:
                lp = torch.log_softmax(logits, -1)
                tok = lp.argmax(-1)
            sel = lp.gather(1, tok[:, None]).squeeze(1)
            nxt = []
            for r, t, l in zip(live, tok.tolist(), sel.tolist()):
                r.output_ids.append(t)
                r.logprobs.append(l)
                if not p.ignore_eos and (t == self.cfg.eos_token_id or t in p.stop_token_ids):
                    r.finish_reason = "stop"
                elif len(r.output_ids) >= p.max_new_tokens or r.num_tokens >= self.max_model_len:
                    r.finish_reason = "length"
                if r.finished:
                    if on_finish is not None:
                        on_finish(r)
                else:
                    nxt.append(r)
            live = nxt
        return groups


register_rollout("hf")(NaiveRollout)


def _vllm_lib(*a, **kw):
    from .vllm_adapter import VLLMRollout          # vLLM itself is imported only when this back-end is constructed

    return VLLMRollout(*a, **kw)


register_rollout("vllm-lib")(_vllm_lib)
