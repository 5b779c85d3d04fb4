"""vLLM as a LIBRARY rollout back-end behind the engine surface (``generate`` / ``set_weights`` / ``release_memory`` /
``resume_memory``) — the counterpart of the reference's vLLM SPMD rollout (3rdparty/verl/verl/workers/rollout/vllm_rollout/
vllm_rollout_spmd.py: token-id prompts, per-token log-probs, sleep / wake_up around training, weights reloaded from the trainer
after every update).  It is a baseline to compare the native engine against, not the product path: ``rollout.name=vllm-lib``.

vLLM needs a checkpoint directory for its model config (``model_path``); the weights that matter arrive through ``set_weights``
(HF names via ``export_hf_state_dict``).  The heavy import happens at construction so the rest of the package never depends on it.
"""
from __future__ import annotations

import torch

from ..models import transformer as tf
from ..models.config import ModelConfig
from .engine import Request, SamplingParams


def _load_weights_on_worker(worker, named):            # runs inside every vLLM worker (collective_rpc)
    model = worker.model_runner.model
    model.load_weights(named)
    return True


class VLLMRollout:
    def __init__(self, cfg: ModelConfig, device, max_model_len: int = 2048, model_path: str | None = None,
                 gpu_memory_utilization: float = 0.5, tensor_parallel_size: int = 1, seed: int = 0, enforce_eager: bool = False,
                 llm=None, **_):
        self.cfg, self.device, self.max_model_len = cfg, torch.device(device), max_model_len
        self.weight_version = 0
        self._asleep = False
        if llm is None:
            if model_path is None:
                raise ValueError("the vLLM back-end needs model_path (an HF checkpoint directory) for its model config")
            from vllm import LLM                      # noqa: WPS433 — optional dependency, imported on use

            llm = LLM(model=model_path, skip_tokenizer_init=True, enable_sleep_mode=True, max_model_len=max_model_len,
                      gpu_memory_utilization=gpu_memory_utilization, tensor_parallel_size=tensor_parallel_size, seed=seed,
                      enforce_eager=enforce_eager, dtype="bfloat16", enable_prefix_caching=True)
        self.llm = llm

    # ------------------------------------------------------------------ weights / memory
    def set_weights(self, embed_w: dict, layer_w: list[dict], head_w: dict, version: int | None = None, presharded: bool = False) -> None:
        if presharded:
            raise NotImplementedError("vLLM shards full HF tensors itself; pass unsharded weights")
        ours = dict(embed_w)
        for i, lw in enumerate(layer_w):
            ours.update({f"layers.{i}.{k}": v for k, v in lw.items()})
        ours.update(head_w)
        named = list(tf.export_hf_state_dict(self.cfg, ours).items())
        if self.cfg.tie_word_embeddings:
            named = [(k, v) for k, v in named if k != "lm_head.weight"]
        self.llm.collective_rpc(_load_weights_on_worker, args=(named,))
        if hasattr(self.llm, "reset_prefix_cache"):
            self.llm.reset_prefix_cache()             # cached KV belongs to the old policy
        self.weight_version = self.weight_version + 1 if version is None else version

    def release_memory(self, tags=("kv_cache",)) -> None:
        if not self._asleep:
            self.llm.sleep(level=1)
            self._asleep = True

    def resume_memory(self, tags=("kv_cache",)) -> None:
        if self._asleep:
            self.llm.wake_up()
            self._asleep = False

    # ------------------------------------------------------------------ generation
    def _sampling_params(self, p: SamplingParams):
        kw = dict(n=p.n, temperature=p.temperature, top_p=p.top_p, top_k=p.top_k if p.top_k and p.top_k > 0 else -1,
                  max_tokens=p.max_new_tokens, logprobs=0, ignore_eos=p.ignore_eos, seed=p.seed, detokenize=False,
                  stop_token_ids=[self.cfg.eos_token_id, *self.cfg.extra_eos_token_ids, *p.stop_token_ids])
        try:
            from vllm import SamplingParams as VSP
        except Exception:                              # adapter tests inject a fake ``llm`` without vLLM installed
            return kw
        return VSP(**kw)

    def generate(self, prompts: list[list[int]], params: SamplingParams | dict | None = None, on_finish=None, mm=None):
        if mm is not None and any(m is not None for m in mm):
            raise NotImplementedError("image prompts go through the native engine")
        p = params if isinstance(params, SamplingParams) else SamplingParams.from_dict(params)
        self.resume_memory()
        outs = self.llm.generate([{"prompt_token_ids": list(map(int, pr))} for pr in prompts], self._sampling_params(p), use_tqdm=False)
        groups = []
        for i, (pr, out) in enumerate(zip(prompts, outs)):
            grp = []
            for j, c in enumerate(out.outputs):
                toks = list(c.token_ids)
                lps = [float(d[t].logprob) for d, t in zip(c.logprobs or [], toks)]
                r = Request(f"v{i}_{j}", list(pr), p, output_ids=toks, logprobs=lps,
                            finish_reason={"stop": "stop", "length": "length", "abort": "abort"}.get(c.finish_reason or "length", "stop"))
                grp.append(r)
                if on_finish is not None:
                    on_finish(r)
            groups.append(grp)
        return groups

    def has_work(self) -> bool:
        return False

    def server_info(self) -> dict:
        return {"#running_req": 0, "#queue_req": 0, "last_gen_throughput": 0.0, "weight_version": self.weight_version}
