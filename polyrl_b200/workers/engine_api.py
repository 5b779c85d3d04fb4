"""Training-engine abstraction used by the role workers (reference: 3rdparty/verl/verl/workers/engine/base.py:26
``BaseEngine`` and its FSDP implementations ``FSDPEngineWithLMHead`` / ``WithValueHead``, fsdp/transformer_impl.py:81-974;
role-level losses verl/workers/roles/utils/losses.py:27-108).

``ZeroEngine`` adapts ``parallel.zero3.Zero3Engine`` to that surface: mode switches, forward-only inference over packed
micro-batches, forward+backward with a pluggable loss, optimizer / LR step, (sharded) checkpoint IO.
"""
from __future__ import annotations

import contextlib
from typing import Callable

import torch

from ..models import get_model_config, pack_padded, unpack_to_padded
from ..parallel.zero3 import Zero3Engine
from ..utils.checkpoint import ShardedCheckpointManager

_ENGINES: dict[tuple[str, str], type] = {}


def register_engine(model_type: str, backend: str = "zero"):
    def deco(cls):
        _ENGINES[(model_type, backend)] = cls
        return cls
    return deco


def get_engine_cls(model_type: str, backend: str = "zero"):
    """``backend``: zero (aliases fsdp / fsdp2) or tp (alias megatron): ``parallel.tp_train.TPTrainEngine`` — constructed with
    explicit (tp_group, dp_group), see tests/tp_train_worker.py."""
    backend = {"fsdp": "zero", "fsdp2": "zero", "megatron": "tp"}.get(backend, backend)
    if backend == "pp" and model_type == "language_model":
        return PipelineLMEngine
    if backend == "tp" and model_type == "language_model":
        from ..parallel.tp_train import TPTrainEngine

        return TPTrainEngine
    if (model_type, backend) not in _ENGINES:
        raise NotImplementedError(f"no training engine for model_type={model_type!r} backend={backend!r}")
    return _ENGINES[(model_type, backend)]


class BaseEngine:
    def initialize(self) -> None:
        raise NotImplementedError

    def train_mode(self):
        raise NotImplementedError

    def eval_mode(self):
        raise NotImplementedError

    def infer_batch(self, batch: dict, **kw) -> dict:
        raise NotImplementedError

    def train_batch(self, batch: dict, loss_fn: Callable, **kw) -> dict:
        raise NotImplementedError

    def optimizer_zero_grad(self) -> None:
        raise NotImplementedError

    def optimizer_step(self) -> dict:
        raise NotImplementedError

    def lr_scheduler_step(self) -> float:
        raise NotImplementedError

    def save_checkpoint(self, path: str, global_step: int = 0, max_ckpt_to_keep=None) -> None:
        raise NotImplementedError

    def load_checkpoint(self, path: str):
        raise NotImplementedError


@register_engine("language_model")
class ZeroEngine(BaseEngine):
    """LM-head engine: ``infer_batch`` -> log_probs / entropy of the response tokens; ``train_batch`` back-propagates
    ``loss_fn(outputs, batch)`` (called once per micro-batch, gradients accumulate in the sharded fp32 buffers)."""

    value_head = False

    def __init__(self, model: str | dict, device, group=None, lr: float = 1e-6, lr_schedule: Callable[[int], float] | None = None,
                 micro_batch_size: int = 8, **engine_kw):
        self.model_cfg = get_model_config(model, value_head=self.value_head) if isinstance(model, str) else model
        self.device, self.group = torch.device(device), group
        self.base_lr, self.lr_schedule = lr, lr_schedule
        self.micro_batch_size = micro_batch_size
        self.engine_kw = engine_kw
        self.engine: Zero3Engine | None = None
        self.lr_step = 0
        self._training = False

    def initialize(self) -> None:
        self.engine = Zero3Engine(self.model_cfg, self.device, group=self.group, lr=self.base_lr, **self.engine_kw)
        self.ckpt = ShardedCheckpointManager(self.engine, self.model_cfg)

    @contextlib.contextmanager
    def train_mode(self):
        prev, self._training = self._training, True
        try:
            yield self
        finally:
            self._training = prev

    @contextlib.contextmanager
    def eval_mode(self):
        prev, self._training = self._training, False
        try:
            with torch.no_grad():
                yield self
        finally:
            self._training = prev

    # ------------------------------------------------------------------ compute
    def _forward_micro(self, mb: dict, temperature: float) -> dict:
        ids, mask = mb["input_ids"].to(self.device), mb["attention_mask"].to(self.device)
        resp_len = mb["responses"].shape[1]
        inp, idx = pack_padded(ids, mask, mb.get("position_ids"))
        labels = torch.roll(inp.input_ids, -1)
        if self.value_head:
            v = self.engine.forward_values(inp)
            values = unpack_to_padded(v, idx, ids.shape[0], ids.shape[1])
            return {"values": values[:, -resp_len - 1:-1]}
        lp, ent = self.engine.forward_logprob_entropy(inp, labels, temperature)
        lp = unpack_to_padded(lp, idx, ids.shape[0], ids.shape[1])[:, -resp_len - 1:-1]
        ent = unpack_to_padded(ent, idx, ids.shape[0], ids.shape[1])[:, -resp_len - 1:-1]
        return {"log_probs": lp, "entropy": ent}

    def _micro_batches(self, batch: dict):
        n = batch["input_ids"].shape[0]
        for s in range(0, n, self.micro_batch_size):
            yield {k: (v[s:s + self.micro_batch_size] if torch.is_tensor(v) else v) for k, v in batch.items()}

    def infer_batch(self, batch: dict, temperature: float = 1.0) -> dict:
        outs: dict[str, list] = {}
        with torch.no_grad():
            for mb in self._micro_batches(batch):
                for k, v in self._forward_micro(mb, temperature).items():
                    outs.setdefault(k, []).append(v)
        return {k: torch.cat(v) for k, v in outs.items()}

    def train_batch(self, batch: dict, loss_fn: Callable, temperature: float = 1.0) -> dict:
        assert self._training, "call inside `with engine.train_mode():`"
        metrics: dict[str, list] = {}
        n = batch["input_ids"].shape[0]
        for mb in self._micro_batches(batch):
            out = self._forward_micro(mb, temperature)
            loss, m = loss_fn(out, mb)
            (loss * (mb["input_ids"].shape[0] / n)).backward()
            for k, v in m.items():
                metrics.setdefault(k, []).append(float(v))
        return {k: sum(v) / len(v) for k, v in metrics.items()}

    # ------------------------------------------------------------------ optimizer
    def optimizer_zero_grad(self) -> None:
        self.engine.zero_grad()

    def optimizer_step(self) -> dict:
        lr = self.lr_schedule(self.lr_step) if self.lr_schedule is not None else self.base_lr
        return {**self.engine.step(lr), "lr": lr}

    def lr_scheduler_step(self) -> float:
        self.lr_step += 1
        return self.lr_schedule(self.lr_step) if self.lr_schedule is not None else self.base_lr

    # ------------------------------------------------------------------ checkpoint
    def save_checkpoint(self, path: str, global_step: int = 0, max_ckpt_to_keep=None) -> None:
        self.ckpt.save(path, global_step, max_ckpt_to_keep, extra={"lr_step": self.lr_step})

    def load_checkpoint(self, path: str):
        extra = self.ckpt.load(path)
        if extra and "lr_step" in extra:
            self.lr_step = int(extra["lr_step"])
        return extra


@register_engine("value_model")
class ZeroEngineWithValueHead(ZeroEngine):
    value_head = True


class TensorParallelLMEngine(ZeroEngine):
    """``BaseEngine`` surface over ``parallel.tp_train.TPTrainEngine`` (dp x tp, optional Megatron sequence parallelism) — the
    reference's Megatron engine with ``tensor_model_parallel_size > 1``.  Every rank of a TP group receives the same batch and
    produces the same log-probs / metrics."""

    def __init__(self, model, device, tp_group, dp_group=None, lr: float = 1e-6, lr_schedule=None, micro_batch_size: int = 8,
                 sequence_parallel: bool = False, **engine_kw):
        super().__init__(model, device, group=None, lr=lr, lr_schedule=lr_schedule, micro_batch_size=micro_batch_size, **engine_kw)
        self.tp_group, self.dp_group, self.sequence_parallel = tp_group, dp_group, sequence_parallel

    def initialize(self) -> None:
        from ..parallel.tp_train import TPTrainEngine

        self.engine = TPTrainEngine(self.model_cfg, self.device, self.tp_group, self.dp_group, lr=self.base_lr,
                                    sequence_parallel=self.sequence_parallel, **self.engine_kw)
        self.ckpt = None


class ExpertParallelLMEngine(ZeroEngine):
    """``BaseEngine`` surface over ``parallel.expert.EPTrainEngine``: the experts of a mixture-of-experts model are split over the
    group, the dense part is data-parallel — every rank trains on its OWN batch chunk and the dispatch / combine all-to-alls
    require all ranks to run the same number of micro-batches."""

    def __init__(self, model, device, ep_group, lr: float = 1e-6, lr_schedule=None, micro_batch_size: int = 8, **engine_kw):
        super().__init__(model, device, group=None, lr=lr, lr_schedule=lr_schedule, micro_batch_size=micro_batch_size, **engine_kw)
        self.ep_group = ep_group

    def initialize(self) -> None:
        from ..parallel.expert import EPTrainEngine

        self.engine = EPTrainEngine(self.model_cfg, self.device, self.ep_group, lr=self.base_lr, **self.engine_kw)
        self.ckpt = None


class PipelineLMEngine(ZeroEngine):
    """``BaseEngine`` surface over ``parallel.pipeline.PipelineEngine`` (dp x pp, 1F1B) — the role of the reference's Megatron
    engine with ``pipeline_model_parallel_size > 1`` (verl/workers/engine/megatron/transformer_impl.py).  Every stage of a pipeline
    receives the same batch; losses and metrics exist on the last stage and are broadcast to the others."""

    def __init__(self, model, device, pp_group, dp_group=None, lr: float = 1e-6, lr_schedule=None, micro_batch_size: int = 8,
                 virtual_stages: int = 1, **engine_kw):
        super().__init__(model, device, group=None, lr=lr, lr_schedule=lr_schedule, micro_batch_size=micro_batch_size, **engine_kw)
        self.pp_group, self.dp_group, self.virtual_stages = pp_group, dp_group, int(virtual_stages)

    def initialize(self) -> None:
        from ..parallel.pipeline import InterleavedPipelineEngine, PipelineEngine

        if self.virtual_stages > 1:       # Megatron's virtual_pipeline_model_parallel_size: interleaved 1F1B over model chunks
            self.engine = InterleavedPipelineEngine(self.model_cfg, self.device, self.pp_group, self.dp_group,
                                                    virtual_stages=self.virtual_stages, lr=self.base_lr, **self.engine_kw)
        else:
            self.engine = PipelineEngine(self.model_cfg, self.device, self.pp_group, self.dp_group, lr=self.base_lr, **self.engine_kw)
        self.ckpt = None

    def _packed(self, mb: dict):
        ids, mask = mb["input_ids"].to(self.device), mb["attention_mask"].to(self.device)
        inp, idx = pack_padded(ids, mask, mb.get("position_ids"))
        return inp, torch.roll(inp.input_ids, -1), idx, ids.shape

    def _unpack(self, lp, ent, idx, shape, resp_len):
        lp = unpack_to_padded(lp, idx, shape[0], shape[1])[:, -resp_len - 1:-1]
        ent = unpack_to_padded(ent, idx, shape[0], shape[1])[:, -resp_len - 1:-1]
        return {"log_probs": lp, "entropy": ent}

    def infer_batch(self, batch: dict, temperature: float = 1.0) -> dict:
        mbs = list(self._micro_batches(batch))
        packed = [self._packed(mb) for mb in mbs]
        res = self.engine.forward_only([(p[0], p[1]) for p in packed], temperature)
        outs = [self._unpack(lp, ent, p[2], p[3], mb["responses"].shape[1]) for (lp, ent), p, mb in zip(res, packed, mbs)]
        return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}

    def train_batch(self, batch: dict, loss_fn: Callable, temperature: float = 1.0) -> dict:
        import torch.distributed as dist

        assert self._training, "call inside `with engine.train_mode():`"
        mbs = list(self._micro_batches(batch))
        packed = [self._packed(mb) for mb in mbs]
        n = batch["input_ids"].shape[0]
        metrics: dict[str, list] = {}

        def stage_loss(lp, ent, k):
            out = self._unpack(lp, ent, packed[k][2], packed[k][3], mbs[k]["responses"].shape[1])
            loss, m = loss_fn(out, mbs[k])
            for key, v in m.items():
                metrics.setdefault(key, []).append(float(v))
            return loss * (mbs[k]["input_ids"].shape[0] / n)

        self.engine.train_batch([(p[0], p[1]) for p in packed], stage_loss, temperature)
        res = [{k: sum(v) / len(v) for k, v in metrics.items()}]
        dist.broadcast_object_list(res, src=dist.get_process_group_ranks(self.pp_group)[-1], group=self.pp_group)
        return res[0]

    def save_checkpoint(self, path: str, global_step: int = 0, max_ckpt_to_keep=None) -> None:
        """Every stage writes its own slice (fp32 masters + moments)."""
        import os

        e = self.engine
        os.makedirs(path, exist_ok=True)
        torch.save({"master": e.master, "exp_avg": e.exp_avg, "exp_avg_sq": e.exp_avg_sq, "opt_step": e.opt_step, "lr_step": self.lr_step,
                    "global_step": global_step}, os.path.join(path, f"pp_stage_{e.stage}_of_{e.pp}.pt"))

    def load_checkpoint(self, path: str):
        import os

        e = self.engine
        sd = torch.load(os.path.join(path, f"pp_stage_{e.stage}_of_{e.pp}.pt"), map_location=e.device, weights_only=True)
        for k in e.master:
            e.master[k].copy_(sd["master"][k])
            e.exp_avg[k].copy_(sd["exp_avg"][k])
            e.exp_avg_sq[k].copy_(sd["exp_avg_sq"][k])
            e.params[k].data.copy_(e.master[k].to(e.dtype))
        e.opt_step, self.lr_step = int(sd["opt_step"]), int(sd["lr_step"])
        return {"lr_step": self.lr_step, "global_step": sd.get("global_step", 0)}


# ---- role-level losses (reference: workers/roles/utils/losses.py) --------------------------------------------------------
def sft_loss(out: dict, mb: dict):
    mask = mb["response_mask"].to(out["log_probs"].device).float()
    loss = -(out["log_probs"] * mask).sum() / mask.sum().clamp_min(1)
    return loss, {"sft/loss": loss.item()}


def ppo_loss(config=None, clip_ratio: float = 0.2, loss_agg_mode: str = "token-mean"):
    from ..algos import core_algos

    # ``config`` may be the actor config tree (policy_loss.loss_mode, entropy_coeff, use_kl_loss, ...) or None (plain dual-clip PPO)
    def get(name, default):
        return getattr(config, name, default) if config is not None else default

    pl = get("policy_loss", None)
    mode = (pl.get("loss_mode") if isinstance(pl, dict) else getattr(pl, "loss_mode", None)) or "vanilla"
    loss_impl = core_algos.get_policy_loss_fn(mode)
    ent_coeff = float(get("entropy_coeff", 0.0) or 0.0)
    use_kl, kl_coef, kl_type = bool(get("use_kl_loss", False)), float(get("kl_loss_coef", 0.001)), get("kl_loss_type", "low_var_kl")

    def fn(out: dict, mb: dict):
        dev = out["log_probs"].device
        mask = mb["response_mask"].to(dev).float()
        kw = dict(cliprange=clip_ratio) if mode == "vanilla" else {}
        riw = mb.get("rollout_is_weights")
        pg, m = loss_impl(mb["old_log_probs"].to(dev), out["log_probs"], mb["advantages"].to(dev), mask, loss_agg_mode=loss_agg_mode,
                          config=config, rollout_is_weights=riw.to(dev) if riw is not None else None, **kw)
        loss, metrics = pg, {"actor/pg_loss": pg.item(), **m}
        if ent_coeff != 0.0:
            ent = core_algos.agg_loss(out["entropy"], mask, loss_agg_mode)
            loss = loss - ent_coeff * ent
            metrics["actor/entropy"] = ent.item()
        if use_kl:
            if "ref_log_prob" not in mb:
                raise ValueError("use_kl_loss needs ref_log_prob in the batch")
            kl = core_algos.agg_loss(core_algos.kl_penalty(out["log_probs"], mb["ref_log_prob"].to(dev), kl_type), mask, loss_agg_mode)
            loss = loss + kl_coef * kl
            metrics.update({"actor/kl_loss": kl.item(), "actor/kl_coef": kl_coef})
        return loss, metrics
    return fn


def value_loss(cliprange_value: float = 0.5):
    from ..algos import core_algos

    def fn(out: dict, mb: dict):
        dev = out["values"].device
        mask = mb["response_mask"].to(dev).float()
        vf, clipfrac = core_algos.compute_value_loss(out["values"], mb["returns"].to(dev), mb["values"].to(dev), mask, cliprange_value)
        return vf, {"critic/vf_loss": vf.item(), "critic/vf_clipfrac": float(clipfrac)}
    return fn
