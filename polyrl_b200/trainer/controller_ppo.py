"""Single-controller PPO / GRPO: ONE driver process owns the data loader, rewards, advantages and the control flow; the model
work runs in a ``WorkerGroup`` (one OS process per GPU, or Ray actors, or in-process workers) whose methods are called through
the ``@register`` dispatch protocol.

This is the reference's execution model — ``RayPPOTrainer.fit`` driving ``RayWorkerGroup``s of fused
actor / rollout / ref (+ critic) workers (3rdparty/verl/verl/trainer/ppo/ray_trainer.py:968-1214,
verl/single_controller/ray/base.py:262-444, verl/workers/fsdp_workers.py:134-1100) — on this framework's own controller
(``polyrl_b200/controller``).  The SPMD trainers (``trainer/stream_trainer.py``, ``trainer/colocated.py``) stay the fast path:
there a batch never leaves the device between phases; here every phase costs one host round trip per worker, which is what
the reference pays through the Ray object store.  Both drive the same ``ActorRolloutRefWorker`` / ``CriticWorker``.

    python -m polyrl_b200.trainer.controller_ppo controller.backend=mp trainer.n_gpus_per_node=2 ...   # one process per GPU
    python -m polyrl_b200.trainer.controller_ppo controller.backend=ray ...                            # Ray actors (if installed)
"""
from __future__ import annotations

import os
import sys
import time
import uuid
from typing import Callable

import numpy as np
import torch

from ..algos import core_algos
from ..config.schema import PPOConfig, load_config, need_critic, need_reference_policy
from ..controller import (ClassWithInitArgs, Dispatch, Execute, LocalWorkerGroup, MPWorkerGroup, Worker,
                          make_nd_compute_dataproto_dispatch_fn, register)
from ..protocol import DataProto, collate_fn
from ..utils import metrics as M
from ..utils.profiler import marked_timer
from . import core


class PPOWorker(Worker):
    """Fused actor + rollout (+ ref) (+ critic) worker of one GPU, exposed to the driver through ``@register`` methods.
    DataProto arguments arrive on the host (they crossed a pipe / the Ray object store), are moved to the device for the call
    and results go back to the host."""

    def __init__(self, config: PPOConfig, init_dist: bool = True):
        super().__init__()
        self.config = config
        self.inner = None
        self.critic = None
        want_cuda = torch.cuda.is_available() and config.trainer.device != "cpu"
        self.device = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}" if want_cuda else "cpu")
        if init_dist and self.world_size > 1:
            import torch.distributed as dist

            if not dist.is_initialized():
                if want_cuda:
                    torch.cuda.set_device(self.device)
                    dist.init_process_group("nccl", rank=self.rank, world_size=self.world_size, device_id=self.device)
                else:
                    dist.init_process_group("gloo", rank=self.rank, world_size=self.world_size)

    # ---- helpers
    def _in(self, data: DataProto) -> DataProto:
        return data.to(self.device)

    @staticmethod
    def _out(data: DataProto) -> DataProto:
        return data.to("cpu")

    # ---- lifecycle
    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def init_model(self) -> dict:
        from ..workers.hybrid_worker import ActorRolloutRefWorker, CriticWorker

        cfg = self.config
        role = "actor_rollout_ref" if need_reference_policy(cfg) else "actor_rollout"
        self.inner = ActorRolloutRefWorker(cfg, role=role, device=self.device, group=None)
        self.inner.init_model()
        if need_critic(cfg):
            self.critic = CriticWorker(cfg, device=self.device, group=None)
            self.critic.init_model(heap=self.inner.actor_engine.heap)
        self.rm = None
        if cfg.reward_model.enable:
            from ..workers.hybrid_worker import RewardModelWorker

            self.rm = RewardModelWorker(cfg, device=self.device, group=None)
            self.rm.init_model(heap=self.inner.actor_engine.heap)
        mc = self.inner.model_cfg
        return {"rank": self.rank, "vocab_size": mc.vocab_size, "eos_token_id": mc.eos_token_id, "pad_token_id": mc.pad_token_id,
                "num_params": self.inner.actor_engine.num_params()}

    # ---- rollout / inference
    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def generate_sequences(self, prompts: DataProto) -> DataProto:
        return self._out(self.inner.generate_sequences(self._in(prompts)))

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def compute_log_prob(self, data: DataProto) -> DataProto:
        return self._out(self.inner.compute_log_prob(self._in(data)))

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def compute_ref_log_prob(self, data: DataProto) -> DataProto:
        return self._out(self.inner.compute_ref_log_prob(self._in(data)))

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def compute_rm_score(self, data: DataProto) -> DataProto:
        return self._out(self.rm.compute_rm_score(self._in(data)))

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def compute_values(self, data: DataProto) -> DataProto:
        return self._out(self.critic.compute_values(self._in(data)))

    # ---- updates: the driver sends one mini-batch per call; every worker trains on its chunk, gradients are averaged by ZeRO
    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def update_actor(self, data: DataProto) -> DataProto:
        return self.inner.update_actor_stream(self._in(data))

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def update_critic(self, data: DataProto) -> DataProto:
        return self.critic.update_critic_stream(self._in(data))

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def update_weight(self) -> bool:
        return bool(self.inner.update_weight_remote())

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def update_reference_policy(self) -> None:
        self.inner.update_reference_policy()

    # ---- implicit process reward model (PRIME recipe): a policy-shaped trainable LM scored against a frozen copy of itself
    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def init_prm(self) -> int:
        from ..algos.prime import ImplicitPRM
        from ..models import get_model_config
        from ..parallel.zero3 import Zero3Engine
        from ..workers.hybrid_worker import maybe_load_pretrained

        rm = self.config.reward_model
        path = rm.model.path if rm.model.path != "tiny-llama" or self.config.actor_rollout_ref.model.path == "tiny-llama" else self.config.actor_rollout_ref.model.path
        mcfg = get_model_config(path, **(rm.model.override_config or {}))
        comm = rm.model.fsdp_config.comm_backend if self.device.type == "cuda" else "nccl"
        heap = self.inner.actor_engine.heap if self.inner is not None and self.inner.actor_engine is not None else None
        o = rm.model.optim
        prm = Zero3Engine(mcfg, self.device, comm=comm, trainable=True, seed=self.config.data.seed, lr=o.lr, betas=tuple(o.betas), eps=o.eps,
                          weight_decay=o.weight_decay, max_grad_norm=o.clip_grad, heap=heap)
        ref = Zero3Engine(mcfg, self.device, comm=comm, trainable=False, seed=self.config.data.seed, heap=heap)
        for e in (prm, ref):
            maybe_load_pretrained(e, rm.model.ref_path if (e is ref and rm.model.ref_path) else path, self.rank)
        self.prm = ImplicitPRM(prm, ref, beta=rm.model.beta_train, micro_batch_size=rm.micro_batch_size_per_gpu or 8)
        return prm.num_params()

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def compute_prm_score(self, data: DataProto) -> DataProto:
        return DataProto.from_dict({"prm_scores": self.prm.score(self._in(data)).float()}).to("cpu")

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def update_prm(self, data: DataProto) -> DataProto:
        data = self._in(data)
        m = self.prm.update(data, data.batch["acc"])
        return DataProto(meta_info={"metrics": {k: [v] for k, v in m.items()}})

    # ---- checkpoints
    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def save_checkpoint(self, path: str, global_step: int = 0, max_ckpt_to_keep: int | None = None) -> None:
        self.inner.save_checkpoint(os.path.join(path, "actor"), None, global_step, max_ckpt_to_keep)
        if self.critic is not None:
            self.critic.save_checkpoint(os.path.join(path, "critic"), None, global_step, max_ckpt_to_keep)

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def load_checkpoint(self, path: str) -> None:
        self.inner.load_checkpoint(os.path.join(path, "actor"))
        if self.critic is not None:
            self.critic.load_checkpoint(os.path.join(path, "critic"))

    @register(dispatch_mode=Dispatch.RANK_ZERO, execute_mode=Execute.RANK_ZERO)
    def weights_checksum(self) -> float:
        s = 0.0
        for _, t in self.inner.actor_engine.named_full_tensors():
            s += float(t.double().abs().sum())
        return s

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def teardown(self) -> None:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()


class PipelinePPOWorker(Worker):
    """``actor.strategy = megatron``: the actor is MODEL-parallel — pipeline (``actor.megatron.pipeline_model_parallel_size`` stages,
    optional virtual stages, ``parallel/pipeline.py``) or tensor parallel (``tensor_model_parallel_size``, ``parallel/tp_train.py``)
    x data-parallel replicas — through the ``BaseEngine`` role worker
    (``workers/roles.py::ActorWorker``), the way the reference's Megatron workers sit behind the same trainer
    (3rdparty/verl/verl/workers/megatron_workers.py:204-600).  Dispatch uses the nd-mesh protocol: every stage of a pipeline
    receives its replica's chunk, results are collected from the last stage.  Rollout stays data-parallel over ALL workers: each
    one holds a full-model engine that is refreshed from the gathered pipeline state after every update (the role of the
    reference's Megatron -> inference-engine resharding).  GRPO-style algorithms only (no critic / reference policy here)."""

    def __init__(self, config: PPOConfig):
        super().__init__()
        import torch.distributed as dist

        self.config = config
        if need_critic(config):
            raise NotImplementedError("the model-parallel actor worker supports critic-free algorithms (GRPO / RLOO / ...)")
        want_cuda = torch.cuda.is_available() and config.trainer.device != "cpu"
        self.device = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}" if want_cuda else "cpu")
        if not dist.is_initialized():
            if want_cuda:
                torch.cuda.set_device(self.device)
                dist.init_process_group("nccl", rank=self.rank, world_size=self.world_size, device_id=self.device)
            else:
                dist.init_process_group("gloo", rank=self.rank, world_size=self.world_size)
        mg = config.actor_rollout_ref.actor.megatron
        self.pp = int(mg.pipeline_model_parallel_size)
        self.tp = int(mg.tensor_model_parallel_size)
        if self.tp > 1 and self.pp > 1:
            raise NotImplementedError("tensor parallelism inside pipeline stages: choose tensor_ OR pipeline_model_parallel_size > 1")
        self.use_tp = self.tp > 1
        self.use_ep = int(mg.expert_model_parallel_size) > 1
        if self.use_ep:
            if self.tp > 1 or self.pp > 1 or int(mg.expert_model_parallel_size) != self.world_size:
                raise NotImplementedError("expert parallelism spans all workers and does not combine with tensor / pipeline parallelism here")
            self.pp = 1                  # no replication of data inside a model-parallel group: every worker has its own chunk
        if self.use_tp:
            self.pp = self.tp            # one model-parallel group per replica; below "pp" is the size of that group
        if (self.pp < 2 and not self.use_ep) or self.world_size % self.pp:
            raise ValueError(f"model-parallel size {self.pp} must be >= 2 and divide the {self.world_size} workers")
        self.dp = self.world_size // self.pp
        self.pp_group = self.dp_group = None
        for d in range(self.dp):                        # every process creates every group, in the same order
            g = dist.new_group(list(range(d * self.pp, (d + 1) * self.pp)))
            if d == self.rank // self.pp:
                self.pp_group = g
        for st in range(self.pp):
            g = dist.new_group(list(range(st, self.world_size, self.pp)))
            if st == self.rank % self.pp:
                self.dp_group = g
        # collect from the last pipeline stage (the head lives there) or from TP rank 0 (every TP rank holds the same result)
        collect = (self.rank % self.pp == 0) if self.use_tp else (self.rank % self.pp == self.pp - 1)
        self._register_dispatch_collect_info("actor", dp_rank=self.rank // self.pp, is_collect=collect, dp_size=self.dp)
        self.actor = self.rollout = self.roll_w = None

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def init_model(self) -> dict:
        from ..models import TransformerWeights, get_model_config
        from ..rollout.engine import RolloutEngine
        import torch.distributed as dist

        from ..workers.engine_api import ExpertParallelLMEngine, TensorParallelLMEngine, get_engine_cls
        from ..workers.hybrid_worker import lr_at_step
        from ..workers.roles import ActorWorker

        cfg = self.config
        arr, a = cfg.actor_rollout_ref, cfg.actor_rollout_ref.actor
        self.model_cfg = get_model_config(arr.model.path, **arr.model.override_config)
        total = a.optim.total_training_steps if a.optim.total_training_steps > 0 else (cfg.trainer.total_training_steps or 0)
        common = dict(lr=a.optim.lr, lr_schedule=lambda step: lr_at_step(a.optim, step, total),
                      micro_batch_size=a.ppo_micro_batch_size_per_gpu or 8, weight_decay=a.optim.weight_decay,
                      max_grad_norm=a.grad_clip, seed=cfg.data.seed)
        if self.use_ep:
            eng = ExpertParallelLMEngine(self.model_cfg, self.device, dist.group.WORLD, **common)
        elif self.use_tp:
            eng = TensorParallelLMEngine(self.model_cfg, self.device, self.pp_group, self.dp_group if self.dp > 1 else None, **common)
        else:
            eng = get_engine_cls("language_model", "pp")(
                self.model_cfg, self.device, self.pp_group, self.dp_group if self.dp > 1 else None,
                virtual_stages=int(a.megatron.virtual_pipeline_model_parallel_size or 1), **common)
        self.actor = ActorWorker(eng, ppo_mini_batch_size=1 << 30, ppo_epochs=1, clip_ratio=a.clip_ratio, loss_agg_mode=a.loss_agg_mode,
                                 loss_config=a)
        self.actor.init_model()
        self.ref = None
        if need_reference_policy(cfg):
            # frozen copy of the initial policy with the same model-parallel layout (same seed => identical initial weights)
            ref_kw = {**common, "lr": 0.0, "lr_schedule": None}
            if self.use_ep:
                ref_eng = ExpertParallelLMEngine(self.model_cfg, self.device, dist.group.WORLD, **ref_kw)
            elif self.use_tp:
                ref_eng = TensorParallelLMEngine(self.model_cfg, self.device, self.pp_group, self.dp_group if self.dp > 1 else None, **ref_kw)
            else:
                ref_eng = get_engine_cls("language_model", "pp")(
                    self.model_cfg, self.device, self.pp_group, self.dp_group if self.dp > 1 else None,
                    virtual_stages=int(a.megatron.virtual_pipeline_model_parallel_size or 1), **ref_kw)
            self.ref = ActorWorker(ref_eng)
            self.ref.init_model()
        r = arr.rollout
        self.roll_w = TransformerWeights(self.model_cfg, device=self.device, dtype=eng.engine.dtype, init=False)
        self.rollout = RolloutEngine(self.model_cfg, self.device, max_num_seqs=r.max_num_seqs,
                                     max_model_len=r.max_model_len or (cfg.data.max_prompt_length + cfg.data.max_response_length),
                                     gpu_memory_utilization=r.gpu_memory_utilization, use_cuda_graph=r.cuda_graph,
                                     max_prefill_tokens=r.max_num_batched_tokens, seed=cfg.data.seed + 17 * self.rank)
        self._refresh_rollout_weights()
        mc = self.model_cfg
        return {"rank": self.rank, "vocab_size": mc.vocab_size, "eos_token_id": mc.eos_token_id, "pad_token_id": mc.pad_token_id,
                "num_params": mc.num_params(), "pp": 1 if (self.use_tp or self.use_ep) else self.pp, "tp": self.tp, "ep": self.world_size if self.use_ep else 1,
                "dp": self.dp,
                "stage_params": (eng.engine.num_local_params() if hasattr(eng.engine, "num_local_params")
                                 else sum(v.numel() for v in eng.engine.master.values()))}

    def _refresh_rollout_weights(self) -> None:
        full = self.actor.engine.engine.gather_full_state()              # every stage's fp32 masters, on every rank
        with torch.no_grad():
            self.roll_w.load_state_dict({k: v.to(self.device) for k, v in full.items()})
        self.rollout.set_weights(self.roll_w.embed, self.roll_w.layers, self.roll_w.head)

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def generate_sequences(self, prompts: DataProto) -> DataProto:
        from ..workers.hybrid_worker import ActorRolloutRefWorker

        # reuse the hybrid worker's request building / post-processing on this worker's engine
        shim = getattr(self, "_gen_shim", None)
        if shim is None:
            shim = self._gen_shim = object.__new__(ActorRolloutRefWorker)
            shim.config, shim.device, shim.rollout, shim.model_cfg = self.config, self.device, self.rollout, self.model_cfg
            shim.rank, shim.world, shim.rollout_tp, shim.tp_group, shim.tp_rank = self.rank, self.world_size, 1, None, 0
            shim.profiler = None
            shim._remote_client = None
        return shim.generate_sequences(prompts.to(self.device)).to("cpu")

    @register(dispatch_mode=make_nd_compute_dataproto_dispatch_fn("actor"))
    def compute_log_prob(self, data: DataProto) -> DataProto:
        data.meta_info.setdefault("temperature", self.config.actor_rollout_ref.rollout.temperature)
        return self.actor.compute_log_prob(data.to(self.device)).to("cpu")

    @register(dispatch_mode=make_nd_compute_dataproto_dispatch_fn("actor"))
    def compute_ref_log_prob(self, data: DataProto) -> DataProto:
        data.meta_info.setdefault("temperature", self.config.actor_rollout_ref.rollout.temperature)
        out = self.ref.compute_log_prob(data.to(self.device))
        return DataProto.from_dict({"ref_log_prob": out.batch["old_log_probs"]}).to("cpu")

    @register(dispatch_mode=make_nd_compute_dataproto_dispatch_fn("actor"))
    def update_actor(self, data: DataProto) -> DataProto:
        data.meta_info.setdefault("temperature", self.config.actor_rollout_ref.rollout.temperature)
        m = self.actor.update_actor(data.to(self.device))
        return DataProto(meta_info={"metrics": {k: [v] for k, v in m.items()}})

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def update_weight(self) -> bool:
        self._refresh_rollout_weights()
        return True

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def save_checkpoint(self, path: str, global_step: int = 0, max_ckpt_to_keep: int | None = None) -> None:
        self.actor.save_checkpoint(os.path.join(path, "actor"), global_step, max_ckpt_to_keep)

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def load_checkpoint(self, path: str) -> None:
        self.actor.load_checkpoint(os.path.join(path, "actor"))
        self._refresh_rollout_weights()

    @register(dispatch_mode=Dispatch.RANK_ZERO, execute_mode=Execute.RANK_ZERO)
    def weights_checksum(self) -> float:
        return float(sum(t.double().abs().sum() for _, t in self.roll_w.named_tensors()))

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def teardown(self) -> None:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()


def make_worker_group(config: PPOConfig, backend: str, world_size: int, env: dict | None = None):
    if config.actor_rollout_ref.actor.strategy == "megatron":
        if backend == "local":
            raise ValueError("the pipeline-parallel actor needs one process per stage (controller.backend=mp | ray)")
        cia = ClassWithInitArgs(PipelinePPOWorker, config)
        if backend == "ray":
            from ..controller.worker_group import RayWorkerGroup

            return RayWorkerGroup(world_size, cia, env=env)
        return MPWorkerGroup(world_size, cia, env=env)
    cia = ClassWithInitArgs(PPOWorker, config)
    if backend == "local":
        if world_size != 1:
            raise ValueError("the in-process backend runs one worker (collectives need one process per rank)")
        return LocalWorkerGroup(1, ClassWithInitArgs(PPOWorker, config, init_dist=False))
    if backend == "mp":
        return MPWorkerGroup(world_size, cia, env=env)
    if backend == "ray":
        from ..controller.worker_group import RayWorkerGroup

        return RayWorkerGroup(world_size, cia, env=env)
    raise ValueError(f"unknown controller backend {backend!r} (local | mp | ray)")


class ControllerPPOTrainer:
    """Driver-side loop.  ``reward_fn(batch) -> [B] scores or [B, R] token scores``; defaults to the synthetic token reward."""

    def __init__(self, config: PPOConfig, worker_group, train_dataset, val_dataset=None, reward_fn: Callable | None = None,
                 tracking=None, sampler=None):
        self.config, self.wg = config, worker_group
        self.train_dataset, self.val_dataset, self.reward_fn, self.tracking = train_dataset, val_dataset, reward_fn, tracking
        self.sampler = sampler
        self.kl_ctrl = (core_algos.get_kl_controller(config.algorithm.kl_ctrl) if config.algorithm.use_kl_in_reward else None)
        self.global_step = 0
        self.model_info = self.wg.init_model()[0]
        self.wg.start_worker_aliveness_check(1.0)

    # ---- data
    def _batches(self):
        bs = self.config.data.train_batch_size
        order = list(self.sampler) if self.sampler is not None else list(range(len(self.train_dataset)))
        for i in range(0, len(order) - bs + 1, bs):
            yield DataProto.from_single_dict(collate_fn([self.train_dataset[j] for j in order[i:i + bs]]))

    def _reward(self, batch: DataProto) -> torch.Tensor:
        mask = batch.batch["response_mask"]
        if self.reward_fn is not None:
            scores = self.reward_fn(batch)
            if torch.is_tensor(scores) and scores.dim() == 2:
                return scores
        elif "rm_scores" in batch.batch:
            return batch.batch["rm_scores"]                  # reward model only (no rule-based scorer configured)
        else:
            from ..data.synthetic import synthetic_token_reward

            resp, lens = batch.batch["responses"].numpy(), mask.sum(-1).numpy()
            rm = batch.non_tensor_batch["reward_model"]
            v = self.model_info["vocab_size"]
            scores = [synthetic_token_reward(resp[i, : lens[i]], rm[i]["ground_truth"], v) for i in range(len(batch))]
        return core.place_outcome_rewards(scores, mask)

    # ---- one PPO step = rollout_and_score -> compute_advantages -> train_on (recipes override the pieces)
    def rollout_and_score(self, prompts: DataProto, timing: dict) -> DataProto:
        """Generation + rewards for one batch of prompts (n samples each, whole groups adjacent)."""
        if "uid" not in prompts.non_tensor_batch:
            prompts.non_tensor_batch["uid"] = np.array([str(uuid.uuid4()) for _ in range(len(prompts))], dtype=object)
        with marked_timer("gen", timing):
            batch = self.wg.generate_sequences(prompts)
        with marked_timer("reward", timing):
            if self.config.reward_model.enable and hasattr(self.wg, "compute_rm_score"):
                batch.batch["rm_scores"] = self.wg.compute_rm_score(batch).batch["rm_scores"]      # model-based reward
            batch.batch["token_level_scores"] = self._reward(batch)
        return batch

    def compute_advantages(self, batch: DataProto, metrics: dict) -> DataProto:
        cfg, algo = self.config, self.config.algorithm
        if algo.use_kl_in_reward:
            batch, klm = core.apply_kl_penalty(batch, self.kl_ctrl, algo.kl_penalty)
            metrics.update(klm)
        else:
            batch.batch["token_level_rewards"] = batch.batch["token_level_scores"]
        batch = core.compute_advantage(batch, algo.adv_estimator, algo.gamma, algo.lam, cfg.actor_rollout_ref.rollout.n,
                                       algo.norm_adv_by_std_in_grpo, algo)
        if algo.use_pf_ppo:
            batch = core.pf_ppo_resample(batch, algo.pf_ppo.get("reweight_method", "pow"), float(algo.pf_ppo.get("weight_pow", 2.0)))
        return batch

    def train_on(self, batch: DataProto, timing: dict) -> dict:
        """old log-probs (+ ref, values), advantages, mini-batch updates, weight hand-off.  Returns the step metrics."""
        cfg, wg = self.config, self.wg
        n = cfg.actor_rollout_ref.rollout.n
        with marked_timer("old_log_prob", timing):
            batch.meta_info["temperature"] = cfg.actor_rollout_ref.rollout.temperature
            if cfg.algorithm.rollout_correction.bypass_mode and "rollout_log_probs" in batch.batch:
                batch.batch["old_log_probs"] = batch.batch["rollout_log_probs"].float()      # engine log-probs serve as pi_old
                entropy = float("nan")
            else:
                old = wg.compute_log_prob(batch)
                batch.batch["old_log_probs"] = old.batch["old_log_probs"]
                entropy = core_algos.masked_mean(old.batch["entropys"], batch.batch["response_mask"].float()).item()
        if need_reference_policy(cfg):
            with marked_timer("ref", timing):
                batch.batch["ref_log_prob"] = wg.compute_ref_log_prob(batch).batch["ref_log_prob"]
        if need_critic(cfg):
            with marked_timer("values", timing):
                batch.batch["values"] = wg.compute_values(batch).batch["values"]
        metrics: dict = {}
        with marked_timer("adv", timing):
            batch = self.compute_advantages(batch, metrics)
        # the driver balances nothing here: whole GRPO groups stay adjacent, chunks are contiguous slices (DP_COMPUTE_PROTO)
        upd: dict[str, list] = {}
        with marked_timer("update_actor", timing):
            mini = cfg.actor_rollout_ref.actor.ppo_mini_batch_size * n
            bounds = list(range(mini, len(batch) + 1, mini)) or [len(batch)]
            if bounds[-1] != len(batch):
                bounds.append(len(batch))
            start = 0
            for bi, end in enumerate(bounds):
                part = batch.slice(start, end)
                if len(part) % wg.world_size != 0:
                    raise ValueError(f"mini-batch of {len(part)} samples does not split over {wg.world_size} workers")
                part.meta_info = {**batch.meta_info, "is_opt_step": True, "is_lr_step": bi == len(bounds) - 1}
                if need_critic(cfg):
                    M.append_to_dict(upd, wg.update_critic(part).meta_info["metrics"])
                if not need_critic(cfg) or self.global_step >= cfg.trainer.critic_warmup:
                    M.append_to_dict(upd, wg.update_actor(part).meta_info["metrics"])
                start = end
        with marked_timer("update_weight", timing):
            wg.update_weight()
        self.global_step += 1
        ruf = cfg.trainer.ref_update_freq
        if need_reference_policy(cfg) and ruf and ruf > 0 and self.global_step % ruf == 0 and hasattr(wg, "update_reference_policy"):
            wg.update_reference_policy()
        mask = batch.batch["response_mask"].float()
        metrics.update({k: float(np.mean(v)) for k, v in upd.items()})
        metrics.update({
            "critic/score/mean": batch.batch["token_level_scores"].sum(-1).mean().item(),
            "critic/advantages/mean": core_algos.masked_mean(batch.batch["advantages"], mask).item(),
            "response_length/mean": mask.sum(-1).mean().item(), "actor/entropy_old": entropy,
            "perf/total_num_tokens": int(batch.batch["attention_mask"].sum().item()), "num_samples": len(batch),
        })
        return metrics

    def step(self, prompts: DataProto) -> dict:
        timing: dict[str, float] = {}
        t_step = time.time()
        batch = self.rollout_and_score(prompts, timing)
        metrics = self.train_on(batch, timing)
        timing["step"] = time.time() - t_step
        metrics.update({f"timing_s/{k}": v for k, v in timing.items()})
        return metrics

    def validate(self) -> dict:
        if self.val_dataset is None or len(self.val_dataset) == 0:
            return {}
        w = self.wg.world_size
        rows = [self.val_dataset[i] for i in range(len(self.val_dataset) - len(self.val_dataset) % w or len(self.val_dataset))]
        prompts = DataProto.from_single_dict(collate_fn(rows))
        prompts.meta_info.update({"validate": True, "do_sample": self.config.actor_rollout_ref.rollout.val_kwargs.get("do_sample", False), "n": 1})
        out = self.wg.generate_sequences(prompts)
        scores = self._reward(out).sum(-1)
        return {"val-core/score/mean@1": scores.mean().item(), "val-aux/response_length/mean": out.batch["response_mask"].sum(-1).float().mean().item()}

    def fit(self, max_steps: int | None = None) -> list[dict]:
        cfg = self.config
        total = max_steps or cfg.trainer.total_training_steps or 10 ** 9
        history: list[dict] = []
        if cfg.trainer.val_before_train and self.val_dataset is not None:
            v = self.validate()
            if self.tracking is not None:
                self.tracking.log(v, step=0)
        for epoch in range(max(1, cfg.trainer.total_epochs)):
            for prompts in self._batches():
                m = self.step(prompts)
                if cfg.trainer.test_freq > 0 and self.global_step % cfg.trainer.test_freq == 0:
                    m.update(self.validate())
                if cfg.trainer.save_freq > 0 and self.global_step % cfg.trainer.save_freq == 0:
                    path = os.path.join(cfg.trainer.default_local_dir, f"global_step_{self.global_step}")
                    self.wg.save_checkpoint(path, self.global_step, cfg.trainer.max_actor_ckpt_to_keep)
                if self.tracking is not None:
                    self.tracking.log(m, step=self.global_step)
                history.append(m)
                if self.global_step >= total:
                    return history
        return history

    def shutdown(self) -> None:
        try:
            self.wg.teardown()
        except Exception:
            pass
        self.wg.shutdown()


def main(argv: list[str] | None = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    backend = "mp"
    rest = []
    for a in argv:
        if a.startswith("controller.backend="):
            backend = a.split("=", 1)[1]
        else:
            rest.append(a)
    cfg = load_config(overrides=rest)
    world = max(1, cfg.trainer.n_gpus_per_node * cfg.trainer.nnodes)
    from ..models import get_model_config
    from .main_stream import build_datasets

    mc = get_model_config(cfg.actor_rollout_ref.model.path, **cfg.actor_rollout_ref.model.override_config)
    train, val, sampler, tok = build_datasets(cfg, mc)
    reward_fn = None
    if tok is not None:
        from ..data.reward_manager import load_reward_manager

        reward_fn = load_reward_manager(cfg, tok)
    from ..utils.tracking import Tracking

    tracking = Tracking(cfg.trainer.project_name, cfg.trainer.experiment_name, cfg.trainer.logger)
    wg = make_worker_group(cfg, backend, world)
    trainer = ControllerPPOTrainer(cfg, wg, train, val, reward_fn, tracking, sampler)
    try:
        trainer.fit()
    finally:
        trainer.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
