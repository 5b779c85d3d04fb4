"""Per-GPU workers: the hybrid actor / rollout / reference worker and the critic worker.

API surface mirrors the reference's stream workers (rlboost/verl_stream/workers/stream_fsdp_workers.py:
``init_model`` :262, ``update_actor_stream`` :378-431, ``update_weight_remote`` :435-444,
``generate_sequences_remote`` :448-495, ``get_stream_batches`` :497-507; critic :511-595) and verl's
``ActorRolloutRefWorker`` (3rdparty/verl/verl/workers/fsdp_workers.py:863-1108), but one worker = one
OS process = one GPU with everything native and in-process: ZeRO-3 engine, rollout engine sharing the
trainer's bf16 weights, and the peer-memory weight-transfer path instead of shm/TCP/ZMQ agents.
"""
from __future__ import annotations

import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from ..config.schema import PPOConfig
from ..models import get_model_config
from ..parallel.zero3 import Zero3Engine
from ..protocol import DataProto
from ..rollout.engine import RolloutEngine, SamplingParams
from .actor import ActorConfig as ActorRuntimeConfig
from .actor import DataParallelPPOActor, DataParallelPPOCritic


def get_response_mask(responses: torch.Tensor, eos_token_id: int, lengths: torch.Tensor | None = None) -> torch.Tensor:
    """1 up to and including the first EOS (reference: verl/utils/torch_functional.py:226)."""
    if lengths is not None:
        ar = torch.arange(responses.shape[1], device=responses.device)[None]
        return (ar < lengths[:, None]).long()
    is_eos = responses == eos_token_id
    after = (is_eos.cumsum(-1) - is_eos.long()) > 0
    return (~after).long()


def lr_at_step(optim, step: int, total_steps: int) -> float:
    """Warmup + constant / cosine schedule (reference: verl/utils/torch_functional.py:509-700)."""
    warm = optim.lr_warmup_steps
    if warm < 0:
        warm = int(optim.lr_warmup_steps_ratio * max(total_steps, 0))
    if warm > 0 and step < warm:
        return optim.lr * (step + 1) / warm
    if optim.warmup_style == "cosine" and total_steps > warm:
        prog = min(1.0, (step - warm) / max(1, total_steps - warm))
        cos = 0.5 * (1.0 + math.cos(math.pi * 2.0 * optim.num_cycles * prog))
        return optim.lr * (optim.min_lr_ratio + (1 - optim.min_lr_ratio) * cos)
    return optim.lr


def maybe_load_pretrained(engine, path: str, rank: int = 0) -> int:
    """``path`` = HF checkpoint directory with weights -> load them unit by unit (value-head engines keep their fresh head);
    presets / config-only directories keep the seeded random init."""
    from pathlib import Path

    p = Path(str(path))
    if not p.is_dir() or not (list(p.glob("*.safetensors")) or list(p.glob("pytorch_model*.bin"))):
        return 0
    from ..models.hf_io import load_into_engine

    n = load_into_engine(engine, p)
    if rank == 0:
        print(f"[worker] loaded {n} tensors from {p}", flush=True)
    return n

from ..utils.profiler import DistProfiler  # noqa: E402


def auto_fsdp_size(configured: int, world: int, comm: str, rank: int = 0, local_world: int | None = None) -> int:
    """Shard-group size of the ZeRO engine.  An explicit ``fsdp_config.fsdp_size`` wins.  Otherwise a job that spans several nodes
    (torchrun exports LOCAL_WORLD_SIZE < WORLD_SIZE) with the peer-memory backend shards INSIDE each node — the symmetric heap and
    its NVLink kernels only exist within one NVSwitch domain — and replicates across nodes, where the gradient all-reduce of the
    HSDP replica group runs over NCCL / the network (reference: fsdp_size / device-mesh (ddp, fsdp), fsdp_workers.py:99-118)."""
    if configured and configured > 0:
        return configured
    lw = local_world if local_world is not None else int(os.environ.get("LOCAL_WORLD_SIZE", world) or world)
    if comm == "p2p" and 0 < lw < world and world % lw == 0:
        if rank == 0:
            print(f"[worker] {world} ranks over {world // lw} nodes: ZeRO shards inside each node ({lw} GPUs, peer memory), "
                  f"replicas across nodes (NCCL)", flush=True)
        return lw
    return configured


class ActorRolloutRefWorker:
    """role in {"actor_rollout", "actor_rollout_ref", "actor", "rollout", "ref"}."""

    def __init__(self, config: PPOConfig, role: str = "actor_rollout_ref", device=None, group=None):
        self.config = config
        self.role = role
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if device is None:
            device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device)
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        self.has_actor = "actor" in role
        self.has_rollout = "rollout" in role
        self.has_ref = "ref" in role
        self.actor_engine: Zero3Engine | None = None
        self.ref_engine: Zero3Engine | None = None
        self.rollout: RolloutEngine | None = None
        self.actor: DataParallelPPOActor | None = None
        self.lr_step = 0
        self.weight_version = 0
        self._rollout_full: dict | None = None
        self.profiler = None
        self.sp_size, self.dp_size, self.dp_rank, self.sp_group = 1, self.world, self.rank, None

    # ------------------------------------------------------------------ init
    def init_model(self) -> None:
        cfg = self.config
        arr = cfg.actor_rollout_ref
        from ..utils.fs import is_non_local, stage_path
        from ..utils.import_utils import import_external_libs

        import_external_libs(getattr(arr.model, "external_lib", None))     # user registrations (presets, reward functions, tools) as an import side effect

        if is_non_local(arr.model.path) or getattr(arr.model, "use_shm", False):      # hdfs:// checkpoints are fetched once per node
            arr.model.path = stage_path(arr.model.path, use_shm=bool(getattr(arr.model, "use_shm", False)))
        self.model_cfg = get_model_config(arr.model.path, **arr.model.override_config)
        comm = arr.actor.fsdp_config.comm_backend
        opt = arr.actor.optim
        seed = cfg.data.seed
        # Ulysses SP: consecutive ranks form an SP group; data parallelism is over the remaining dimension
        sp = max(1, arr.actor.ulysses_sequence_parallel_size)
        if self.world % sp != 0:
            raise ValueError(f"world size {self.world} not divisible by ulysses_sequence_parallel_size {sp}")
        self.sp_size, self.dp_size, self.dp_rank = sp, self.world // sp, self.rank // sp
        self.sp_group = None
        if sp > 1:
            base = dist.get_process_group_ranks(self.group) if self.group is not None else list(range(self.world))
            for g in range(self.dp_size):       # every process creates every group, in the same order
                grp = dist.new_group([base[g * sp + i] for i in range(sp)])
                if g == self.dp_rank:
                    self.sp_group = grp
        # HSDP: fsdp_size < world => shard inside consecutive blocks of fsdp_size ranks, replicate across blocks
        fs = auto_fsdp_size(arr.actor.fsdp_config.fsdp_size, self.world, comm, self.rank)
        self.shard_group, self.replica_group = self.group, None
        if dist.is_initialized() and 0 < fs < self.world:
            if self.world % fs != 0:
                raise ValueError(f"world size {self.world} not divisible by fsdp_size {fs}")
            base = dist.get_process_group_ranks(self.group) if self.group is not None else list(range(self.world))
            for b in range(self.world // fs):
                grp = dist.new_group(base[b * fs:(b + 1) * fs])
                if b == self.rank // fs:
                    self.shard_group = grp
            for i in range(fs):
                grp = dist.new_group(base[i::fs])
                if i == self.rank % fs:
                    self.replica_group = grp
        from ..models.lora import LoraConfig

        m = arr.model
        self.lora_cfg = LoraConfig.from_model_cfg(m.lora_rank, m.lora_alpha, m.target_modules, m.exclude_modules)
        if self.has_actor:
            self.actor_engine = Zero3Engine(self.model_cfg, self.device, group=self.shard_group, comm=comm, trainable=True,
                                            seed=seed, lr=opt.lr, betas=tuple(opt.betas), eps=opt.eps,
                                            weight_decay=opt.weight_decay, max_grad_norm=arr.actor.grad_clip,
                                            sp_group=self.sp_group, replica_group=self.replica_group, lora=self.lora_cfg,
                                            freeze_vision_tower=m.freeze_vision_tower)
            self._maybe_load_pretrained(self.actor_engine, arr.model.path)
            if m.lora_adapter_path and self.actor_engine.lora is not None:
                from ..models.lora import load_peft_adapter

                n_ad = load_peft_adapter(self.actor_engine, m.lora_adapter_path)
                if self.rank == 0:
                    print(f"[worker] loaded {n_ad} adapter tensors from {m.lora_adapter_path}", flush=True)
            fc = arr.actor.fsdp_config
            eng = self.actor_engine
            want = str(fc.resident_params).lower()
            replica_bytes = 2 * eng.num_params()
            hbm = torch.cuda.get_device_properties(self.device).total_memory if self.device.type == "cuda" else 1 << 62
            if eng.world > 1 and (want == "true" or (want == "auto" and replica_bytes <= hbm // 4)):
                eng.make_resident()
            eng.act_ckpt = str(fc.activation_checkpoint) if arr.model.enable_gradient_checkpointing else "none"
            eng.act_reserve_bytes = int(fc.activation_reserve_gb * (1 << 30))
            eng.act_offload = bool(fc.activation_offload)
            eng.moe_aux = bool(self.model_cfg.num_experts) and arr.actor.router_aux_loss_coef > 0
            if self.rank == 0 and self.device.type == "cuda":
                try:        # what the layout decisions will look like on this box (utils/hbm_sizing.py uses the same arithmetic)
                    from ..utils.hbm_sizing import plan

                    p = plan(self.model_cfg, n_gpus=eng.world, rollout_tp=max(1, arr.rollout.tensor_model_parallel_size),
                             tokens_per_microbatch=arr.actor.ppo_max_token_len_per_gpu, max_num_seqs=arr.rollout.max_num_seqs,
                             max_model_len=cfg.data.max_prompt_length + cfg.data.max_response_length, hbm_gb=hbm / (1 << 30),
                             reserve_gb=fc.activation_reserve_gb)
                    print(f"[worker] HBM plan: {p.decisions}", flush=True)
                except Exception as e:  # the plan is informational only
                    print(f"[worker] HBM plan unavailable: {e!r}", flush=True)
            n_dp = self.dp_size
            a = arr.actor
            mini = max(1, a.ppo_mini_batch_size * arr.rollout.n // n_dp)      # per-GPU samples per optimizer step
            micro = a.ppo_micro_batch_size_per_gpu or (max(1, a.ppo_micro_batch_size // n_dp) if a.ppo_micro_batch_size else None)
            rt = ActorRuntimeConfig(ppo_mini_batch_size=mini, ppo_micro_batch_size_per_gpu=micro,
                                    use_dynamic_bsz=a.use_dynamic_bsz, ppo_max_token_len_per_gpu=a.ppo_max_token_len_per_gpu,
                                    clip_ratio=a.clip_ratio, clip_ratio_low=a.clip_ratio_low, clip_ratio_high=a.clip_ratio_high,
                                    clip_ratio_c=a.clip_ratio_c, loss_agg_mode=a.loss_agg_mode, entropy_coeff=a.entropy_coeff, router_aux_loss_coef=a.router_aux_loss_coef,
                                    use_kl_loss=a.use_kl_loss, kl_loss_coef=a.kl_loss_coef, kl_loss_type=a.kl_loss_type,
                                    ppo_epochs=a.ppo_epochs, policy_loss_mode=a.policy_loss.loss_mode,
                                    policy_loss=a.policy_loss.to_dict(), grad_clip=a.grad_clip)
            if a.policy_loss.loss_mode == "flowrl":
                from .actor import FlowRLActor

                self.actor = FlowRLActor(self.actor_engine, rt, dp_group=self.group, seed=seed)
            else:
                self.actor = DataParallelPPOActor(self.actor_engine, rt, dp_group=self.group)
        if self.has_ref:
            # reference policy = frozen copy of the initial actor weights (same seed => identical init)
            self.ref_engine = Zero3Engine(self.model_cfg, self.device, group=self.shard_group, comm=comm, trainable=False,
                                          seed=seed, heap=self.actor_engine.heap if self.actor_engine else None,
                                          sp_group=self.sp_group)
            self._maybe_load_pretrained(self.ref_engine, arr.model.path)
            self.ref_actor = DataParallelPPOActor(self.ref_engine, ActorRuntimeConfig(), dp_group=self.group)
        if self.has_rollout:
            r = arr.rollout
            max_len = r.max_model_len or (cfg.data.max_prompt_length + cfg.data.max_response_length)
            # tensor-parallel rollout (reference: rollout.tensor_model_parallel_size -> SGLang tp_size, mesh (dp, infer_tp),
            # collect from tp-rank 0; stream_fsdp_workers.py:126-153): consecutive ranks form one engine replica that decodes
            # the union of its ranks' prompts in lock-step; every rank keeps its own slice of the results.
            self.rollout_tp = max(1, r.tensor_model_parallel_size)
            self.tp_group, self.tp_rank, tp_ctx = None, 0, None
            if self.rollout_tp > 1:
                if self.world % self.rollout_tp != 0:
                    raise ValueError(f"world size {self.world} not divisible by rollout TP {self.rollout_tp}")
                from ..rollout.tp import TPContext

                base = dist.get_process_group_ranks(self.group) if self.group is not None else list(range(self.world))
                for g in range(self.world // self.rollout_tp):
                    grp = dist.new_group(base[g * self.rollout_tp:(g + 1) * self.rollout_tp])
                    if g == self.rank // self.rollout_tp:
                        self.tp_group = grp
                self.tp_rank = self.rank % self.rollout_tp
                heap = None
                if self.device.type == "cuda":
                    from ..parallel.symm import SymmetricHeap, make_heap

                    heap = make_heap(group=self.tp_group, device=self.device)
                tp_ctx = TPContext(self.model_cfg, heap, self.tp_rank, self.rollout_tp,
                                   max_batch=r.max_num_seqs * self.rollout_tp, group=self.tp_group)
            # a TP replica serves the prompts of all its ranks: tp x the per-GPU sequence slots (each holds 1/tp of the KV heads)
            self.rollout = RolloutEngine(self.model_cfg, self.device, max_num_seqs=r.max_num_seqs * self.rollout_tp, max_model_len=max_len,
                                         gpu_memory_utilization=r.gpu_memory_utilization, use_cuda_graph=r.cuda_graph,
                                         max_prefill_tokens=r.max_num_batched_tokens,
                                         seed=seed + 17 * (self.rank // self.rollout_tp), tp=tp_ctx,
                                         weight_dtype=r.weight_dtype,
                                         enable_prefix_cache=(r.enable_prefix_caching or r.multi_turn.enable) and tp_ctx is None,
                                         kv_vmm=True if r.kv_vmm else None)
            if self.has_actor:
                self.update_weight_remote()
            if r.skip_rollout:
                # debugging aid (verl rollout.skip_rollout / skip_dump_dir): generations are dumped once and replayed afterwards,
                # so the training half of a step can be iterated on a fixed batch without running the engine
                from ..utils.rollout_skip import RolloutSkip

                self._rollout_skip = RolloutSkip(os.path.join(r.skip_dump_dir, f"rank{self.rank}"), cfg.trainer.experiment_name, r.n,
                                                 cfg.data.train_batch_size)
                self._rollout_skip.wrap_generate_sequences(self)

    def _maybe_load_pretrained(self, engine, path: str) -> None:
        maybe_load_pretrained(engine, path, self.rank)

    # ------------------------------------------------------------------ weight sync (colocated)
    @DistProfiler.annotate("update_weight_remote")
    def update_weight_remote(self) -> bool:
        """Publish the current actor weights to the rollout engine.  world==1: the engine *aliases* the
        trainer's bf16 shard (zero copies).  world>1: every rank gathers a full bf16 replica into
        engine-owned buffers with the peer-memory all-gather (or NCCL in baseline mode)."""
        eng = self.actor_engine
        if eng is None or self.rollout is None:
            return False
        self.weight_version += 1
        base = eng.base_units
        if eng.vision is not None:
            self._publish_vision_tower(eng)
        if eng.lora is not None:
            # adapters are merged into an engine-owned replica (reference: layered summon + merge, fsdp_utils.py:569-694)
            from ..models.lora import merge_layer_

            if self._rollout_full is None:
                self._rollout_full = {u.name: torch.empty(u.full_numel, device=self.device, dtype=eng.param_dtype) for u in base}
            views = []
            for u in base:
                full = u.acquire()
                self._rollout_full[u.name].copy_(full)
                u.release()
                views.append(u.views(self._rollout_full[u.name]))
            for i, lu in enumerate(eng.lora_layers):
                lfull = lu.acquire()
                merge_layer_(self.model_cfg, views[1 + i], lu.views(lfull), eng.lora.scale)
                lu.release()
        elif eng.world == 1:
            views = [u.views(u.shard) for u in base]
        elif eng.resident:
            eng.ensure_resident()           # one gather per optimizer step; the engine aliases the trainer's replica
            views = [u.views(u._resident) for u in base]
        elif getattr(self, "rollout_tp", 1) > 1:
            # ZeRO-3 (no resident replica: 70B-class) + TP rollout: gather ONE block at a time, keep only this rank's TP slice.
            # Peak extra memory = one full block + 1/tp of the model (a full replica per GPU would not fit in HBM).
            tp = self.rollout.tp
            keep = self._rollout_full if isinstance(self._rollout_full, dict) else {}

            def stash(name: str, tensors: dict) -> dict:
                out = {}
                for k, v in tensors.items():
                    dst = keep.get((name, k))
                    if dst is None or dst.shape != v.shape:
                        dst = keep[(name, k)] = torch.empty(v.shape, device=self.device, dtype=v.dtype)
                    dst.copy_(v)
                    out[k] = dst
                return out

            sliced = []
            for i, u in enumerate(base):
                full = u.acquire()
                v = u.views(full)
                sliced.append(stash(u.name, tp.shard_layer(v) if 0 < i < len(base) - 1 else v))
                u.release()
            self._rollout_full = keep
            self.rollout.set_weights(sliced[0], sliced[1:-1], sliced[-1], version=self.weight_version, presharded=True)
            return True
        else:
            if self._rollout_full is None:
                self._rollout_full = {u.name: torch.empty(u.full_numel, device=self.device, dtype=eng.param_dtype)
                                      for u in base}
            if eng.heap is not None:
                eng.heap.barrier()
            for u in base:
                u._gather_into(self._rollout_full[u.name])
            views = [u.views(self._rollout_full[u.name]) for u in base]
        self.rollout.set_weights(views[0], views[1:-1], views[-1], version=self.weight_version)
        return True

    def _publish_vision_tower(self, eng) -> None:
        """Vision-language actors: the engine runs the tower at prefill.  world == 1 aliases the trainer's shard, otherwise the tower
        (small next to the text stack) is gathered into an engine-owned replica."""
        from ..models.vision import tower_view

        u = eng.vision
        if eng.world == 1:
            flat = u.shard
        elif eng.resident:
            eng.ensure_resident()
            flat = u._resident
        else:
            if getattr(self, "_vision_full", None) is None:
                self._vision_full = torch.empty(u.full_numel, device=self.device, dtype=eng.param_dtype)
            if eng.heap is not None:
                eng.heap.barrier()
            u._gather_into(self._vision_full)
            flat = self._vision_full
        self.rollout.set_vision_weights(tower_view(self.model_cfg, u.views(flat)))

    # ------------------------------------------------------------------ local instance (disaggregated hybrid mode)
    def start_local_server(self, manager_endpoint: str, port: int) -> None:
        """Serve this rank's colocated engine to the rollout manager as a LOCAL instance (time-boxed by the manager)."""
        from ..rollout.server import EngineServer

        self.local_server = EngineServer(self.rollout, port=port).start()
        self.local_server.register_with_manager(manager_endpoint, local=True)
        self._in_rollout_mode = True          # the engine owns its KV pool until the first trainer_mode()

    def rollout_mode(self) -> None:
        """Give HBM to the local engine (resume its KV pool) — reference: fsdp_workers.py:651-750 rollout_mode()."""
        srv = getattr(self, "local_server", None)
        if srv is not None and not self._in_rollout_mode:
            srv._command("resume", {"tags": ["kv_cache"]}, wait=True)
            self._in_rollout_mode = True

    def trainer_mode(self) -> None:
        """Abort whatever the local engine still runs (the manager already re-dispatched it) and take the HBM back."""
        srv = getattr(self, "local_server", None)
        if srv is not None and self._in_rollout_mode:
            srv._command("release", {"tags": ["kv_cache"]}, wait=True)
            self._in_rollout_mode = False

    # ------------------------------------------------------------------ rollout
    @DistProfiler.annotate("generate_sequences")
    @torch.no_grad()
    def generate_sequences(self, prompts: DataProto) -> DataProto:
        """Generate ``n`` responses per prompt with the colocated engine (sync path / validation)."""
        cfg = self.config
        r = cfg.actor_rollout_ref.rollout
        meta = prompts.meta_info
        do_sample = meta.get("do_sample", r.do_sample)
        validate = meta.get("validate", False)
        n = 1 if validate else meta.get("n", r.n)
        temperature = (r.val_kwargs.get("temperature", 0.0) if validate else meta.get("temperature", r.temperature))
        if not do_sample:
            temperature = 0.0
        resp_len = meta.get("response_length", cfg.data.max_response_length)
        sp = SamplingParams(max_new_tokens=resp_len, temperature=temperature, top_p=meta.get("top_p", r.top_p),
                            top_k=meta.get("top_k", r.top_k), n=n, ignore_eos=meta.get("ignore_eos", r.ignore_eos))
        raw = prompts.non_tensor_batch.get("raw_prompt_ids")
        if raw is None:
            ids, mask = prompts.batch["input_ids"], prompts.batch["attention_mask"].bool()
            raw = [ids[i][mask[i]].tolist() for i in range(len(prompts))]
        t0 = time.time()
        self.rollout.resume_memory(("kv_cache",))
        if r.multi_turn.enable:
            out = self._generate_multi_turn(prompts, raw, sp, n, resp_len)
        else:
            mine = [list(map(int, p)) for p in raw]
            mm = None
            if "multi_modal_inputs" in prompts.non_tensor_batch:
                mm = [None if m is None else {"pixel_values": m["pixel_values"], "grid_thw": torch.as_tensor(m["image_grid_thw"]).tolist(),
                                              "second_per_grid_ts": m.get("second_per_grid_ts")}
                      for m in prompts.non_tensor_batch["multi_modal_inputs"]]
            if getattr(self, "rollout_tp", 1) > 1:
                # every TP rank must feed the SAME request stream (prompts and their images) to its engine shard
                parts: list = [None] * self.rollout_tp
                dist.all_gather_object(parts, (mine, mm), group=self.tp_group)
                off = sum(len(p[0]) for p in parts[: self.tp_rank])
                all_p = [p for part in parts for p in part[0]]
                all_mm = None
                if any(part[1] is not None for part in parts):
                    all_mm = [m for part in parts for m in (part[1] if part[1] is not None else [None] * len(part[0]))]
                groups = self.rollout.generate(all_p, sp, mm=all_mm)[off:off + len(mine)]
            else:
                groups = self.rollout.generate(mine, sp, mm=mm)
            out = self._postprocess(prompts, groups, n, resp_len)
        gen_s = time.time() - t0
        out.meta_info["timing"] = {"generate_sequences": gen_s}
        if r.free_cache_engine and not meta.get("keep_cache", False):
            self.rollout.release_memory(("kv_cache",))
        return out

    def _generate_multi_turn(self, prompts: DataProto, raw, sp: SamplingParams, n: int, resp_len: int) -> DataProto:
        """Tool / interaction trajectories inside the engine's continuous batch (reference: request-level path of
        sglang_rollout.py:807-1497 and experimental/agent_loop/tool_agent_loop.py)."""
        from ..agent import MultiTurnRollout, load_tools
        from ..agent.agent_loop import outputs_to_dataproto
        from ..agent.interactions import load_interactions
        from ..data.rl_dataset import load_tokenizer

        mt = self.config.actor_rollout_ref.rollout.multi_turn
        if getattr(self, "_multi_turn", None) is None:
            tok = load_tokenizer(self.config.actor_rollout_ref.model.tokenizer_path)
            from ..agent import get_agent_loop_cls

            agent_cfg = dict(self.config.actor_rollout_ref.rollout.agent or {})
            name = agent_cfg.get("default_agent_loop")
            if name in ("collabllm_agent",):
                from ..recipes import collabllm  # noqa: F401 — registers the simulated-user interaction
            loop_cls = get_agent_loop_cls(name) if name else MultiTurnRollout
            self._multi_turn = loop_cls(
                self.rollout, tok, **dict(agent_cfg.get("agent_loop_kwargs") or {}),
                tools=load_tools(mt.tool_config_path), interactions=load_interactions(mt.interaction_config_path),
                parser=mt.format, response_length=resp_len, max_assistant_turns=mt.max_assistant_turns,
                max_user_turns=mt.max_user_turns, max_parallel_calls=mt.max_parallel_calls,
                max_tool_response_length=mt.max_tool_response_length, tool_response_truncate_side=mt.tool_response_truncate_side,
                model_cfg=self.model_cfg, image_min_pixels=self.config.data.image_min_pixels, image_max_pixels=self.config.data.image_max_pixels)
        self._multi_turn.response_length = resp_len
        nt = prompts.non_tensor_batch
        samples = []
        for i, p in enumerate(raw):
            s = {"prompt_ids": list(map(int, p))}
            for key in ("tools_kwargs", "interaction_kwargs"):
                if key in nt and nt[key][i] is not None:
                    s[key] = nt[key][i]
            if "raw_prompt" in nt:
                s["messages"] = list(nt["raw_prompt"][i])
            if "multi_modal_inputs" in nt and nt["multi_modal_inputs"][i] is not None:
                s["multi_modal_inputs"] = nt["multi_modal_inputs"][i]
            samples += [dict(s) for _ in range(n)]
        outs = self._multi_turn.generate(samples, {"temperature": sp.temperature, "top_p": sp.top_p, "top_k": sp.top_k})
        rep = prompts.repeat(n, interleave=True) if n > 1 else prompts
        dp = outputs_to_dataproto(outs, rep.batch["input_ids"].shape[1], resp_len, self.model_cfg.pad_token_id, model_cfg=self.model_cfg)
        dev = rep.batch["input_ids"].device
        for k, v in dp.batch.items():
            rep.batch[k] = v.to(dev)
        rep.non_tensor_batch.update(dp.non_tensor_batch)
        return rep

    def _postprocess(self, prompts: DataProto, groups, n: int, resp_len: int) -> DataProto:
        """Pad responses to ``response_length`` and build input_ids / masks / position ids (reference:
        sglang_rollout_remote.py:318-391)."""
        pad = self.model_cfg.pad_token_id
        dev = prompts.batch["input_ids"].device
        rep = prompts.repeat(n, interleave=True) if n > 1 else prompts
        b = len(rep)
        resp = torch.full((b, resp_len), pad, dtype=torch.long)
        lps = torch.zeros(b, resp_len, dtype=torch.float32)
        lens = torch.zeros(b, dtype=torch.long)
        aborted = np.zeros(b, dtype=bool)
        k = 0
        for g in groups:
            for req in g:
                m = min(len(req.output_ids), resp_len)
                resp[k, :m] = torch.tensor(req.output_ids[:m], dtype=torch.long)
                lps[k, :m] = torch.tensor(req.logprobs[:m])
                lens[k] = m
                aborted[k] = req.finish_reason == "abort"
                k += 1
        resp, lps, lens = resp.to(dev), lps.to(dev), lens.to(dev)
        p_ids, p_mask = rep.batch["input_ids"], rep.batch["attention_mask"]
        response_mask = get_response_mask(resp, self.model_cfg.eos_token_id, lens)
        input_ids = torch.cat([p_ids, resp], dim=1)
        attention_mask = torch.cat([p_mask, response_mask], dim=1)
        p_pos = rep.batch.get("position_ids")
        if p_pos is None:
            p_pos = (p_mask.cumsum(-1) - 1).clamp_min(0)
        delta = torch.arange(1, resp_len + 1, device=dev)[None]
        if p_pos.dim() == 3:                  # M-RoPE [b, 3, S]: generated tokens continue all three axes after the largest position
            nxt = p_pos.amax(dim=(1, 2)).view(-1, 1, 1) + delta.view(1, 1, -1)
            position_ids = torch.cat([p_pos, nxt.expand(-1, 3, -1)], dim=2)
        else:
            position_ids = torch.cat([p_pos, p_pos[:, -1:] + delta], dim=1)
        rep.batch["prompts"] = p_ids
        rep.batch["responses"] = resp
        rep.batch["input_ids"] = input_ids
        rep.batch["attention_mask"] = attention_mask
        rep.batch["position_ids"] = position_ids
        rep.batch["response_mask"] = response_mask
        rep.batch["rollout_log_probs"] = lps
        rep.non_tensor_batch["aborted"] = aborted
        return rep

    # ------------------------------------------------------------------ streaming rollout through the manager
    def attach_rollout_client(self, client) -> None:
        """Rank 0 holds the manager connection (``rollout.remote_client.RemoteRolloutClient``)."""
        self.rollout_client = client
        self._stream_iter = None

    def generate_sequences_remote(self, prompts: DataProto) -> bool:
        """Submit the (unsharded) prompt batch to the rollout manager; returns once the request is in flight.
        Results are consumed with ``get_stream_batches`` (reference API: stream_fsdp_workers.py:448-507)."""
        if getattr(self, "rollout_client", None) is not None and self.rank == 0:
            r = self.config.actor_rollout_ref.rollout
            self._stream_iter = iter(self.rollout_client.generate_sequences_remote(
                prompts, {"temperature": r.temperature, "top_p": r.top_p, "top_k": r.top_k}))
        return True

    def get_stream_batches(self) -> DataProto | None:
        """Next streamed micro-batch (rank 0) or ``None`` at the end of the stream."""
        it = getattr(self, "_stream_iter", None)
        if it is None:
            return None
        try:
            return next(it)
        except StopIteration:
            self._stream_iter = None
            return None

    # ------------------------------------------------------------------ log-probs
    @DistProfiler.annotate("compute_log_prob")
    @torch.no_grad()
    def compute_log_prob(self, data: DataProto) -> DataProto:
        r = self.config.actor_rollout_ref.rollout
        data.meta_info.setdefault("micro_batch_size", r.log_prob_micro_batch_size_per_gpu)
        data.meta_info.setdefault("max_token_len", r.log_prob_max_token_len_per_gpu)
        data.meta_info.setdefault("use_dynamic_bsz", r.log_prob_use_dynamic_bsz)
        data.meta_info.setdefault("temperature", r.temperature)
        lp, ent = self.actor.compute_log_prob(data)
        return DataProto.from_dict({"old_log_probs": lp, "entropys": ent}, meta_info={"temperature": r.temperature})

    @torch.no_grad()
    def update_reference_policy(self) -> None:
        """Copy the current actor weights into the reference policy (``trainer.ref_update_freq``: self-play recipes such as SPIN move
        the reference every k steps; reference: recipe/spin/spin_trainer.py ref_update_freq)."""
        if self.ref_engine is None or self.actor_engine is None:
            return
        for src, dst in zip(self.actor_engine.base_units, self.ref_engine.base_units):
            dst.shard.copy_(src.shard)
            dst.drop_gathered()
        if getattr(self.actor_engine, "vision", None) is not None and getattr(self.ref_engine, "vision", None) is not None:
            self.ref_engine.vision.shard.copy_(self.actor_engine.vision.shard)
            self.ref_engine.vision.drop_gathered()
        self.ref_engine._resident_stale = self.ref_engine._reload_all = True

    @DistProfiler.annotate("compute_ref_log_prob")
    @torch.no_grad()
    def compute_ref_log_prob(self, data: DataProto) -> DataProto:
        ref = self.config.actor_rollout_ref.ref
        data.meta_info["micro_batch_size"] = ref.log_prob_micro_batch_size_per_gpu
        data.meta_info["max_token_len"] = ref.log_prob_max_token_len_per_gpu
        data.meta_info["use_dynamic_bsz"] = ref.log_prob_use_dynamic_bsz
        data.meta_info.setdefault("temperature", self.config.actor_rollout_ref.rollout.temperature)
        lp, _ = self.ref_actor.compute_log_prob(data, calculate_entropy=False)
        return DataProto.from_dict({"ref_log_prob": lp})

    # ------------------------------------------------------------------ update
    @DistProfiler.annotate("update_actor_stream")
    def update_actor_stream(self, data: DataProto) -> DataProto:
        meta = data.meta_info
        is_opt = meta.get("is_opt_step", True)
        is_lr = meta.get("is_lr_step", is_opt)
        opt = self.config.actor_rollout_ref.actor.optim
        total = opt.total_training_steps if opt.total_training_steps > 0 else (self.config.trainer.total_training_steps or 0)
        lr = lr_at_step(opt, self.lr_step, total)
        t0 = time.time()
        offload = self.config.actor_rollout_ref.actor.fsdp_config.optimizer_offload
        if offload:
            self.actor_engine.load_optimizer()
        metrics = self.actor.update_policy_stream(data, is_opt_step=is_opt, lr=lr)
        if offload and is_opt:
            self.actor_engine.offload_optimizer()
        metrics["actor/lr"] = [lr]
        # MFU of this update (fwd + recompute + bwd counted as the standard 6·N·T model FLOPs)
        from ..utils.flops_counter import FlopsCounter

        seqlens = data.batch["attention_mask"].sum(-1).tolist()
        achieved, peak = FlopsCounter(self.model_cfg).estimate_flops([int(x) for x in seqlens], time.time() - t0)
        metrics["perf/mfu/actor"] = [achieved / peak if peak != float("inf") else 0.0]
        if is_lr:
            self.lr_step += 1
        metrics["perf/update_actor_s"] = [time.time() - t0]
        if self.device.type == "cuda":
            metrics["perf/max_memory_allocated_gb"] = [torch.cuda.max_memory_allocated(self.device) / 2 ** 30]
            metrics["perf/max_memory_reserved_gb"] = [torch.cuda.max_memory_reserved(self.device) / 2 ** 30]
        return DataProto(meta_info={"metrics": metrics})

    update_actor = update_actor_stream

    def optimizer_step_only(self, is_lr_step: bool = True) -> dict:
        """Optimizer step on the gradients accumulated so far, with the SCHEDULED learning rate — used when a mini-batch
        boundary coincides with the end of a streamed slice, so there is no micro-batch left to ride the step on."""
        opt = self.config.actor_rollout_ref.actor.optim
        total = opt.total_training_steps if opt.total_training_steps > 0 else (self.config.trainer.total_training_steps or 0)
        lr = lr_at_step(opt, self.lr_step, total)
        offload = self.config.actor_rollout_ref.actor.fsdp_config.optimizer_offload
        if offload:
            self.actor_engine.load_optimizer()
        st = self.actor_engine.step(lr)
        if offload:
            self.actor_engine.offload_optimizer()
        if is_lr_step:
            self.lr_step += 1
        return {"actor/grad_norm": [st["grad_norm"]], "actor/lr": [lr]}

    # ------------------------------------------------------------------ checkpoint / profile
    def save_checkpoint(self, local_path: str, hdfs_path: str | None = None, global_step: int = 0,
                        max_ckpt_to_keep: int | None = None) -> None:
        extra = {"lr_step": self.lr_step}
        if hasattr(self.actor, "extra_state"):
            extra["actor_extra"] = self.actor.extra_state()          # e.g. FlowRL's log-Z head
        self._ckpt_manager().save(local_path, global_step, max_ckpt_to_keep, extra=extra)

    def _ckpt_manager(self):
        """ONE manager per worker, configured from ``actor.checkpoint.{save,load}_contents`` (e.g. ``hf_model``)."""
        if getattr(self, "_ckpt_mgr", None) is None:
            from ..utils.checkpoint import ShardedCheckpointManager

            ck = self.config.actor_rollout_ref.actor.checkpoint
            self._ckpt_mgr = ShardedCheckpointManager(self.actor_engine, self.model_cfg, tuple(ck.save_contents),
                                                      tuple(ck.load_contents), async_save=ck.async_save)
        return self._ckpt_mgr

    def load_checkpoint(self, local_path: str | None, hdfs_path: str | None = None, del_local_after_load: bool = False) -> None:
        if not local_path:
            return
        extra = self._ckpt_manager().load(local_path)
        self.lr_step = (extra or {}).get("lr_step", 0)
        if hasattr(self.actor, "load_extra_state"):
            self.actor.load_extra_state((extra or {}).get("actor_extra"))
        if self.rollout is not None:
            self.update_weight_remote()

    def _ensure_profiler(self):
        from ..utils.profiler import DistProfiler

        if self.profiler is None:
            role_cfg = self.config.actor_rollout_ref.actor.profiler
            # a role config that was never enabled means "no per-role filter": the trainer's global_profiler.steps decide
            self.profiler = DistProfiler(self.rank, self.config.global_profiler, role_cfg if role_cfg.enable else None, role=self.role)
            if self.actor is not None:
                self.actor.profiler = self.profiler
        return self.profiler

    def start_profile(self, **kw) -> None:
        self._ensure_profiler()
        self.profiler.start(**kw)

    def stop_profile(self) -> None:
        if self.profiler is not None:
            self.profiler.stop()

    def dump_memory_snapshot(self, tag: str = "manual", sub_dir: str | None = None) -> None:
        from ..utils.profiler import dump_memory_snapshot

        dump_memory_snapshot(self.config.global_profiler.save_path, tag, sub_dir, self.rank)


class RewardModelWorker:
    """Model-based reward: a value-head transformer scores each (prompt, response); the scalar at the last valid response
    token becomes ``rm_scores`` (reference: verl/workers/fsdp_workers.py:1573-1925 ``RewardModelWorker.compute_rm_score``)."""

    def __init__(self, config: PPOConfig, device=None, group=None):
        self.config, self.group = config, group
        if device is None:
            device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device)

    def init_model(self, heap=None) -> None:
        rm = self.config.reward_model
        self.model_cfg = get_model_config(rm.model.path, **{**rm.model.override_config, "value_head": True})
        self.engine = Zero3Engine(self.model_cfg, self.device, group=self.group, comm="nccl", trainable=False,
                                  seed=self.config.data.seed + 2, heap=heap)
        maybe_load_pretrained(self.engine, rm.model.path, dist.get_rank(self.group) if dist.is_initialized() else 0)
        rt = ActorRuntimeConfig(ppo_micro_batch_size_per_gpu=rm.micro_batch_size_per_gpu, use_dynamic_bsz=rm.use_dynamic_bsz,
                                ppo_max_token_len_per_gpu=rm.forward_max_token_len_per_gpu)
        self.scorer = DataParallelPPOCritic(self.engine, rt, dp_group=self.group)
        # reward models with their own vocabulary / chat template (verl reward_model.model.input_tokenizer): the policy's tokens are
        # decoded and re-encoded with the reward model's tokenizer before scoring
        self.src_tok = self.rm_tok = None
        if rm.model.input_tokenizer:
            from ..data.rl_dataset import load_tokenizer

            self.src_tok = load_tokenizer(rm.model.input_tokenizer)
            self.rm_tok = load_tokenizer(rm.model.tokenizer_path or rm.model.path)

    def _switch_tokenizer(self, data: DataProto) -> DataProto:
        """(prompt, response) pairs re-tokenised for the reward model: left-padded prompt + right-padded response of the SAME
        widths as the input batch (longer re-encodings are truncated from the prompt's left / the response's right)."""
        ids, mask = data.batch["input_ids"], data.batch["attention_mask"].bool()
        r = data.batch["responses"].shape[1]
        p = ids.shape[1] - r
        pad = self.rm_tok.pad_token_id or 0
        out_ids = torch.full_like(ids, pad)
        out_mask = torch.zeros_like(data.batch["attention_mask"])
        for i in range(ids.shape[0]):
            prompt = self.src_tok.decode(ids[i, :p][mask[i, :p]].tolist())
            resp = self.src_tok.decode(ids[i, p:][mask[i, p:]].tolist())
            pi = list(self.rm_tok.encode(prompt, add_special_tokens=False))[-p:]
            ri = list(self.rm_tok.encode(resp, add_special_tokens=False))[:r]
            if pi:
                out_ids[i, p - len(pi):p] = torch.tensor(pi, dtype=ids.dtype, device=ids.device)
                out_mask[i, p - len(pi):p] = 1
            if ri:
                out_ids[i, p:p + len(ri)] = torch.tensor(ri, dtype=ids.dtype, device=ids.device)
                out_mask[i, p:p + len(ri)] = 1
        pos = (out_mask.cumsum(-1) - 1).clamp_min(0)
        new = DataProto.from_dict({"input_ids": out_ids, "attention_mask": out_mask, "position_ids": pos, "responses": out_ids[:, p:],
                                   "response_mask": out_mask[:, p:]})
        new.meta_info = dict(data.meta_info)
        return new

    @torch.no_grad()
    def compute_rm_score(self, data: DataProto) -> DataProto:
        if self.rm_tok is not None:
            scored = self.compute_rm_score_tokens(self._switch_tokenizer(data))
            # the scalar goes back onto the last valid token of the ORIGINAL response
            mask = data.batch["response_mask"] if "response_mask" in data.batch else data.batch["attention_mask"][:, -data.batch["responses"].shape[1]:]
            last = (mask.sum(-1).long() - 1).clamp_min(0)
            out = torch.zeros(mask.shape, dtype=torch.float32, device=mask.device)
            out[torch.arange(out.shape[0], device=out.device), last] = scored.batch["rm_scores"].sum(-1).to(out.device)
            return DataProto.from_dict({"rm_scores": out * (mask.sum(-1, keepdim=True) > 0)})
        return self.compute_rm_score_tokens(data)

    @torch.no_grad()
    def compute_rm_score_tokens(self, data: DataProto) -> DataProto:
        data.meta_info.setdefault("micro_batch_size", self.config.reward_model.micro_batch_size_per_gpu)
        values = self.scorer.compute_values(data)                     # value BEFORE each response token
        mask = data.batch["response_mask"] if "response_mask" in data.batch else data.batch["attention_mask"][:, -values.shape[1]:]
        last = (mask.sum(-1).long() - 1).clamp_min(0)
        score = values.gather(1, last[:, None]).squeeze(1)
        out = torch.zeros_like(values)
        out[torch.arange(out.shape[0], device=out.device), last] = score
        return DataProto.from_dict({"rm_scores": out * (mask.sum(-1, keepdim=True) > 0)})


class CriticWorker:
    def __init__(self, config: PPOConfig, device=None, group=None):
        self.config, self.group = config, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if device is None:
            device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device)
        self.lr_step = 0
        self.profiler = None

    def start_profile(self, **kw) -> None:
        if self.profiler is None:
            role_cfg = getattr(self.config.critic, "profiler", None)
            self.profiler = DistProfiler(self.rank, self.config.global_profiler, role_cfg if (role_cfg is not None and role_cfg.enable) else None,
                                         role="critic")
        self.profiler.start(**kw)

    def stop_profile(self) -> None:
        if self.profiler is not None:
            self.profiler.stop()

    def init_model(self, heap=None) -> None:
        cfg = self.config
        c = cfg.critic
        path = c.model.path if c.model.path != "tiny-llama" or cfg.actor_rollout_ref.model.path == "tiny-llama" \
            else cfg.actor_rollout_ref.model.path
        self.model_cfg = get_model_config(path, **{**c.model.override_config, "value_head": True})
        self.engine = Zero3Engine(self.model_cfg, self.device, group=self.group, comm=c.fsdp_config.comm_backend,
                                  trainable=True, seed=cfg.data.seed + 1, lr=c.optim.lr, betas=tuple(c.optim.betas),
                                  eps=c.optim.eps, weight_decay=c.optim.weight_decay, max_grad_norm=c.grad_clip, heap=heap)
        maybe_load_pretrained(self.engine, path, self.rank)
        n_dp = self.world
        mini = max(1, c.ppo_mini_batch_size * cfg.actor_rollout_ref.rollout.n // n_dp)
        rt = ActorRuntimeConfig(ppo_mini_batch_size=mini, ppo_micro_batch_size_per_gpu=c.ppo_micro_batch_size_per_gpu,
                                use_dynamic_bsz=c.use_dynamic_bsz, ppo_max_token_len_per_gpu=c.ppo_max_token_len_per_gpu,
                                loss_agg_mode=c.loss_agg_mode, grad_clip=c.grad_clip, cliprange_value=c.cliprange_value)
        self.critic = DataParallelPPOCritic(self.engine, rt, dp_group=self.group)

    @DistProfiler.annotate("compute_values")
    @torch.no_grad()
    def compute_values(self, data: DataProto) -> DataProto:
        c = self.config.critic
        data.meta_info.setdefault("micro_batch_size", c.ppo_micro_batch_size_per_gpu)
        data.meta_info.setdefault("max_token_len", c.forward_max_token_len_per_gpu)
        data.meta_info.setdefault("use_dynamic_bsz", c.use_dynamic_bsz)
        return DataProto.from_dict({"values": self.critic.compute_values(data)})

    @DistProfiler.annotate("update_critic_stream")
    def update_critic_stream(self, data: DataProto) -> DataProto:
        meta = data.meta_info
        is_opt = meta.get("is_opt_step", True)
        opt = self.config.critic.optim
        total = opt.total_training_steps if opt.total_training_steps > 0 else (self.config.trainer.total_training_steps or 0)
        lr = lr_at_step(opt, self.lr_step, total)
        metrics = self.critic.update_critic_stream(data, is_opt_step=is_opt, lr=lr)
        metrics["critic/lr"] = [lr]
        if meta.get("is_lr_step", is_opt):
            self.lr_step += 1
        return DataProto(meta_info={"metrics": metrics})

    update_critic = update_critic_stream

    def optimizer_step_only(self, is_lr_step: bool = True) -> dict:
        opt = self.config.critic.optim
        total = opt.total_training_steps if opt.total_training_steps > 0 else (self.config.trainer.total_training_steps or 0)
        lr = lr_at_step(opt, self.lr_step, total)
        st = self.engine.step(lr)
        if is_lr_step:
            self.lr_step += 1
        return {"critic/grad_norm": [st["grad_norm"]], "critic/lr": [lr]}

    def _ckpt_manager(self):
        if getattr(self, "_ckpt_mgr", None) is None:
            from ..utils.checkpoint import ShardedCheckpointManager

            ck = self.config.critic.checkpoint
            self._ckpt_mgr = ShardedCheckpointManager(self.engine, self.model_cfg, tuple(ck.save_contents), tuple(ck.load_contents))
        return self._ckpt_mgr

    def save_checkpoint(self, local_path, hdfs_path=None, global_step=0, max_ckpt_to_keep=None) -> None:
        self._ckpt_manager().save(local_path, global_step, max_ckpt_to_keep, extra={"lr_step": self.lr_step})

    def load_checkpoint(self, local_path, hdfs_path=None, del_local_after_load=False) -> None:
        if not local_path:
            return
        extra = self._ckpt_manager().load(local_path)
        self.lr_step = (extra or {}).get("lr_step", 0)
