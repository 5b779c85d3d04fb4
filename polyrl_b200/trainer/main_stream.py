"""Entry point of the streaming trainer:  ``python -m polyrl_b200.trainer.main_stream [--config-file F] key=value ...``

Launch it once per GPU (``torchrun --nproc-per-node N -m polyrl_b200.trainer.main_stream ...`` or the
``--spawn`` helper).  Ranks ``[0, N - n_rollout_gpus)`` are trainer ranks (ZeRO-3 group); with
``trainer.n_rollout_gpus > 0`` the remaining ranks host native rollout engines behind the SGLang-compatible
HTTP front-end and register with the C++ rollout manager, which rank 0 spawns.  With ``n_rollout_gpus = 0``
the run is colocated (every GPU rolls out, then trains).

Reference flow: rlboost/verl_stream/trainer/main_stream.py:40-391 (Hydra main -> TaskRunner: build workers,
datasets, reward fns, spawn manager via ``cargo run``, init_workers, register_weight_senders, validate
``min_stream_batch_size``, ``fit()``).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

from ..config.schema import PPOConfig, load_config, need_critic, need_reference_policy, validate_config


STEP_CALLBACK = None        # set by bench.py --mode disagg: callable(trainer, global_step, metrics) on every trainer rank


def _init_dist(device_type: str, timeout_s: int | None = None):
    """``timeout_s`` = actor_rollout_ref.nccl_timeout: how long a collective may wait for a straggling rank before the job aborts."""
    import datetime

    kw = {"timeout": datetime.timedelta(seconds=int(timeout_s))} if timeout_s else {}
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if os.environ["MASTER_ADDR"] in ("127.0.0.1", "localhost"):
            os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")       # single node: never resolve the (possibly unresolvable) hostname
        if device_type == "cuda":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
            from ..utils.affinity import set_numa_affinity

            set_numa_affinity(int(os.environ.get("LOCAL_RANK", 0)))          # opt-in: POLYRL_NUMA_AFFINITY=1
            dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))), **kw)
        else:
            dist.init_process_group("gloo", **kw)
    return int(os.environ.get("RANK", 0)), world


def build_datasets(cfg: PPOConfig, model_cfg, tokenizer=None):
    from ..data.rl_dataset import ResumableRandomSampler, RLHFDataset, load_tokenizer
    from ..data.synthetic import SyntheticPromptDataset

    d = cfg.data
    from ..utils.fs import stage_path

    d.train_files, d.val_files = stage_path(d.train_files), stage_path(d.val_files)        # remote (hdfs://) datasets: once per node
    if d.synthetic or not d.train_files:
        train = SyntheticPromptDataset(model_cfg.vocab_size, d.synthetic_num_samples, tuple(d.synthetic_prompt_len_range),
                                       d.max_prompt_length, model_cfg.pad_token_id, seed=d.seed)
        val = SyntheticPromptDataset(model_cfg.vocab_size, max(8, d.val_batch_size or 64), tuple(d.synthetic_prompt_len_range),
                                     d.max_prompt_length, model_cfg.pad_token_id, seed=d.seed + 7)
        tok = None
    else:
        tok = tokenizer or load_tokenizer(d.tokenizer or cfg.actor_rollout_ref.model.tokenizer_path or cfg.actor_rollout_ref.model.path)
        if cfg.actor_rollout_ref.model.custom_chat_template and hasattr(tok, "chat_template"):
            tok.chat_template = cfg.actor_rollout_ref.model.custom_chat_template          # verl model.custom_chat_template
        raw = d.return_raw_chat or cfg.actor_rollout_ref.rollout.multi_turn.enable       # agent loops re-template the chat
        kw = dict(model_cfg=model_cfg, image_key=d.image_key, video_key=d.video_key, min_pixels=d.image_min_pixels, max_pixels=d.image_max_pixels,
                  apply_chat_template_kwargs=dict(d.apply_chat_template_kwargs or {}))
        if (d.custom_cls or {}).get("name") and (d.custom_cls or {}).get("path"):   # data.custom_cls.{path,name}: user dataset class with RLHFDataset's constructor
            from ..data.dynamic import load_extern_type

            RLHFDataset = load_extern_type(d.custom_cls.get("path"), d.custom_cls["name"])    # noqa: N806
        if (d.datagen or {}).get("name"):
            from ..data.dynamic import DynamicGenDataset, load_extern_type

            gen = load_extern_type(d.datagen.get("path"), d.datagen["name"])(d.datagen)
            train = DynamicGenDataset(d.train_files, tok, d.prompt_key, d.max_prompt_length, d.filter_overlong_prompts, d.truncation, raw,
                                      generator=gen, **kw)
        else:
            train = RLHFDataset(d.train_files, tok, d.prompt_key, d.max_prompt_length, d.filter_overlong_prompts, d.truncation, raw, **kw)
        val = (RLHFDataset(d.val_files, tok, d.prompt_key, d.max_prompt_length, d.filter_overlong_prompts, d.truncation, raw, **kw)
               if d.val_files else None)
    # data.train_max_samples / val_max_samples: only the first N rows (quick experiments on a slice of a big file)
    for ds, cap in ((train, d.train_max_samples), (val, d.val_max_samples)):
        if ds is not None and cap and cap > 0 and hasattr(ds, "rows") and len(ds.rows) > cap:
            ds.rows = ds.rows[:cap]
    if (d.sampler or {}).get("class_name"):
        from ..data.dynamic import load_extern_type

        sampler = load_extern_type(d.sampler.get("class_path"), d.sampler["class_name"])(train, d)
    else:
        sampler = ResumableRandomSampler(len(train), d.seed, d.shuffle)
    return train, val, sampler, tok


def run_rollout_rank(cfg: PPOConfig, rank: int, trainer_ranks, rollout_ranks, device, manager_endpoint: str, tp_groups=None) -> None:
    """One rollout GPU.  With ``rollout.tensor_model_parallel_size = t > 1`` consecutive rollout ranks form one TP-t engine:
    its first rank owns the HTTP front-end and the manager registration, the others mirror its scheduler commands in
    lock-step (rollout/server.py); every rank holds a full bf16 replica that the trainers push into and slices it locally."""
    from ..models import get_model_config
    from ..rollout.engine import RolloutEngine
    from ..rollout.server import EngineServer, run_tp_follower
    from ..weight_transfer.sync import WeightSync

    arr = cfg.actor_rollout_ref
    model_cfg = get_model_config(arr.model.path, **arr.model.override_config)
    r = arr.rollout
    sync = WeightSync(model_cfg, trainer_ranks, rollout_ranks, device, backend=arr.actor.fsdp_config.comm_backend)
    max_len = r.max_model_len or (cfg.data.max_prompt_length + cfg.data.max_response_length)
    tp = max(1, r.tensor_model_parallel_size)
    tp_ctx, cmd_group, tp_rank, replica = None, None, 0, rollout_ranks.index(rank)
    if tp > 1:
        from ..rollout.tp import TPContext

        nccl_group, cmd_group = tp_groups
        tp_rank, replica = rollout_ranks.index(rank) % tp, rollout_ranks.index(rank) // tp
        heap = None
        if torch.device(device).type == "cuda":
            from ..parallel.symm import SymmetricHeap, make_heap

            heap = make_heap(group=nccl_group, device=torch.device(device))
        tp_ctx = TPContext(model_cfg, heap, tp_rank, tp, max_batch=r.max_num_seqs, group=nccl_group)
    eng = RolloutEngine(model_cfg, device, max_num_seqs=r.max_num_seqs, max_model_len=max_len,
                        gpu_memory_utilization=r.gpu_memory_utilization, use_cuda_graph=r.cuda_graph,
                        max_prefill_tokens=r.max_num_batched_tokens, seed=cfg.data.seed + 17 * replica, tp=tp_ctx)
    e, l, h = sync.engine_views()
    eng.set_weights(e, l, h)
    if model_cfg.vision is not None:
        eng.set_vision_weights(sync.vision_view())       # the tower lives in the same pushed replica

    def on_update(body: dict):
        sync.receive(int(body.get("weight_version", 0)))
        if tp_ctx is not None:
            eng.set_weights(*sync.engine_views())          # re-slice the compacted qkv copies from the fresh replica
        return True, "ok"

    if tp_rank > 0:
        run_tp_follower(eng, cmd_group, weight_update_fn=on_update)
        return
    srv = EngineServer(eng, port=int(os.environ.get("POLYRL_ENGINE_PORT_BASE", 30000)) + rank, weight_update_fn=on_update,
                       tp_cmd_group=cmd_group).start()
    srv.register_with_manager(manager_endpoint, local=False)
    print(f"[rollout rank {rank}] engine serving on {srv.endpoint}", flush=True)
    srv.serve_until_shutdown()


def run_trainer_rank(cfg: PPOConfig, rank: int, trainer_ranks, rollout_ranks, device, manager_endpoint: str | None, trainer_group):
    from ..data.reward_manager import load_reward_manager
    from ..rollout.remote_client import RemoteRolloutClient
    from ..utils.tracking import Tracking
    from ..weight_transfer.agent import TrainerWeightSender
    from ..weight_transfer.sync import WeightSync
    from ..workers.hybrid_worker import ActorRolloutRefWorker, CriticWorker
    from .stream_trainer import StreamPPOTrainer

    elastic = bool(cfg.trainer.elastic_rollout)
    disagg = len(rollout_ranks) > 0 or elastic
    hybrid = disagg and (cfg.trainer.local_rollout or not rollout_ranks)      # without in-job rollout GPUs the local engines must generate
    role = ("actor" if (disagg and not hybrid) else "actor_rollout") + ("_ref" if need_reference_policy(cfg) else "")
    worker = ActorRolloutRefWorker(cfg, role=role, device=device, group=trainer_group)
    worker.init_model()
    critic = None
    if need_critic(cfg):
        critic = CriticWorker(cfg, device=device, group=trainer_group)
        critic.init_model(heap=worker.actor_engine.heap)
    train, val, sampler, tok = build_datasets(cfg, worker.model_cfg)
    reward_fn = None if tok is None else load_reward_manager(cfg, tok)
    client = None
    if disagg:
        sync = (WeightSync(worker.model_cfg, trainer_ranks, rollout_ranks, device, backend=cfg.actor_rollout_ref.actor.fsdp_config.comm_backend)
                if rollout_ranks else None)
        # engines of this job's own rollout ranks take the peer-memory / broadcast push; any other registered instance (elastic
        # engines on other hosts) is served over TCP at its handshake address
        tp_ = max(1, cfg.actor_rollout_ref.rollout.tensor_model_parallel_size)
        base_port = int(os.environ.get("POLYRL_ENGINE_PORT_BASE", 30000))
        own = {f"127.0.0.1:{base_port + r}" for i, r in enumerate(rollout_ranks) if i % tp_ == 0}
        sender = TrainerWeightSender(manager_endpoint, sync, group=trainer_group, sync_endpoints=own)
        r = cfg.actor_rollout_ref.rollout
        if rank == 0:
            client = RemoteRolloutClient(manager_endpoint, r.n, cfg.data.max_response_length, worker.model_cfg.pad_token_id,
                                         worker.model_cfg.eos_token_id, stream_size=r.min_stream_batch_size)
            client.wait_healthy()
            sender.register_as_sender()
            # wait until every rollout rank has registered and passed the health gate
            t0 = time.time()
            while time.time() - t0 < 600:
                st = client.instances_status()
                n_inst = len(rollout_ranks) // max(1, cfg.actor_rollout_ref.rollout.tensor_model_parallel_size)
                if st.get("pending_instances", 0) + st.get("ready_instances", 0) >= n_inst:
                    break
                time.sleep(0.2)
            time.sleep(0.5)
        else:
            from .stream_trainer import RemoteRolloutPlaceholder

            client = RemoteRolloutPlaceholder()      # generation is remote for this rank too; rank 0 owns the connection
        if hybrid:
            worker.start_local_server(manager_endpoint, int(os.environ.get("POLYRL_ENGINE_PORT_BASE", 30000)) + 100 + rank)
        if dist.is_initialized():
            dist.barrier(group=trainer_group)

        alias_local = worker.update_weight_remote      # colocated path: the local engine aliases the trainer's fresh weights

        def update_weight_remote():
            sender.update(worker.actor_engine, rank)
            if hybrid:
                alias_local()
            return True

        worker.update_weight_remote = update_weight_remote      # disaggregated: push over NVLink / broadcast

        class _NoopVersion:
            """the sender bumps the version itself; the trainer's explicit bump becomes a no-op"""

        if rank == 0:
            client.update_weight_version = lambda: sender.version
    tracking = Tracking(cfg.trainer.project_name, cfg.trainer.experiment_name, cfg.trainer.logger) if rank == 0 else None
    trainer = StreamPPOTrainer(cfg, worker, critic, reward_fn, reward_fn, train, val, client, trainer_group, tracking, sampler)
    if cfg.reward_model.enable:
        from ..workers.hybrid_worker import RewardModelWorker

        trainer.rm = RewardModelWorker(cfg, device=device, group=trainer_group)
        trainer.rm.init_model(heap=worker.actor_engine.heap)
    trainer.step_callback = STEP_CALLBACK
    out = trainer.fit()
    return out


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(add_help=True)
    ap.add_argument("--config-file", action="append", default=[])
    ap.add_argument("--manager-endpoint", default=None, help="use an already running rollout manager")
    ap.add_argument("overrides", nargs="*")
    a = ap.parse_args(argv)
    cfg = load_config(a.config_file, a.overrides)
    device_type = cfg.trainer.device if (cfg.trainer.device != "cuda" or torch.cuda.is_available()) else "cpu"
    rank, world = _init_dist(device_type, cfg.actor_rollout_ref.nccl_timeout)
    if os.environ.get("POLYRL_DUMP_STACKS_AFTER_S"):
        # hang diagnosis: every rank writes the Python stacks of all its threads to a file after N seconds (and every N after that)
        import faulthandler

        _fh = open(os.path.join(os.environ.get("POLYRL_DUMP_STACKS_DIR", "/tmp"), f"polyrl_stacks_rank{rank}.txt"), "w")
        faulthandler.dump_traceback_later(float(os.environ["POLYRL_DUMP_STACKS_AFTER_S"]), repeat=True, file=_fh)
    n_roll = cfg.trainer.n_rollout_gpus
    n_train = world - n_roll
    validate_config(cfg, n_gpus=world)
    trainer_ranks = list(range(n_train))
    rollout_ranks = list(range(n_train, world))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if device_type == "cuda" else torch.device("cpu")
    trainer_group = dist.new_group(trainer_ranks) if (dist.is_initialized() and n_roll > 0) else None
    # TP groups of the rollout ranks: every process must take part in every new_group call, in the same order
    my_tp_groups = None
    tp = max(1, cfg.actor_rollout_ref.rollout.tensor_model_parallel_size)
    if dist.is_initialized() and n_roll > 0 and tp > 1:
        if n_roll % tp != 0:
            raise ValueError(f"trainer.n_rollout_gpus ({n_roll}) must be a multiple of rollout.tensor_model_parallel_size ({tp})")
        for c in range(n_roll // tp):
            chunk = rollout_ranks[c * tp:(c + 1) * tp]
            g_main = dist.new_group(chunk)
            # command channel of the TP engine: on GPUs the TP (NCCL) group itself carries the small scheduler messages — a 4-GPU
            # run that also created gloo sub-groups hung before the manager was even spawned; gloo only on the CPU plumbing path
            g_cmd = g_main if device_type == "cuda" else dist.new_group(chunk, backend="gloo")
            if rank in chunk:
                my_tp_groups = (g_main, g_cmd)
    manager_endpoint = a.manager_endpoint or cfg.actor_rollout_ref.rollout.rollout_manager.endpoint
    mgr_proc = None
    elastic = bool(cfg.trainer.elastic_rollout)
    if (n_roll > 0 or elastic) and manager_endpoint is None:
        from ..weight_transfer.launcher import spawn_rollout_manager

        port = cfg.actor_rollout_ref.rollout.rollout_manager.port
        if rank == 0:
            mgr_proc, port = spawn_rollout_manager(port=port, config_path=cfg.actor_rollout_ref.rollout.rollout_manager.config_path,
                                                   initial_local_gen_s=int(cfg.actor_rollout_ref.rollout.rollout_manager.max_local_gen_s))
        box = [port]
        if dist.is_initialized():
            dist.broadcast_object_list(box, src=0)
        manager_endpoint = f"127.0.0.1:{box[0]}"
    try:
        if rank in rollout_ranks:
            run_rollout_rank(cfg, rank, trainer_ranks, rollout_ranks, device, manager_endpoint, my_tp_groups)
        else:
            out = run_trainer_rank(cfg, rank, trainer_ranks, rollout_ranks, device, manager_endpoint if (n_roll > 0 or elastic) else None,
                                   trainer_group)
            if rank == 0:
                print("[main_stream] final metrics:", {k: v for k, v in sorted(out.items()) if k.startswith(("perf/", "critic/score", "actor/pg_loss", "training/"))}, flush=True)
                if n_roll > 0:        # tell the rollout ranks to exit
                    import http.client
                    import json as _json

                    h, p = manager_endpoint.rsplit(":", 1)
                    c = http.client.HTTPConnection(h, int(p), timeout=10)
                    c.request("GET", "/get_instances_status")
                    st = _json.loads(c.getresponse().read().decode())
                    eps = st.get("ready_instance_endpoints", []) + st.get("pending_instance_endpoints", [])
                    c.request("POST", "/shutdown_instances", body=_json.dumps({"endpoints": eps}), headers={"Content-Type": "application/json"})
                    c.getresponse().read()
    finally:
        if mgr_proc is not None:
            time.sleep(0.5)
            mgr_proc.terminate()
    if dist.is_initialized():
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass
    return 0


if __name__ == "__main__":
    sys.exit(main())
