"""``bench.py --mode disagg``: the flagship PolyRL deployment (BASELINE config 3) as a benchmark.

N GPUs of one node, one process per GPU under torchrun: ranks [0, N/2) are ZeRO trainers, ranks [N/2, N) host native rollout
engines (one TP engine of size min(4, N/2) per group) behind the C++ rollout manager that trainer rank 0 spawns.  Every step:
rank 0 submits the prompt batch to the manager, engines generate and STREAM finished GRPO groups back, the trainers run
old-log-prob / advantages / update on each streamed micro-batch while the engines are still decoding (bubble filling), then all
trainer ranks push their updated shards straight into the rollout GPUs' weight buffers over NVLink (``weight_transfer``) and the
manager bumps the weight version.  Same metric as the colocated mode: samples/s over the whole job.

Reference: examples/scripts/run_async_grpo_pipeline.sh:16-59 and rlboost/verl_stream/trainer/ppo/stream_ray_trainer.py:282-689.
"""
from __future__ import annotations

import json
import os
import sys
import time


def main(args) -> int:
    import torch
    import torch.distributed as dist

    from . import main_stream as ms

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    if world < 2 or world % 2:
        if rank == 0:
            print(json.dumps({"mode": "disagg", "unavailable": f"disaggregated mode needs an even number of GPUs >= 2, got {world}"}))
        return 0
    n_roll = world // 2
    n_train = world - n_roll
    tp = args.rollout_tp if args.rollout_tp > 1 else min(4, n_roll)
    while n_roll % tp:
        tp //= 2
    P = (args.total_prompts if args.scaling == "strong" else args.prompts_per_gpu * world)
    n, R = args.n, args.response_length
    total_steps = args.warmup + args.steps
    stream = max(n * n_train, (min(16, P) * n) // (n * n_train) * (n * n_train))       # prompts*n per streamed batch, whole groups
    algo = os.environ.get("BENCH_DISAGG_ALGO", "grpo")
    overrides = [
        f"actor_rollout_ref.model.path={args.model}", f"algorithm.adv_estimator={'gae' if algo == 'ppo' else 'grpo'}",
        f"actor_rollout_ref.rollout.n={n}", "data.synthetic=True", f"data.synthetic_num_samples={P * (total_steps + 2)}",
        f"data.max_prompt_length={args.prompt_len_max}", f"data.max_response_length={R}", f"data.train_batch_size={P}",
        f"actor_rollout_ref.actor.ppo_mini_batch_size={P}", "actor_rollout_ref.actor.use_dynamic_bsz=True",
        f"actor_rollout_ref.actor.ppo_max_token_len_per_gpu={args.max_token_len}",
        "actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True",
        f"actor_rollout_ref.rollout.log_prob_max_token_len_per_gpu={args.max_token_len}",
        "actor_rollout_ref.rollout.name=native-disaggregated", f"actor_rollout_ref.rollout.min_stream_batch_size={stream}",
        f"actor_rollout_ref.rollout.max_num_seqs={min(256, P * n // max(1, n_roll // tp))}",
        f"actor_rollout_ref.rollout.tensor_model_parallel_size={tp}",
        "actor_rollout_ref.rollout.rollout_manager.port=0", "actor_rollout_ref.rollout.rollout_manager.max_local_gen_s=600",
        f"actor_rollout_ref.actor.fsdp_config.comm_backend={args.comm}", "actor_rollout_ref.actor.optim.lr=1e-6",
        f"trainer.n_gpus_per_node={world}", f"trainer.n_rollout_gpus={n_roll}", "trainer.logger=[console]",
        f"trainer.total_training_steps={total_steps}", "trainer.total_epochs=1000", "trainer.save_freq=-1", "trainer.test_freq=-1",
        "trainer.val_before_train=False",
        # the reference's flagship script is the ASYNC pipeline (examples/scripts/run_async_grpo_pipeline.sh): the rollout GPUs
        # generate batch k+1 (with the weights of step k-1 -> one step off-policy) while the trainers update on batch k
        f"trainer.one_step_off={os.environ.get('BENCH_DISAGG_SYNC', '0') != '1'}",
    ]
    if algo == "ppo":
        overrides += [f"critic.model.path={args.model}", "critic.optim.lr=1e-6", "critic.use_dynamic_bsz=True",
                      f"critic.ppo_max_token_len_per_gpu={args.max_token_len}", f"critic.ppo_mini_batch_size={P}"]
    state = {"events": [], "timing": {}, "t0": None, "result": None}

    def cb(trainer, step, out):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        state["events"].append((step, ev, time.perf_counter()))
        state["timing"] = {k: round(float(v), 3) for k, v in out.items() if k.startswith("timing_s/")}
        if step == total_steps:
            torch.cuda.synchronize()
            e0 = next(e for s, e, _ in state["events"] if s == args.warmup)
            w0 = next(w for s, _, w in state["events"] if s == args.warmup)
            dev_ms = e0.elapsed_time(ev)
            wall_ms = (time.perf_counter() - w0) * 1e3
            t = torch.tensor([dev_ms, wall_ms], device="cuda", dtype=torch.float64)
            if trainer.group is not None or dist.is_initialized():
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=trainer.group)
            from .. import ops

            state["result"] = {"dev_ms": float(t[0]), "wall_ms": float(t[1]), "launches": ops.total_launches(),
                               "tokens": float(out.get("perf/total_num_tokens", 0)), "gen_s": out.get("timing_s/gen")}

    ms.STEP_CALLBACK = cb
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    # keep the trainer's own console output off stdout (the driver wants exactly ONE JSON line there)
    real_stdout = sys.stdout
    sys.stdout = sys.stderr
    try:
        rc = ms.main(overrides)
    finally:
        sys.stdout = real_stdout
    if rank == 0 and state["result"] is not None:
        r = state["result"]
        samples = args.steps * P * n
        value = samples / (r["dev_ms"] / 1e3)
        out = {
            "metric": "PPO-step samples/sec (rollout+update, device-timed max over ranks)", "value": round(value, 4), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["dev_ms"] / args.steps, 2),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic GSM8K-shaped prompts (64-192 tok), random-init weights", "mode": "disagg", "impl": "own",
            "config": {"model": args.model, "algorithm": "PPO+critic" if algo == "ppo" else "GRPO", "global_batch": P,
                       "samples_per_step": P * n, "n": n, "seq_len": args.prompt_len_max + R, "response_length": R,
                       "parallelism": f"disaggregated: {n_roll // tp} x TP{tp} rollout engine(s) on {n_roll} GPUs behind the C++ manager + "
                                      f"zero x{n_train} trainers ({args.comm}); streamed micro-batches of {stream} samples; weight push "
                                      "over NVLink every step",
                       "min_stream_batch_size": stream, "old_log_prob_recompute": True,
                       "one_step_off_policy": os.environ.get("BENCH_DISAGG_SYNC", "0") != "1"},
            "e2e": {"value": round(samples / (r["wall_ms"] / 1e3), 4), "unit": "samples/s",
                    "note": "prompts travel host -> manager (HTTP) -> engines, responses stream back through the manager to trainer rank 0"},
            "timing_s": state["timing"], "gpu_launches": int(r["launches"]),
        }
        print(json.dumps(out), flush=True)
    return rc
