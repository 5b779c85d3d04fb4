#!/usr/bin/env python
"""Flagship benchmark: PPO(GRPO)-step samples/sec (rollout + update), Llama-3-8B, colocated rollout+train.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W``; for N>1 launched under
``python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`` (one rank per GPU, NCCL).
Prints ONE JSON line from rank 0.  Weak scaling: every GPU keeps ``--prompts-per-gpu`` prompts x n samples.

A "step" is a complete PPO step through the public API ``ColocatedPPO.step``: H2D of the prompt batch from
pinned host memory -> rollout (prefill + CUDA-graph decode with our kernels) -> reward -> old-log-prob
recompute -> GRPO advantages -> actor update (ZeRO-3 fwd/bwd with fused kernels, AdamW) -> weight sync ->
D2H of the step metrics.  Random-init weights and synthetic GSM8K-shaped prompts (no network).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference", "baseline"],
                    help="own = product; reference = the unmodified reference (unavailable offline); baseline = this framework with the "
                         "named ops routed through stock libraries (BASELINE.md section 3): cuBLAS GEMMs + unfused SwiGLU, materialised-"
                         "logits chunked CE, NCCL all-gather per layer / fp32 reduce-scatter, full recompute")
    ap.add_argument("--model", default=os.environ.get("BENCH_MODEL", "llama3-8b"))
    ap.add_argument("--prompts-per-gpu", type=int, default=int(os.environ.get("BENCH_PROMPTS_PER_GPU", 16)))
    ap.add_argument("--n", type=int, default=int(os.environ.get("BENCH_N", 8)))
    ap.add_argument("--prompt-len-max", type=int, default=int(os.environ.get("BENCH_PROMPT_LEN", 256)))
    ap.add_argument("--response-length", type=int, default=int(os.environ.get("BENCH_RESPONSE_LEN", 512)))
    ap.add_argument("--max-token-len", type=int, default=int(os.environ.get("BENCH_MAX_TOKEN_LEN", 16384)))
    ap.add_argument("--comm", default=os.environ.get("BENCH_COMM", "p2p"), choices=["p2p", "nccl"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--rollout-weight-dtype", default="bf16", choices=["bf16", "fp8"],
                    help="fp8: rollout projections re-quantised to e4m3 every step (BASELINE config 4); training stays bf16")
    return ap.parse_args()


def reference_arm(args) -> int:
    """The reference (Terra-Flux/PolyRL) installs here only with --no-deps; its own entry point
    (python -m rlboost.verl_stream.trainer.main_stream) imports hydra / ray / omegaconf / tensordict / verl /
    sglang and spawns the Rust manager with `cargo run` — none of which exist on this image (no network)."""
    why = ("reference needs ray, sglang==0.5.5, hydra-core, omegaconf, tensordict, rpyc, verl (empty submodule) and a Rust "
           "toolchain (cargo); none are installed/installable offline — see baseline/install_attempt*.log, DESIGN.md")
    if int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps({"impl": "reference", "unavailable": why}), flush=True)
    return 0


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "clocks.mem,temperature.gpu")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc, self.thr = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        import time as _t

        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")] + [_t.time()])

    def phase_medians(self, intervals: list[tuple[float, float]]) -> dict:
        """Median SM / memory clock and max temperature inside vs outside the given wall-clock intervals (rollout phases)."""
        inside, outside, mem, temp = [], [], [], []
        for r in self.rows:
            try:
                smv, t = float(r[1]), float(r[-1])
                mem.append(float(r[9])); temp.append(float(r[10]))
            except Exception:
                continue
            (inside if any(a <= t <= b for a, b in intervals) else outside).append(smv)
        med = lambda v: sorted(v)[len(v) // 2] if v else None
        return {"sm_mhz_rollout": med(inside), "sm_mhz_train": med(outside), "mem_mhz": med(mem), "temp_c_max": max(temp) if temp else None}

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(s for s, p in zip(sm, pw) if p > 0.3 * max(pw)) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


def main() -> int:
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    if args.impl == "baseline":
        os.environ["POLYRL_GEMM_BACKEND"] = "cublas"
        os.environ["POLYRL_CE_BACKEND"] = "torch"
        os.environ["POLYRL_DECODE_GEMM"] = "cublas"
        args.comm = "nccl"

    # stdout carries exactly ONE JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "bench.py needs a CUDA device"}))
        return 1
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if world != args.gpus and rank == 0:
        print(f"[bench] WORLD_SIZE={world} but --gpus={args.gpus}; using WORLD_SIZE", file=sys.stderr)

    from polyrl_b200 import ops
    from polyrl_b200.config.schema import load_config
    from polyrl_b200.data.synthetic import SyntheticPromptDataset
    from polyrl_b200.protocol import DataProto, collate_fn
    from polyrl_b200.trainer.colocated import ColocatedPPO

    ops.ext()   # fail loudly if the sm_100a extension is missing
    P, n, R = args.prompts_per_gpu, args.n, args.response_length
    per_gpu = P * n
    cfg = load_config(overrides=[
        f"actor_rollout_ref.model.path={args.model}", "algorithm.adv_estimator=grpo", f"actor_rollout_ref.rollout.n={n}",
        f"data.max_prompt_length={args.prompt_len_max}", f"data.max_response_length={R}",
        f"data.train_batch_size={P * world}", f"actor_rollout_ref.actor.ppo_mini_batch_size={P * world}",
        "actor_rollout_ref.actor.use_dynamic_bsz=True", f"actor_rollout_ref.actor.ppo_max_token_len_per_gpu={args.max_token_len}",
        "actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True",
        f"actor_rollout_ref.rollout.log_prob_max_token_len_per_gpu={args.max_token_len}",
        f"actor_rollout_ref.rollout.max_num_seqs={per_gpu}", f"actor_rollout_ref.rollout.cuda_graph={not args.no_graph}",
        f"actor_rollout_ref.actor.fsdp_config.comm_backend={args.comm}", "actor_rollout_ref.rollout.free_cache_engine=True",
        f"trainer.n_gpus_per_node={world}", "actor_rollout_ref.actor.optim.lr=1e-6",
        f"actor_rollout_ref.rollout.weight_dtype={args.rollout_weight_dtype}",
        *(["actor_rollout_ref.actor.fsdp_config.resident_params=false", "actor_rollout_ref.actor.fsdp_config.activation_checkpoint=full"]
          if args.impl == "baseline" else []),
    ])
    ppo = ColocatedPPO(cfg, device=device, group=None)
    mcfg = ppo.worker.model_cfg
    ds = SyntheticPromptDataset(mcfg.vocab_size, 1 << 20, (64, 192), args.prompt_len_max, mcfg.pad_token_id, seed=1)

    def host_batch(step: int) -> DataProto:
        base = (step * world + rank) * P
        b = DataProto.from_single_dict(collate_fn([ds[base + i] for i in range(P)]))
        b.batch = b.batch.pin_memory()           # inputs start in pinned host memory every step
        b.meta_info["ignore_eos"] = False
        return b

    h2d_bytes = sum(v.numel() * v.element_size() for v in host_batch(0).batch.values())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        ppo.step(host_batch(w))
    sync_all()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ops.total_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    d2h_bytes = 0
    samples = 0
    t_wall0 = time.perf_counter()
    gen_intervals = []
    for k in range(args.steps):
        hb = host_batch(args.warmup + k)         # host-side batch assembly is part of the e2e region
        ev[k][0].record()
        t_wall = time.time()
        m = ppo.step(hb)                         # includes H2D of hb and D2H of losses/metrics (.item())
        ev[k][1].record()
        gen_intervals.append((t_wall, t_wall + m.get("timing_s/gen", 0.0)))
        if rank == 0:
            print("[bench] step", k, {kk: round(v, 3) for kk, v in m.items() if kk.startswith("timing_s/")},
                  "engine", {kk: round(float(v), 3) for kk, v in ppo.worker.rollout.stats.items()}, "cpus", len(os.sched_getaffinity(0)),
                  "saved_act_layers", ppo.worker.actor_engine.saved_layers_last, "resident", ppo.worker.actor_engine.resident,
                  "mem_GB", round(torch.cuda.max_memory_allocated() / 2**30, 1),
                  "launches", dict(sorted(ops.LAUNCHES.items())), file=sys.stderr, flush=True)
        d2h_bytes += 8 * sum(1 for v in m.values() if isinstance(v, float))
        samples += int(m["num_samples"])
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    clocks.update(sampler.phase_medians(gen_intervals))
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms, t_wall * 1e3, float(samples), float(ops.total_launches() - launches0)], device=device, dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, total_samples, launches = tmax[0].item(), tmax[1].item(), tsum[2].item(), tsum[3].item()
    else:
        dev_ms, wall_ms, total_samples, launches = t[0].item(), t[1].item(), t[2].item(), t[3].item()
    value = total_samples / (dev_ms / 1e3)
    e2e = total_samples / (wall_ms / 1e3)
    if rank == 0:
        out = {
            "metric": "PPO-step samples/sec (rollout+update, device-timed max over ranks)", "value": round(value, 4),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dev_ms / args.steps, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic GSM8K-shaped prompts (64-192 tok), random-init weights",
            "config": {"model": args.model, "algorithm": "GRPO", "global_batch": P * world, "samples_per_step": per_gpu * world,
                       "n": n, "seq_len": args.prompt_len_max + R, "response_length": R,
                       "parallelism": f"colocated dp{world} rollout replicas + zero x{world} ({args.comm}; fp32 master/moments/grads "
                                      f"sharded, bf16 replica {'resident' if ppo.worker.actor_engine.resident or world == 1 else 'gathered per layer'})",
                       "activation_checkpoint": f"auto: {ppo.worker.actor_engine.saved_layers_last}/{ppo.worker.model_cfg.num_layers} blocks "
                                                "saved, rest recomputed",
                       "l2": "working set >> L2: 16 GB bf16 weights + KV cache streamed every step; no cache reuse across iterations",
                       "optimizer": "AdamW fp32 master+moments, grad clip 1.0", "old_log_prob_recompute": True,
                       "cuda_graph_decode": not args.no_graph, "rollout_weight_dtype": args.rollout_weight_dtype},
            "clocks": clocks,
            "e2e": {"value": round(e2e, 4), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": int(d2h_bytes / max(1, args.steps)) + per_gpu * R * 12},
            "gpu_launches": int(launches), "impl": args.impl,
            "tokens_per_s": round(value * (args.prompt_len_max / 2 + R), 1),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
