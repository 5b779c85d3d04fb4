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
                    help="own = product (also times the library baseline afterwards, see --baseline-steps); reference = the unmodified "
                         "reference (unavailable offline); baseline = ONLY this framework with the named ops routed through stock "
                         "libraries (BASELINE.md section 3): cuBLAS GEMMs + unfused SwiGLU, materialised-logits chunked CE, cuBLAS "
                         "decode GEMMs, NCCL collectives.  Memory policy (resident replica, saved activations) is IDENTICAL to own")
    ap.add_argument("--baseline-steps", type=int, default=int(os.environ.get("BENCH_BASELINE_STEPS", 3)),
                    help="after the own arm, rebuild the job with the library switches and time this many steps in the same "
                         "invocation -> vs_baseline (0 = skip)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --prompts-per-gpu prompts on every GPU; strong: --total-prompts split over the GPUs")
    ap.add_argument("--total-prompts", type=int, default=int(os.environ.get("BENCH_TOTAL_PROMPTS", 128)))
    ap.add_argument("--rollout-tp", type=int, default=int(os.environ.get("BENCH_ROLLOUT_TP", 1)),
                    help="tensor-parallel size of the colocated rollout engine (TP groups of consecutive ranks)")
    ap.add_argument("--no-selfcheck", action="store_true")
    ap.add_argument("--model", default=os.environ.get("BENCH_MODEL", "llama3-8b"))
    ap.add_argument("--prompts-per-gpu", type=int, default=int(os.environ.get("BENCH_PROMPTS_PER_GPU", 16)))
    ap.add_argument("--n", type=int, default=int(os.environ.get("BENCH_N", 8)))
    ap.add_argument("--prompt-len-max", type=int, default=int(os.environ.get("BENCH_PROMPT_LEN", 256)))
    ap.add_argument("--response-length", type=int, default=int(os.environ.get("BENCH_RESPONSE_LEN", 512)))
    ap.add_argument("--max-token-len", type=int, default=int(os.environ.get("BENCH_MAX_TOKEN_LEN", 16384)))
    ap.add_argument("--mode", default="colocated", choices=["colocated", "disagg"],
                    help="colocated: every GPU rolls out and trains; disagg: N/2 rollout GPUs (TP engine servers behind the C++ "
                         "manager) stream to N/2 ZeRO trainers with a weight push every step (BASELINE config 3)")
    ap.add_argument("--comm", default=os.environ.get("BENCH_COMM", "p2p"), choices=["p2p", "nccl"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--rollout-weight-dtype", default="bf16", choices=["bf16", "fp8"],
                    help="fp8: rollout projections re-quantised to e4m3 every step (BASELINE config 4); training stays bf16")
    return ap.parse_args()


def reference_arm(args) -> int:
    """The reference (Terra-Flux/PolyRL) installs here only with --no-deps; its own entry point
    (python -m rlboost.verl_stream.trainer.main_stream) imports hydra / ray / omegaconf / tensordict / verl /
    sglang and spawns the Rust manager with `cargo run` — none of which exist on this image (no network)."""
    why = ("reference installs with --no-deps (baseline/_ref) but cannot run: its entry point needs ray, sglang==0.5.5, hydra-core, "
           "omegaconf, tensordict, rpyc and a Rust toolchain (cargo run for the manager); none are installed/installable offline — "
           "see baseline/install_attempt*.log, DESIGN.md section 5")
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


# library arm: backend switches + a larger transient reserve for the materialised [chunk, V] fp32 logits of the eager CE
LIB_ENV = {"POLYRL_GEMM_BACKEND": "cublas", "POLYRL_CE_BACKEND": "torch", "POLYRL_DECODE_GEMM": "cublas", "POLYRL_ACT_RESERVE_GB": "30"}


def build_job(args, world, device, library: bool):
    """Construct the colocated job.  ``library`` flips ONLY the backend switches (cuBLAS GEMMs / materialised-logits CE / cuBLAS
    decode GEMMs / NCCL collectives); shapes, memory policy and algorithm are the same as the product arm."""
    from polyrl_b200.config.schema import load_config
    from polyrl_b200.trainer.colocated import ColocatedPPO

    for k, v in LIB_ENV.items():
        if library:
            os.environ[k] = v
        else:
            os.environ.pop(k, None)
    P, n, R = args.P, args.n, args.response_length
    cfg = load_config(overrides=[
        f"actor_rollout_ref.model.path={args.model}", "algorithm.adv_estimator=grpo", f"actor_rollout_ref.rollout.n={n}",
        f"data.max_prompt_length={args.prompt_len_max}", f"data.max_response_length={R}",
        f"data.train_batch_size={P * world}", f"actor_rollout_ref.actor.ppo_mini_batch_size={P * world}",
        "actor_rollout_ref.actor.use_dynamic_bsz=True", f"actor_rollout_ref.actor.ppo_max_token_len_per_gpu={args.max_token_len}",
        "actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True",
        f"actor_rollout_ref.rollout.log_prob_max_token_len_per_gpu={args.max_token_len}",
        f"actor_rollout_ref.rollout.max_num_seqs={P * n}", f"actor_rollout_ref.rollout.cuda_graph={not args.no_graph}",
        f"actor_rollout_ref.actor.fsdp_config.comm_backend={'nccl' if library else args.comm}",
        "actor_rollout_ref.rollout.free_cache_engine=True",
        f"actor_rollout_ref.rollout.tensor_model_parallel_size={args.rollout_tp}",
        f"trainer.n_gpus_per_node={world}", "actor_rollout_ref.actor.optim.lr=1e-6",
        f"actor_rollout_ref.rollout.weight_dtype={args.rollout_weight_dtype}",
    ])
    return ColocatedPPO(cfg, device=device, group=None)


def selfcheck(ppo, world, device) -> dict:
    """Cross-rank correctness evidence for the driver (the multi-GPU unit tests skip on its 1-GPU test box): peer-memory
    reduce-scatter vs NCCL on a bucket of this engine's size class, bit-identical bf16 replicas on every rank after an optimizer
    step, and TP lock-step of the rollout engine when --rollout-tp > 1."""
    import torch
    import torch.distributed as dist

    out: dict = {"world": world}
    eng = ppo.worker.actor_engine
    if world > 1:
        # (a) replica checksum: every rank must hold the same bf16 parameters after the step's all-gather
        cs = torch.zeros(2, device=device, dtype=torch.float64)
        for u in eng.units[:8]:
            full = u._resident if getattr(u, "_resident", None) is not None else None
            if full is None:
                continue
            f = full.float()
            cs[0] += f.double().sum()
            cs[1] += f.double().abs().sum()
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["replica_checksum_equal_across_ranks"] = bool(torch.equal(lo, hi))
        out["replica_checksum"] = [float(lo[0]), float(lo[1])]
        # (b) peer-memory reduce-scatter vs NCCL reduce-scatter on the same data
        if eng.comm == "p2p" and eng.heap is not None and hasattr(eng, "selfcheck_reduce_scatter"):
            out.update(eng.selfcheck_reduce_scatter())
    tp = getattr(ppo.worker, "rollout_tp", 1)
    if tp > 1 and hasattr(ppo.worker.rollout, "lockstep_digest"):
        d = torch.tensor([ppo.worker.rollout.lockstep_digest()], device=device, dtype=torch.int64)
        lo, hi = d.clone(), d.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=ppo.worker.tp_group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=ppo.worker.tp_group)
        out["tp_lockstep"] = bool(torch.equal(lo, hi))
    return out


def run_arm(args, ppo, world, rank, local_rank, device, steps, warmup, with_clocks: bool):
    import torch
    import torch.distributed as dist

    from polyrl_b200 import ops
    from polyrl_b200.data.synthetic import SyntheticPromptDataset
    from polyrl_b200.protocol import DataProto, collate_fn

    P = args.P
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

    check = {}
    for w in range(warmup):
        ppo.step(host_batch(w))
        if w == 0 and with_clocks and not args.no_selfcheck:
            check = selfcheck(ppo, world, device)
    sync_all()
    sampler = ClockSampler(local_rank)
    if with_clocks:
        sampler.start()
    launches0 = ops.total_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    d2h_bytes, samples, tokens = 0, 0, 0
    t_wall0 = time.perf_counter()
    gen_intervals, last_timing = [], {}
    for k in range(steps):
        hb = host_batch(warmup + k)              # host-side batch assembly is part of the e2e region
        ev[k][0].record()
        t_wall = time.time()
        m = ppo.step(hb)                         # includes H2D of hb and D2H of losses/metrics (.item())
        ev[k][1].record()
        gen_intervals.append((t_wall, t_wall + m.get("timing_s/gen", 0.0)))
        last_timing = {kk: round(v, 3) for kk, v in m.items() if kk.startswith("timing_s/")}
        if rank == 0:
            print("[bench] step", k, last_timing,
                  "engine", {kk: round(float(v), 3) for kk, v in ppo.worker.rollout.stats.items()}, "cpus", len(os.sched_getaffinity(0)),
                  "saved_act_layers", ppo.worker.actor_engine.saved_layers_last, "resident", ppo.worker.actor_engine.resident,
                  "mem_GB", round(torch.cuda.max_memory_allocated() / 2**30, 1),
                  "launches", dict(sorted(ops.LAUNCHES.items())), file=sys.stderr, flush=True)
        d2h_bytes += 8 * sum(1 for v in m.values() if isinstance(v, float))
        samples += int(m["num_samples"])
        tokens += int(m.get("perf/total_num_tokens", 0))
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if with_clocks else {}
    if with_clocks:
        clocks.update(sampler.phase_medians(gen_intervals))
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms, t_wall * 1e3, float(samples), float(ops.total_launches() - launches0), float(tokens)], device=device,
                     dtype=torch.float64)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms = tmax[0].item(), tmax[1].item()
        total_samples, launches, total_tokens = tsum[2].item(), tsum[3].item(), tsum[4].item()
    else:
        dev_ms, wall_ms, total_samples, launches, total_tokens = (x.item() for x in t)
    return {"dev_ms": dev_ms, "wall_ms": wall_ms, "samples": total_samples, "launches": launches, "tokens": total_tokens,
            "clocks": clocks, "h2d": h2d_bytes, "d2h": int(d2h_bytes / max(1, steps)), "timing": last_timing, "selfcheck": check,
            "saved_layers": ppo.worker.actor_engine.saved_layers_last, "resident": bool(ppo.worker.actor_engine.resident)}


def main() -> int:
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    if args.mode == "disagg":
        from polyrl_b200.trainer.bench_disagg import main as disagg_main

        return disagg_main(args)

    # stdout carries exactly ONE JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    # ---- library baseline FIRST, as a child process per rank (own rendezvous port): its memory, CUDA graphs and symmetric heaps are
    # gone when it exits, so the product arm below runs on a clean device with the same shapes and the same memory policy
    base, base_err = None, None
    if args.impl == "own" and args.baseline_steps > 0:
        env = dict(os.environ)
        env["MASTER_PORT"] = str(int(os.environ.get("MASTER_PORT", "29500")) + 13)
        env["BENCH_CHILD"] = "1"
        env["TORCHELASTIC_USE_AGENT_STORE"] = "False"      # the children rendezvous on their own port: child rank 0 hosts the store
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "baseline", "--gpus", str(args.gpus), "--steps", str(args.baseline_steps),
               "--warmup", str(3 if args.warmup >= 3 else max(1, args.warmup)), "--baseline-steps", "0", "--no-selfcheck", "--model", args.model,
               "--prompts-per-gpu", str(args.prompts_per_gpu), "--n", str(args.n), "--prompt-len-max", str(args.prompt_len_max),
               "--response-length", str(args.response_length), "--max-token-len", str(args.max_token_len), "--scaling", args.scaling,
               "--total-prompts", str(args.total_prompts), "--rollout-tp", str(args.rollout_tp),
               "--rollout-weight-dtype", args.rollout_weight_dtype] + (["--no-graph"] if args.no_graph else [])
        try:
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=float(os.environ.get("BENCH_BASELINE_TIMEOUT_S", 900)))
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if rank == 0:
                sys.stderr.write(r.stderr[-6000:])
                if r.returncode == 0 and lines:
                    b = json.loads(lines[-1])
                    base = {"value": b["value"], "unit": b["unit"], "steps": b["steps"], "ms_per_step": b["ms_per_step"],
                            "what": "same framework, child process of this invocation, same shapes and memory policy; GEMMs -> cuBLAS (+ eager "
                                    "SwiGLU), linear-CE -> chunked materialised logits, decode GEMMs -> cuBLAS, collectives -> NCCL "
                                    "(BASELINE.md section 3)",
                            "timing_s": b.get("timing_s"), "activation_checkpoint": b["config"]["activation_checkpoint"]}
                else:
                    base_err = f"baseline child rc={r.returncode}: {r.stderr[-300:]}"
        except Exception as e:      # never let the comparison arm take the product number down with it
            base_err = f"{type(e).__name__}: {str(e)[:200]}"
        time.sleep(float(os.environ.get("BENCH_COOLDOWN_S", 10)))     # let the power/thermal state settle before the product arm

    import torch
    import torch.distributed as dist

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

    ops.ext()   # fail loudly if the sm_100a extension is missing
    if args.scaling == "strong":
        if args.total_prompts % world:
            raise SystemExit(f"--total-prompts {args.total_prompts} not divisible by {world} GPUs")
        args.P = args.total_prompts // world
    else:
        args.P = args.prompts_per_gpu
    n, R = args.n, args.response_length
    per_gpu = args.P * n

    own_only_library = args.impl == "baseline"
    ppo = build_job(args, world, device, library=own_only_library)
    res = run_arm(args, ppo, world, rank, local_rank, device, args.steps, args.warmup, with_clocks=True)
    value = res["samples"] / (res["dev_ms"] / 1e3)
    e2e = res["samples"] / (res["wall_ms"] / 1e3)
    parallelism = (f"colocated dp{world} rollout replicas" if args.rollout_tp == 1 else
                   f"colocated {world // args.rollout_tp} x TP{args.rollout_tp} rollout engines") + \
                  (f" + zero x{world} ({'nccl' if own_only_library else args.comm}; fp32 master/moments/grads sharded, bf16 replica "
                   f"{'resident' if res['resident'] or world == 1 else 'gathered per layer'})")
    num_layers = ppo.worker.model_cfg.num_layers

    if rank == 0:
        out = {
            "metric": "PPO-step samples/sec (rollout+update, device-timed max over ranks)", "value": round(value, 4),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(res["dev_ms"] / args.steps, 2), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": round(value / base["value"], 4) if base else None,
            "dtype": "bf16", "data": "synthetic GSM8K-shaped prompts (64-192 tok), random-init weights",
            "config": {"model": args.model, "algorithm": "GRPO", "global_batch": args.P * world, "samples_per_step": per_gpu * world,
                       "n": n, "seq_len": args.prompt_len_max + R, "response_length": R, "parallelism": parallelism,
                       "activation_checkpoint": f"auto: {res['saved_layers']}/{num_layers} blocks saved, rest recomputed",
                       "max_tokens_per_microbatch": args.max_token_len,
                       "l2": "working set >> L2: 16 GB bf16 weights + KV cache streamed every step; no cache reuse across iterations",
                       "optimizer": "AdamW fp32 master+moments, grad clip 1.0", "old_log_prob_recompute": True,
                       "cuda_graph_decode": not args.no_graph, "rollout_weight_dtype": args.rollout_weight_dtype,
                       "rollout_tp": args.rollout_tp},
            "clocks": res["clocks"],
            "e2e": {"value": round(e2e, 4), "unit": "samples/s", "h2d_bytes_per_step": res["h2d"],
                    "d2h_bytes_per_step": res["d2h"] + per_gpu * R * 12},
            "gpu_launches": int(res["launches"]), "impl": args.impl,
            "tokens_per_s": round(res["tokens"] / (res["dev_ms"] / 1e3), 1),
            "timing_s": res["timing"], "selfcheck": res["selfcheck"],
            "baseline": base if base else ({"unavailable": base_err} if base_err else None),
            "vs_baseline_note": "BASELINE.md publishes no comparable number and the reference cannot run offline; vs_baseline = value / "
                                "the library-baseline arm timed by this same invocation just before the product arm (see 'baseline')" if base else None,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
