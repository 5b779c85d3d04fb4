"""Multi-GPU checks, run under torchrun (one rank per GPU) by tests/test_multigpu.py or directly:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_worker.py
Every check compares a hand-written peer-memory / fused path against NCCL or a single-GPU reference."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from polyrl_b200.ops import ext
    from polyrl_b200.parallel.symm import SymmetricHeap

    C = ext()
    heap = SymmetricHeap(device=dev)
    results = {}

    # ---- 1. barrier + all-gather pull + reduce-scatter-accumulate vs NCCL
    n = 1 << 22
    shard = heap.alloc(n, torch.bfloat16)
    torch.manual_seed(100 + rank)
    shard.local.copy_(torch.randn(n, device=dev).to(torch.bfloat16))
    heap.barrier()
    out = torch.empty(world * n, device=dev, dtype=torch.bfloat16)
    C.p2p_allgather_pull(shard.ptrs, out.data_ptr(), n * 2, rank, 32)
    want = torch.empty_like(out)
    dist.all_gather_into_tensor(want, shard.local)
    torch.cuda.synchronize()
    results["allgather_pull_exact"] = bool(torch.equal(out, want))

    full = heap.alloc(world * n, torch.bfloat16)
    full.local.copy_(torch.randn(world * n, device=dev).to(torch.bfloat16))
    heap.barrier()
    acc = torch.ones(n, device=dev)
    ss = torch.zeros(1, device=dev)
    C.p2p_reduce_scatter_accum(full.ptrs, rank * n, acc, 0.5, rank, ss, 32)
    ref_in = full.local.float().clone()
    ref_out = torch.empty(n, device=dev)
    dist.reduce_scatter_tensor(ref_out, ref_in)
    want_acc = 1.0 + 0.5 * ref_out
    torch.cuda.synchronize()
    results["reduce_scatter_accum_err"] = rel(acc, want_acc)
    results["reduce_scatter_sumsq_err"] = abs(ss.item() - want_acc.pow(2).sum().item()) / want_acc.pow(2).sum().item()
    heap.barrier()

    # ---- 2. bandwidth of the peer kernels vs NCCL (CUDA events, max over ranks)
    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    nb = 256 << 20   # 256 MiB shard
    big = heap.alloc(nb // 2, torch.bfloat16)
    gout = torch.empty(world * nb // 2, device=dev, dtype=torch.bfloat16)
    heap.barrier()
    t_pull = timeit(lambda: C.p2p_allgather_pull(big.ptrs, gout.data_ptr(), nb, rank, 0))
    t_nccl = timeit(lambda: dist.all_gather_into_tensor(gout, big.local))
    results["allgather_256MiB_ms"] = {"p2p_pull": round(t_pull, 3), "nccl": round(t_nccl, 3),
                                      "p2p_GBps_per_gpu_in": round((world - 1) * nb / t_pull / 1e6, 1)}
    peer = (rank + 1) % world
    t_copy = timeit(lambda: C.p2p_copy(big.local.data_ptr(), big.ptrs[peer], nb, 0))
    results["peer_write_GBps"] = round(nb / t_copy / 1e6, 1)

    # reduce-scatter (+ fp32 accumulate, scale) of a layer-sized bf16 gradient buffer vs NCCL reduce_scatter + accumulate kernel
    rs_n = (218_112_000 // (8 * world)) * 8                       # one Llama-3-8B block / world, bf16 elements per rank chunk
    gbuf = heap.alloc(world * rs_n, torch.bfloat16)
    gbuf.local.normal_()
    acc_big = torch.zeros(rs_n, device=dev)
    heap.barrier()

    def rs_p2p():
        heap.barrier()
        C.p2p_reduce_scatter_accum(gbuf.ptrs, rank * rs_n, acc_big, 0.5, rank, None, 32)
        heap.barrier()

    rs_in = torch.empty(world * rs_n, device=dev)
    rs_out = torch.empty(rs_n, device=dev)

    def rs_nccl():                                                  # the reference path: cast to fp32, NCCL reduce-scatter, accumulate
        rs_in.copy_(gbuf.local)
        dist.reduce_scatter_tensor(rs_out, rs_in)
        C.grad_accum_(acc_big, rs_out, 0.5, None)

    t_rs, t_rs_nccl = timeit(rs_p2p, 5), timeit(rs_nccl, 5)
    results["reduce_scatter_layer_ms"] = {"p2p_fused_accum": round(t_rs, 3), "nccl_fp32_rs_plus_accum": round(t_rs_nccl, 3),
                                          "p2p_GBps_per_gpu_in": round((world - 1) * rs_n * 2 / t_rs / 1e6, 1)}

    # ---- 3. all-gather fused GEMM vs (NCCL all-gather + matmul)
    M, K, R = 4096, 4096, 1536 // (world // 2 if world > 2 else 1)
    R = max(256, (R // 256) * 256)
    wsh = heap.alloc(R * K, torch.bfloat16)
    torch.manual_seed(7 + rank)
    wsh.local.copy_((torch.randn(R * K, device=dev) * 0.05).to(torch.bfloat16))
    gathered = torch.empty(world * R * K, device=dev, dtype=torch.bfloat16)
    flags = torch.zeros(world * 8, device=dev, dtype=torch.int32)
    torch.manual_seed(3)
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    heap.barrier()
    epoch = [0]

    def fused():
        epoch[0] += 1
        return C.ag_gemm_bf16(a, wsh.ptrs, gathered.data_ptr(), flags.data_ptr(), epoch[0], R, rank, 3)

    d = fused()
    wfull = torch.empty(world * R * K, device=dev, dtype=torch.bfloat16)
    dist.all_gather_into_tensor(wfull, wsh.local)
    want = a.float() @ wfull.view(world * R, K).float().t()
    torch.cuda.synchronize()
    results["ag_gemm_err"] = rel(d, want)
    results["ag_gemm_gather_exact"] = bool(torch.equal(gathered, wfull))

    def baseline():
        dist.all_gather_into_tensor(wfull, wsh.local)
        return a @ wfull.view(world * R, K).t()

    results["ag_gemm_ms"] = {"fused": round(timeit(fused), 3), "nccl_ag_then_cublas": round(timeit(baseline), 3)}
    heap.barrier()

    # ---- 4. ZeRO-3 engine: p2p and nccl comm give the same update as a single-GPU engine
    from polyrl_b200.models import get_model_config, pack_padded
    from polyrl_b200.parallel.zero3 import Zero3Engine

    cfg = get_model_config("tiny-llama-128")
    torch.manual_seed(11)
    ids = torch.randint(3, cfg.vocab_size, (4, 48), device=dev)
    mask = torch.ones_like(ids)
    inp, _ = pack_padded(ids, mask)
    labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), device=dev)

    def run(engine):
        for _ in range(2):
            lp, ent = engine.forward_logprob_entropy(inp, labels, 1.0)
            (-(lp.mean()) + 0.01 * ent.mean()).backward()
        st = engine.step()
        torch.cuda.synchronize()
        return st, {n_: t.clone() for n_, t in engine.named_full_tensors()}

    single_group = dist.new_group([rank]) if world > 1 else None
    groups = [dist.new_group([r]) for r in range(world)]
    ref_eng = Zero3Engine(cfg, dev, group=groups[rank], comm="nccl", lr=1e-2, weight_decay=0.0, seed=5)
    st_ref, w_ref = run(ref_eng)
    for comm in ("p2p", "nccl"):
        eng = Zero3Engine(cfg, dev, comm=comm, lr=1e-2, weight_decay=0.0, seed=5)
        eng.grad_sink_enabled = True      # opt-in (POLYRL_GRAD_SINK=1): wgrads written straight into the reduce-scatter slot
        st, w = run(eng)
        err = max(rel(w[k], w_ref[k]) for k in w_ref)
        results[f"zero3_{comm}_weight_err"] = err
        results[f"zero3_{comm}_gradnorm"] = (round(st["grad_norm"], 4), round(st_ref["grad_norm"], 4))
        if comm == "p2p":                 # wgrad GEMMs store straight into the reduce-scatter slot (gradient sink)
            results["zero3_p2p_grad_staging"] = dict(eng.grad_stage_stats)
        dist.barrier()

    # ---- 5. trainer->rollout weight push over peer memory (all ranks train, all ranks host a replica buffer)
    from polyrl_b200.weight_transfer.sync import WeightSync

    sync = WeightSync(cfg, list(range(world)), list(range(world)), dev, backend="p2p")
    eng = Zero3Engine(cfg, dev, comm="p2p", seed=9, trainable=False)
    sync.push(eng, 1)
    sync.heap.barrier()
    torch.cuda.synchronize()
    e_w, l_w, h_w = sync.engine_views()
    full_ref = {n_: t.clone() for n_, t in eng.named_full_tensors()}
    ok = torch.equal(e_w["embed"], full_ref["embed"]) and all(
        torch.equal(l_w[i][k], full_ref[f"layers.{i}.{k}"]) for i in range(cfg.num_layers) for k in l_w[i])
    results["weight_push_exact"] = bool(ok)

    if rank == 0:
        import json

        print("MULTIGPU_RESULTS " + json.dumps(results), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
