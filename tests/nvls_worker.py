"""torchrun worker: NVLS heap (VMM + POSIX-fd sharing + multicast) and the multimem collectives against NCCL.
Prints ``NVLS_RESULTS {json}`` on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, iters=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=dev)
    from polyrl_b200.ops import ext
    from polyrl_b200.parallel.nvls import NvlsAllReduce, NvlsHeap

    C = ext()
    res = {"world": world, "multicast_supported": bool(C.nvls_supported(dev.index))}
    heap = NvlsHeap(device=dev)
    res["multicast"] = heap.multicast
    # ---- unicast peer mappings (VMM replaces legacy IPC)
    n = 1 << 20
    buf = heap.alloc(n, torch.bfloat16)
    buf.local.copy_(torch.full((n,), float(rank + 1), device=dev, dtype=torch.bfloat16))
    torch.cuda.synchronize()
    dist.barrier()
    # the peer mapping is usable by OUR kernels from this device: pull-gather through the mapped pointers
    gathered = torch.empty(world * n, device=dev, dtype=torch.bfloat16)
    C.p2p_allgather_pull(buf.peer_ptrs(), gathered.data_ptr(), n * 2, rank, 0)
    torch.cuda.synchronize()
    res["vmm_peer_kernel_pull_exact"] = all(bool((gathered[r * n:(r + 1) * n] == float(r + 1)).all().item()) for r in range(world))
    res["mc_mapped"] = bool(buf.mc_ptr)
    res["multicast_error"] = heap.multicast_error
    if heap.multicast and buf.mc_ptr:
        # ---- one-shot in-switch all-reduce vs NCCL
        B, H = 128, 4096
        ar = NvlsAllReduce(heap, 256 * 8192)
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        x = torch.randn(B * H, device=dev, generator=g).to(torch.bfloat16)
        r_ = torch.randn(B * H, device=dev, generator=torch.Generator(device=dev).manual_seed(7)).to(torch.bfloat16)
        ar.inp.local[: B * H].copy_(x)
        got = ar(B * H).clone()
        want = x.float().clone()
        dist.all_reduce(want)
        torch.cuda.synchronize()
        res["allreduce_max_abs_err"] = float((got.float() - want).abs().max().item())
        res["allreduce_ref_scale"] = float(want.abs().max().item())
        got2 = ar(B * H, residual=r_).clone()
        res["allreduce_residual_max_abs_err"] = float((got2.float() - (want + r_.float())).abs().max().item())
        # identical on every rank (lock-step requirement of TP engines)
        chk = got.float().sum().double().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res["allreduce_identical_across_ranks"] = bool(torch.equal(lo, hi))
        # graph replay keeps epochs in step
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            out_view = ar(B * H)
        for _ in range(3):
            gph.replay()
        torch.cuda.synchronize()
        res["allreduce_graph_replay_err"] = float((out_view.float() - want).abs().max().item())
        # ---- timing: NVLS vs NCCL vs the round-1 peer-pointer one-shot kernel
        for (b, h) in ((128, 4096), (64, 4096), (256, 4096), (512, 4096)):
            ne = b * h
            xb = torch.randn(ne, device=dev).to(torch.bfloat16)
            t_nvls = timeit(lambda: ar(ne))
            t_nccl = timeit(lambda: dist.all_reduce(xb))
            res[f"allreduce_us_B{b}"] = {"nvls": round(t_nvls * 1e3, 2), "nccl": round(t_nccl * 1e3, 2)}
        # in-graph (how the TP engine runs it): 16 all-reduces per replay
        gph2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph2):
            for _ in range(16):
                ar(128 * 4096)
        res["allreduce_us_B128_in_graph"] = round(timeit(lambda: gph2.replay(), iters=20, warm=5) / 16 * 1e3, 2)
        # ---- reduce-scatter + fp32 accumulate in the switch vs NCCL
        ng = 1 << 24
        gb = heap.alloc(ng, torch.bfloat16)
        gsrc = torch.randn(ng, device=dev, generator=g).to(torch.bfloat16)
        gb.local.copy_(gsrc)
        torch.cuda.synchronize()
        dist.barrier()
        chunk = ng // world
        acc = torch.zeros(chunk, device=dev, dtype=torch.float32)
        ssq = torch.zeros(1, device=dev, dtype=torch.float32)
        C.nvls_reduce_scatter_accum(gb.mc_ptr, rank * chunk, acc, 1.0 / world, ssq, 0)
        want_rs = torch.empty(chunk, device=dev, dtype=torch.float32)
        dist.reduce_scatter_tensor(want_rs, gsrc.float())
        want_rs /= world
        torch.cuda.synchronize()
        res["rs_accum_max_abs_err"] = float((acc - want_rs).abs().max().item())
        res["rs_sumsq_rel_err"] = float(((ssq - want_rs.pow(2).sum()).abs() / want_rs.pow(2).sum()).item())
        t_rs = timeit(lambda: C.nvls_reduce_scatter_accum(gb.mc_ptr, rank * chunk, acc, 1.0 / world, None, 0), iters=20, warm=3)
        out32 = torch.empty(chunk, device=dev, dtype=torch.float32)
        g32 = gsrc.float()
        t_rs_nccl = timeit(lambda: dist.reduce_scatter_tensor(out32, g32), iters=20, warm=3)
        res["rs_accum_ms"] = {"nvls": round(t_rs, 4), "nccl_fp32": round(t_rs_nccl, 4),
                              "nvls_GBps_in": round((world - 1) * chunk * 2 / t_rs / 1e6, 1)}
        # ---- multicast broadcast store (all-gather of bf16 shards)
        dstb = heap.alloc(ng, torch.bfloat16)
        shard = torch.full((chunk,), float(rank + 1), device=dev, dtype=torch.bfloat16)
        dist.barrier()
        C.nvls_bcast_store(shard, dstb.mc_ptr + rank * chunk * 2, 0)
        torch.cuda.synchronize()
        dist.barrier()
        ok = all(bool((dstb.local[r * chunk:(r + 1) * chunk] == float(r + 1)).all().item()) for r in range(world))
        res["bcast_allgather_exact"] = ok
        t_bc = timeit(lambda: C.nvls_bcast_store(shard, dstb.mc_ptr + rank * chunk * 2, 0), iters=20, warm=3)
        res["bcast_allgather_ms"] = round(t_bc, 4)
        res["bcast_allgather_GBps_out"] = round(chunk * 2 / t_bc / 1e6, 1)
    dist.barrier()
    heap.close()
    if rank == 0:
        print("NVLS_RESULTS " + json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
