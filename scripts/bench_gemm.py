"""Micro-benchmark: tcgen05 GEMM (own kernel) vs cuBLAS (torch.matmul) at training / decode shapes.
CUDA-event timing, warm-up, L2 flushed between iterations by rotating over buffers larger than L2."""
import json
import sys

import torch

from polyrl_b200.ops import ext


def bench(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    for s, e in evs:
        flush.zero_()
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2]


def main():
    C = ext()
    shapes = [
        (16384, 6144, 4096, 0), (16384, 4096, 4096, 0), (16384, 28672, 4096, 0), (16384, 4096, 14336, 0),
        (8192, 8192, 8192, 0), (256, 6144, 4096, 0), (256, 4096, 4096, 0), (256, 28672, 4096, 0), (256, 4096, 14336, 0),
        (256, 128256, 4096, 0),
    ]
    out = []
    for M, N, K, bn in shapes:
        a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
        d = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        t_own = bench(lambda: C.gemm_bf16(a, b, None, d, bn))
        t_lib = bench(lambda: torch.matmul(a, b.t(), out=d))
        fl = 2.0 * M * N * K
        rec = {"M": M, "N": N, "K": K, "own_ms": round(t_own, 4), "cublas_ms": round(t_lib, 4),
               "own_tflops": round(fl / t_own / 1e9, 1), "cublas_tflops": round(fl / t_lib / 1e9, 1)}
        if M <= 512:
            for fbn in (64, 128, 256):
                rec[f"own_bn{fbn}_ms"] = round(bench(lambda: C.gemm_bf16(a, b, None, d, fbn)), 4)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    # decode-shape GEMMs measured INSIDE a CUDA graph (as the engine runs them: no host launch overhead), cycling over
    # enough distinct weight copies that every call streams its weights from HBM (sum of copies > L2).
    ws = torch.empty(24 * 1024 * 1024, device="cuda", dtype=torch.float32)
    ctr = torch.zeros(4096, device="cuda", dtype=torch.int32)

    def graph_time(fns, reps=5):
        for f in fns:
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for f in fns:
                f()
        g.replay(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record(); g.replay(); e_.record(); torch.cuda.synchronize()
            ts.append(s_.elapsed_time(e_) / len(fns))
        return sorted(ts)[len(ts) // 2]

    for B in (64, 128, 256):
        tot = {"skinny": 0.0, "cublas": 0.0, "tile128": 0.0}
        for N, K in ((6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336), (128256, 4096)):
            copies = max(2, int(300e6 // (N * K * 2)) + 1)
            x = torch.randn(B, K, device="cuda", dtype=torch.bfloat16)
            Ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(copies)]
            mode = 1 if N == 28672 else 0
            t_s = graph_time([(lambda w=w: C.skinny_gemm(x, w, None, None, None, None, ws, ctr, mode, 0)) for w in Ws])
            if mode == 1:
                t_l = graph_time([(lambda w=w: torch.nn.functional.silu((x @ w.t())[:, :N // 2]) * (x @ w.t())[:, N // 2:]) for w in Ws[:1]]) if False else \
                    graph_time([(lambda w=w: (lambda gu: torch.nn.functional.silu(gu[:, :N // 2]) * gu[:, N // 2:])(x @ w.t())) for w in Ws])
                t_o = graph_time([(lambda w=w: C.gemm_swiglu(x, w, False)) for w in Ws])
            else:
                t_l = graph_time([(lambda w=w: x @ w.t()) for w in Ws])
                t_o = graph_time([(lambda w=w: C.gemm_bf16(x, w, None, None, 0)) for w in Ws])
            rec = {"op": "decode_gemm_in_graph", "B": B, "N": N, "K": K, "swiglu": bool(mode), "skinny_ms": round(t_s, 4),
                   "cublas_ms": round(t_l, 4), "tile128_ms": round(t_o, 4), "skinny_GBps": round(N * K * 2 / t_s / 1e6, 1),
                   "skinny_hbm_frac_of_measured": round(N * K * 2 / t_s / 1e6 / 6569.6, 3),
                   "cublas_hbm_frac_of_measured": round(N * K * 2 / t_l / 1e6 / 6569.6, 3)}
            print(json.dumps(rec), flush=True)
            if N != 128256:
                tot["skinny"] += t_s; tot["cublas"] += t_l; tot["tile128"] += t_o
            del Ws
        print(json.dumps({"op": "decode_layer_gemms_total_in_graph", "B": B, **{k: round(v, 4) for k, v in tot.items()}}), flush=True)
    # fused swiglu vs unfused
    M, F, K = 16384, 14336, 4096
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    wgu = torch.randn(2 * F, K, device="cuda", dtype=torch.bfloat16) * 0.02
    t_f = bench(lambda: C.gemm_swiglu(x, wgu, False))
    def unfused():
        gu = x @ wgu.t()
        g, u = gu[:, :F], gu[:, F:]
        return torch.nn.functional.silu(g) * u
    t_u = bench(unfused)
    print(json.dumps({"op": "gate_up+swiglu", "fused_ms": round(t_f, 4), "unfused_cublas_ms": round(t_u, 4)}), flush=True)
    # fused linear-CE vs materialised logits
    T, V, H = 16384, 128256, 4096
    h = torch.randn(T, H, device="cuda", dtype=torch.bfloat16) * 0.5
    w = torch.randn(V, H, device="cuda", dtype=torch.bfloat16) * 0.02
    lab = torch.randint(0, V, (T,), device="cuda")
    t_f = bench(lambda: C.linear_ce_fwd(h, w, lab, 1.0), iters=5)
    def mat():
        z = (h @ w.t()).float()
        lse = torch.logsumexp(z, -1)
        return z.gather(-1, lab[:, None]).squeeze(-1) - lse
    t_u = bench(mat, iters=5)
    print(json.dumps({"op": "linear_ce_fwd", "fused_ms": round(t_f, 4), "materialised_ms": round(t_u, 4),
                      "fused_tflops": round(2.0 * T * V * H / t_f / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    sys.exit(main())
