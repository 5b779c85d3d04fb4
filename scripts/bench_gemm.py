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
    # decode-shape GEMMs: skinny (swap-AB + split-K) kernel vs cuBLAS vs the 128-row-tile kernel
    ws = torch.empty(24 * 1024 * 1024, device="cuda", dtype=torch.float32)
    ctr = torch.zeros(4096, device="cuda", dtype=torch.int32)
    for B in (64, 128, 256):
        tot = {"skinny": 0.0, "cublas": 0.0}
        for N, K in ((6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336), (128256, 4096)):
            x = torch.randn(B, K, device="cuda", dtype=torch.bfloat16)
            w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02
            mode = 1 if N == 28672 else 0
            t_s = bench(lambda: C.skinny_gemm(x, w, None, None, None, None, ws, ctr, mode, 0))
            if mode == 1:
                def lib():
                    gu = x @ w.t()
                    return torch.nn.functional.silu(gu[:, :N // 2]) * gu[:, N // 2:]
            else:
                lib = lambda: x @ w.t()
            t_l = bench(lib)
            rec = {"op": "decode_gemm", "B": B, "N": N, "K": K, "swiglu": bool(mode), "skinny_ms": round(t_s, 4), "cublas_ms": round(t_l, 4),
                   "skinny_GBps": round(N * K * 2 / t_s / 1e6, 1), "hbm_frac_of_measured": round(N * K * 2 / t_s / 1e6 / 6569.6, 3)}
            print(json.dumps(rec), flush=True)
            if N != 128256:
                tot["skinny"] += t_s; tot["cublas"] += t_l
        print(json.dumps({"op": "decode_layer_gemms_total", "B": B, "skinny_ms": round(tot["skinny"], 4), "cublas_ms": round(tot["cublas"], 4)}), flush=True)
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
