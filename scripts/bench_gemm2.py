"""CTA-pair tcgen05 GEMM family (gemm2): correctness in the NT / NN / TN layouts against an fp32 PyTorch
reference, then a micro-benchmark against cuBLAS (torch.matmul) and the round-1 1-CTA kernel on the Llama-3-8B
training shapes.  CUDA-event timing, 3 warm-ups, L2 flushed between iterations (256 MiB memset)."""
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


def operands(M, N, K, a_mn, b_mn, scale=1.0):
    a = (torch.randn((K, M) if a_mn else (M, K), device="cuda") * scale).to(torch.bfloat16)
    b = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * scale).to(torch.bfloat16)
    return a, b


def ref_mm(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A @ B


def check(C, M, N, K, a_mn, b_mn, bn, residual=False):
    a, b = operands(M, N, K, a_mn, b_mn, 0.5)
    r = (torch.randn(M, N, device="cuda")).to(torch.bfloat16) if residual else None
    d = C.gemm2_bf16(a, b, a_mn, b_mn, r, None, bn)
    torch.cuda.synchronize()
    ref = ref_mm(a, b, a_mn, b_mn)
    if residual:
        ref = ref + r.float()
    err = (d.float() - ref).abs().max().item()
    tol = 0.02 * ref.abs().max().item() + 1e-2
    ok = bool(err <= tol) and bool(torch.isfinite(d.float()).all())
    rec = {"check": "gemm2", "M": M, "N": N, "K": K, "a_mn": a_mn, "b_mn": b_mn, "bn": bn, "residual": residual,
           "max_err": round(err, 5), "tol": round(tol, 5), "ok": ok}
    print(json.dumps(rec), flush=True)
    return ok


def main():
    C = ext()
    only_check = "--check" in sys.argv
    ok = True
    for a_mn, b_mn in ((False, False), (False, True), (True, True), (True, False)):
        for bn in (256, 128):
            ok &= check(C, 512, 512, 256, a_mn, b_mn, bn)
            ok &= check(C, 1000, 456 if not b_mn else 448, 200 if not (a_mn or b_mn) else 208, a_mn, b_mn, bn)   # ragged edges
            ok &= check(C, 2048 + 64, 1024 + 64, 1024, a_mn, b_mn, bn, residual=True)
    ok &= check(C, 8192, 6144, 4096, False, False, 256)
    ok &= check(C, 8192, 4096, 6144, False, True, 256)
    ok &= check(C, 6144, 4096, 8192, True, True, 256)
    print(json.dumps({"all_checks_ok": bool(ok)}), flush=True)
    if only_check or not ok:
        return 0 if ok else 1
    T = 16384
    layers = [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336)]   # (out, in) of the Llama-3-8B projections
    for out_f, in_f in layers:
        x = torch.randn(T, in_f, device="cuda", dtype=torch.bfloat16)
        w = torch.randn(out_f, in_f, device="cuda", dtype=torch.bfloat16) * 0.02
        dy = torch.randn(T, out_f, device="cuda", dtype=torch.bfloat16)
        y = torch.empty(T, out_f, device="cuda", dtype=torch.bfloat16)
        dx = torch.empty(T, in_f, device="cuda", dtype=torch.bfloat16)
        dw = torch.empty(out_f, in_f, device="cuda", dtype=torch.bfloat16)
        cases = [
            ("NT fwd", lambda bn: C.gemm2_bf16(x, w, False, False, None, y, bn), lambda: torch.matmul(x, w.t(), out=y),
             lambda: C.gemm_bf16(x, w, None, y, 0), (T, out_f, in_f)),
            ("NN dgrad", lambda bn: C.gemm2_bf16(dy, w, False, True, None, dx, bn), lambda: torch.matmul(dy, w, out=dx), None,
             (T, in_f, out_f)),
            ("TN wgrad", lambda bn: C.gemm2_bf16(dy, x, True, True, None, dw, bn), lambda: torch.matmul(dy.t(), x, out=dw), None,
             (out_f, in_f, T)),
        ]
        for name, own, lib, old, (m, n, k) in cases:
            fl = 2.0 * m * n * k
            t_lib = bench(lib)
            rec = {"op": name, "M": m, "N": n, "K": k, "cublas_ms": round(t_lib, 4), "cublas_tflops": round(fl / t_lib / 1e9, 1)}
            for bn in (256, 128):
                t = bench(lambda: own(bn))
                rec[f"gemm2_bn{bn}_ms"] = round(t, 4)
                rec[f"gemm2_bn{bn}_tflops"] = round(fl / t / 1e9, 1)
            t_auto = bench(lambda: own(0))
            rec["gemm2_auto_ms"] = round(t_auto, 4)
            rec["gemm2_auto_tflops"] = round(fl / t_auto / 1e9, 1)
            rec["gemm2_vs_cublas"] = round(t_lib / t_auto, 3)
            rec["frac_of_measured_burst_1688.7"] = round(fl / t_auto / 1e9 / 1688.7, 3)
            if old is not None:
                t_old = bench(old)
                rec["r1_1cta_ms"] = round(t_old, 4)
                rec["r1_1cta_tflops"] = round(fl / t_old / 1e9, 1)
            print(json.dumps(rec), flush=True)
    for m, n, k in ((8192, 8192, 8192),):
        a, b = operands(m, n, k, False, False)
        d = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        fl = 2.0 * m * n * k
        t_lib = bench(lambda: torch.matmul(a, b.t(), out=d))
        t = bench(lambda: C.gemm2_bf16(a, b, False, False, None, d, 256))
        print(json.dumps({"op": "NT square", "M": m, "N": n, "K": k, "cublas_tflops": round(fl / t_lib / 1e9, 1),
                          "gemm2_tflops": round(fl / t / 1e9, 1), "gemm2_vs_cublas": round(t_lib / t, 3)}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
