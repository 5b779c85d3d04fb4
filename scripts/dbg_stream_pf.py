import json, os, sys, torch
sys.path.insert(0, "scripts")
from bench_decode_gemm import graph_time
from polyrl_b200.ops import ext
C = ext()
for (B, N, K, mode, cfgs) in [(128, 4096, 4096, 0, [(4, 0), (3, 1), (2, 1)]), (128, 6144, 4096, 0, [(2, 0), (2, 1), (3, 0)]),
                              (128, 28672, 4096, 1, [(1, 0), (1, 1)]), (128, 4096, 14336, 0, [(4, 0), (3, 1), (2, 1)]),
                              (128, 128256, 4096, 0, [(1, 0), (1, 1)])]:
    copies = max(2, int(300e6 // (N * K * 2)) + 1)
    x = torch.randn(B, K, device="cuda", dtype=torch.bfloat16)
    Ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(copies)]
    rec = {"B": B, "N": N, "K": K, "mode": mode}
    if mode:
        rec["cublas_us"] = round(1e3 * graph_time([(lambda w=w: (lambda gu: torch.nn.functional.silu(gu[:, :N // 2]) * gu[:, N // 2:])(x @ w.t())) for w in Ws]), 2)
        rec["r1_tile_us"] = round(1e3 * graph_time([(lambda w=w: C.gemm_swiglu(x, w, False)) for w in Ws]), 2)
    else:
        rec["cublas_us"] = round(1e3 * graph_time([(lambda w=w: x @ w.t()) for w in Ws]), 2)
    for pf in (0, 8, 16, 32, 64):
        os.environ["PRL_STREAM_PF"] = str(pf)
        for (S, pair) in cfgs:
            t = graph_time([(lambda w=w: C.stream_gemm(x, w, None, None, None, None, mode, S, None, pair)) for w in Ws])
            rec[f"pf{pf}_S{S}_{'pair' if pair else 'single'}_us"] = round(t * 1e3, 2)
    print(json.dumps(rec), flush=True)
