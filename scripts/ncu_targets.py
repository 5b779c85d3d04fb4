"""Tiny launch scripts for ncu captures (one kernel family per invocation; see profiles/README.md)."""
import sys

import torch

sys.path.insert(0, ".")
from polyrl_b200.ops import ext  # noqa: E402

C = ext()
what = sys.argv[1]
torch.manual_seed(0)
if what == "gemm":          # training-shape tcgen05 GEMM (gate/up projection, fused SwiGLU)
    x = torch.randn(16384, 4096, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(28672, 4096, device="cuda", dtype=torch.bfloat16) * 0.02
    for _ in range(4):
        C.gemm_swiglu(x, w, False)
elif what == "gemm_plain":
    x = torch.randn(16384, 4096, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16) * 0.02
    for _ in range(4):
        C.gemm_bf16(x, w, None, None, 0)
elif what == "skinny":      # decode-shape GEMM
    x = torch.randn(128, 4096, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16) * 0.02
    ws = torch.empty(24 * 1024 * 1024, device="cuda", dtype=torch.float32)
    ctr = torch.zeros(4096, device="cuda", dtype=torch.int32)
    for _ in range(4):
        C.skinny_gemm(x, w, None, None, None, None, ws, ctr, 0, 0)
elif what == "linear_ce":
    h = torch.randn(16384, 4096, device="cuda", dtype=torch.bfloat16) * 0.5
    w = torch.randn(128256, 4096, device="cuda", dtype=torch.bfloat16) * 0.02
    lab = torch.randint(0, 128256, (16384,), device="cuda")
    for _ in range(3):
        C.linear_ce_fwd(h, w, lab, 1.0)
elif what == "decode_attn":
    B, nq, nkv, d, ctx = 128, 32, 8, 128, 640
    nblk = B * (ctx // 16 + 1) + 1
    kc = torch.randn(nblk, nkv, 16, d, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(nblk, nkv, 16, d, device="cuda", dtype=torch.bfloat16)
    bt = torch.arange(1, 1 + B * (ctx // 16 + 1), device="cuda", dtype=torch.int32).reshape(B, -1)
    sl = torch.full((B,), ctx, device="cuda", dtype=torch.int32)
    q = torch.randn(B, nq * d, device="cuda", dtype=torch.bfloat16)
    for _ in range(4):
        C.paged_decode(q, kc, vc, bt, sl, nq, nkv, d, d ** -0.5, 1, None)
torch.cuda.synchronize()
