"""One launch each of the round-2 flagship kernels (CTA-pair GEMM forward, flash attention forward / backward) for an
`ncu --set full` capture:  ncu --set full --clock-control none --import-source on -k regex:'gemm2_sm100|attn_fwd_kernel|attn_bwd_kernel'
-c 3 -o gpurun_out/r2_kernels python scripts/ncu_r2_kernels.py"""
import torch

from polyrl_b200.ops import ext

C = ext()
torch.manual_seed(0)
a = torch.randn(8192, 4096, device="cuda", dtype=torch.bfloat16)
b = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
C.gemm2_bf16(a, b, False, False, None, None, 0)
nq, nkv, d, seq, nseq = 32, 8, 128, 1024, 8
T = seq * nseq
qkv = (torch.randn(T, (nq + 2 * nkv) * d, device="cuda") * 0.5).to(torch.bfloat16)
q = qkv[:, : nq * d].view(T, nq, d)
k = qkv[:, nq * d:(nq + nkv) * d].view(T, nkv, d)
v = qkv[:, (nq + nkv) * d:].view(T, nkv, d)
cu = torch.arange(0, T + 1, seq, device="cuda", dtype=torch.int32)
o, lse = C.attn_fwd(q, k, v, cu, cu, d ** -0.5, True, True)
do = torch.randn_like(o)
C.attn_bwd(q, k, v, o, do, lse, cu, cu, d ** -0.5, True)
torch.cuda.synchronize()
print("ncu_r2_kernels done", flush=True)
