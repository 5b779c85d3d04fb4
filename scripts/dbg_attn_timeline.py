import torch
from polyrl_b200.ops import ext
C = ext()
nq, nkv, d = 32, 8, 128
seq_len, nseq = 2048, 4
T = seq_len * nseq
qkv = (torch.randn(T, (nq + 2 * nkv) * d, device="cuda") * 0.5).to(torch.bfloat16)
q = qkv[:, : nq * d].view(T, nq, d); k = qkv[:, nq * d:(nq + nkv) * d].view(T, nkv, d); v = qkv[:, (nq + nkv) * d:].view(T, nkv, d)
cu = torch.arange(0, T + 1, seq_len, device="cuda", dtype=torch.int32)
dbg = torch.zeros(256, device="cuda", dtype=torch.int64)
for _ in range(3):
    C.attn_fwd(q, k, v, cu, cu, d ** -0.5, True, True, dbg)
torch.cuda.synchronize()
t = dbg.tolist()
t0 = t[0]
print("start->first s_full:", t[1] - t0)
for j in range(12):
    a = t[1 + 5 * j:6 + 5 * j]
    if a[0] == 0: break
    print(f"tile {j}: s_full@{a[0]-t0} pass1 {a[1]-a[0]} pass2 {a[2]-a[1]} wait_PV {a[3]-a[2]} accum {a[4]-a[3]} | total {a[4]-a[0]}")
