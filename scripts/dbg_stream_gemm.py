import json, os, sys, torch
sys.path.insert(0, "scripts")
from bench_decode_gemm import graph_time
from polyrl_b200.ops import ext
C = ext()
ws = torch.empty(48 * 1024 * 1024, device="cuda", dtype=torch.float32)
ctr = torch.zeros(4096, device="cuda", dtype=torch.int32)
for (B, N, K, mode) in [(128, 4096, 4096, 0), (64, 4096, 4096, 0), (256, 4096, 4096, 0), (128, 28672, 4096, 1)]:
    copies = max(2, int(300e6 // (N * K * 2)) + 1)
    x = torch.randn(B, K, device="cuda", dtype=torch.bfloat16)
    Ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(copies)]
    rec = {"B": B, "N": N, "K": K, "mode": mode}
    for nacc in (1, 2, 4):
        os.environ["PRL_STREAM_NACC"] = str(nacc)
        for dbg in (0, 1, 2):
            os.environ["PRL_STREAM_DBG"] = str(dbg)
            t = graph_time([(lambda w=w: C.stream_gemm(x, w, None, None, None, None, mode, 1)) for w in Ws])
            rec[f"nacc{nacc}_dbg{dbg}_us"] = round(t * 1e3, 2)
    os.environ["PRL_STREAM_DBG"] = "0"
    for nacc in (1, 2, 4):
        os.environ["PRL_STREAM_NACC"] = str(nacc)
        for S in (2, 4):
            if (N + 127) // 128 * S > 148:
                continue
            t = graph_time([(lambda w=w: C.stream_gemm(x, w, None, None, None, None, mode, S)) for w in Ws])
            rec[f"nacc{nacc}_S{S}_us"] = round(t * 1e3, 2)
    print(json.dumps(rec), flush=True)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader", shell=True, capture_output=True, text=True).stdout)
