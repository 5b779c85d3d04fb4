"""Paged flash-decode microbench (CUDA events around CUDA-graph replays; several layer caches cycled so every call streams
its K/V from HBM):  PYTHONPATH=. python scripts/bench_decode_attn.py"""
import json

import torch

from polyrl_b200.ops import ext, paged_decode_attention


def main():
    C = ext()
    dev = "cuda"
    nq, nkv, d, page = 32, 8, 128, 16
    out_rows = []
    import os

    for B, ctx, splits in [(b, c, sp) for (b, c) in ((128, 428), (128, 640), (256, 640), (32, 2048))
                           for sp in ([0] if os.environ.get("DECODE_ATTN_SPLITS", "") == "" else [int(x) for x in os.environ["DECODE_ATTN_SPLITS"].split(",")])]:
        pages = (ctx + page - 1) // page
        nblk = B * pages + 1
        layers = 4
        kc = [torch.randn(nblk, nkv, page, d, device=dev, dtype=torch.bfloat16) for _ in range(layers)]
        vc = [torch.randn(nblk, nkv, page, d, device=dev, dtype=torch.bfloat16) for _ in range(layers)]
        bt = (torch.randperm(nblk - 1, device=dev)[: B * pages].view(B, pages) + 1).int()
        sl = torch.full((B,), ctx, device=dev, dtype=torch.int32)
        q = torch.randn(B, nq * d, device=dev, dtype=torch.bfloat16)

        def fn():
            for i in range(layers):
                paged_decode_attention(q, kc[i], vc[i], bt, sl, nq, nkv, d, d ** -0.5, max_len=ctx, splits=splits)

        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 20
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (reps * layers) * 1e3
        bytes_ = 2 * B * ctx * nkv * d * 2
        out_rows.append({"B": B, "ctx": ctx, "splits": splits, "us": round(us, 2), "TBps": round(bytes_ / us / 1e6, 3),
                         "frac_of_hbm_copy_6.57": round(bytes_ / us / 1e6 / 6.57, 3)})
        del kc, vc
    for r in out_rows:
        print("DECODE_ATTN " + json.dumps(r), flush=True)


if __name__ == "__main__":
    import os
    import subprocess
    import sys

    if os.environ.get("_DECODE_ATTN_CHILD") == "1":
        main()
    else:                                   # the occupancy variant is latched per process: one child per setting
        for minb in os.environ.get("DECODE_ATTN_MINB", "3").split(","):
            print(f"== POLYRL_DECODE_MINB={minb}", flush=True)
            subprocess.run([sys.executable, __file__], env={**os.environ, "_DECODE_ATTN_CHILD": "1", "POLYRL_DECODE_MINB": minb})
