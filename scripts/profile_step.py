"""Kernel-time breakdown (torch.profiler / CUPTI) of (a) one ZeRO-3 training micro-batch (fwd + recompute + bwd) and
(b) decode steps of the rollout engine, Llama-3-8B shapes.  Shares only — numbers under a profiler are not bench values."""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polyrl_b200.models import get_model_config, pack_padded  # noqa: E402
from polyrl_b200.parallel.zero3 import Zero3Engine  # noqa: E402
from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams  # noqa: E402


def top(prof, n=22):
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
        if t > 0 and e.device_type.name == "CUDA":
            rows.append((t, e.count, e.key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    out = [{"share": round(r[0] / total, 4), "ms": round(r[0] / 1e3, 3), "calls": r[1], "kernel": r[2][:110]} for r in rows[:n]]
    return total / 1e3, out


def main():
    model = os.environ.get("PROF_MODEL", "llama3-8b")
    layers = int(os.environ.get("PROF_LAYERS", "8"))
    cfg = get_model_config(model, num_layers=layers)
    dev = torch.device("cuda:0")
    what = sys.argv[1] if len(sys.argv) > 1 else "both"
    if what in ("train", "both"):
        eng = Zero3Engine(cfg, dev, lr=1e-6)
        b, s = 24, 680
        ids = torch.randint(16, cfg.vocab_size, (b, s), device=dev)
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), device=dev)

        def step():
            lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
            (-(lp.mean())).backward()

        step(); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step(); torch.cuda.synchronize()
        total, rows = top(prof)
        print(json.dumps({"phase": "train_microbatch", "tokens": int(inp.input_ids.numel()), "layers": layers, "gpu_ms": round(total, 2)}))
        for r in rows:
            print(json.dumps(r))
        del eng
        torch.cuda.empty_cache()
    if what in ("decode", "both"):
        from polyrl_b200.models import TransformerWeights

        w = TransformerWeights(cfg, device=dev, dtype=torch.bfloat16)
        B = int(os.environ.get("PROF_BATCH", "128"))
        eng = RolloutEngine(cfg, dev, max_num_seqs=B, max_model_len=1024, kv_cache_tokens=B * 1024, use_cuda_graph=False)
        eng.set_weights(w.embed, w.layers, w.head)
        for i in range(B):
            eng.add_request(list(range(20, 148)), SamplingParams(max_new_tokens=400, ignore_eos=True))
        for _ in range(300):
            eng.step()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(4):
                eng.step()
            torch.cuda.synchronize()
        total, rows = top(prof)
        print(json.dumps({"phase": "decode_4_steps", "batch": B, "ctx": 428, "layers": layers, "gpu_ms": round(total, 2)}))
        for r in rows:
            print(json.dumps(r))


if __name__ == "__main__":
    main()
