"""tcgen05 varlen flash-attention (csrc/attn_fwd.cu, attn_bwd.cu) vs an fp32 PyTorch reference, and timing vs the flash-attn 2
library at the training shapes (Llama-3-8B: 32 query / 8 kv heads, d=128, ~16K packed tokens of 640-1280-token sequences)."""
import json
import sys

import torch

from polyrl_b200.ops import ext


def ref_attn(q, k, v, cu_q, cu_k, causal=True):
    """fp32 reference, sequence by sequence, bottom-right aligned causal mask."""
    nq, nkv = q.shape[1], k.shape[1]
    out = torch.empty_like(q, dtype=torch.float32)
    cq, ck = cu_q.tolist(), cu_k.tolist()
    for i in range(len(cq) - 1):
        qs, ks, vs = q[cq[i]:cq[i + 1]].float(), k[ck[i]:ck[i + 1]].float(), v[ck[i]:ck[i + 1]].float()
        lq, lk = qs.shape[0], ks.shape[0]
        if lq == 0:
            continue
        ks, vs = ks.repeat_interleave(nq // nkv, dim=1), vs.repeat_interleave(nq // nkv, dim=1)
        s = torch.einsum("qhd,khd->hqk", qs, ks) * (q.shape[-1] ** -0.5)
        if causal:
            mask = torch.arange(lk, device=q.device)[None, :] <= (torch.arange(lq, device=q.device)[:, None] + (lk - lq))
            s = s.masked_fill(~mask[None], float("-inf"))
        out[cq[i]:cq[i + 1]] = torch.einsum("hqk,khd->qhd", torch.softmax(s, -1), vs)
    return out


def bench(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


def main():
    C = ext()
    torch.manual_seed(0)
    nq, nkv, d = 32, 8, 128
    ok = True
    for name, lq_list, lk_list in [
        ("self", [640, 700, 33, 128, 1, 257, 129], None),
        ("prefix", [64, 1, 200, 128], [64 + 300, 1 + 17, 200 + 128, 128 + 1024]),
        ("one_long", [1500], None),
    ]:
        lk_list = lk_list or lq_list
        cu_q = torch.tensor([0] + list(torch.tensor(lq_list).cumsum(0)), device="cuda", dtype=torch.int32)
        cu_k = torch.tensor([0] + list(torch.tensor(lk_list).cumsum(0)), device="cuda", dtype=torch.int32)
        Tq, Tk = int(cu_q[-1]), int(cu_k[-1])
        # q / k / v as strided views of one packed row (how the model produces them) for the self-attention case
        if lk_list is lq_list:
            qkv = (torch.randn(Tq, (nq + 2 * nkv) * d, device="cuda") * 0.5).to(torch.bfloat16)
            q = qkv[:, : nq * d].view(Tq, nq, d)
            k = qkv[:, nq * d:(nq + nkv) * d].view(Tq, nkv, d)
            v = qkv[:, (nq + nkv) * d:].view(Tq, nkv, d)
        else:
            q = (torch.randn(Tq, nq, d, device="cuda") * 0.5).to(torch.bfloat16)
            k = (torch.randn(Tk, nkv, d, device="cuda") * 0.5).to(torch.bfloat16)
            v = (torch.randn(Tk, nkv, d, device="cuda") * 0.5).to(torch.bfloat16)
        o, lse = C.attn_fwd(q, k, v, cu_q, cu_k, d ** -0.5, True, True)
        torch.cuda.synchronize()
        want = ref_attn(q, k, v, cu_q, cu_k)
        err = (o.float() - want).abs().max().item()
        rel = ((o.float() - want).norm() / want.norm()).item()
        good = rel < 1e-2 and bool(torch.isfinite(o.float()).all())
        ok &= good
        print(json.dumps({"check": "attn_fwd", "case": name, "Tq": Tq, "Tk": Tk, "max_abs_err": round(err, 5), "rel_err": round(rel, 5),
                          "ok": bool(good)}), flush=True)
    # ---------------- backward vs autograd of the fp32 reference
    for name, lq_list, lk_list in [("self", [640, 700, 33, 128, 1, 257, 129], None), ("prefix", [64, 1, 200, 128], [364, 18, 328, 1152])]:
        lk_list = lk_list or lq_list
        cu_q = torch.tensor([0] + list(torch.tensor(lq_list).cumsum(0)), device="cuda", dtype=torch.int32)
        cu_k = torch.tensor([0] + list(torch.tensor(lk_list).cumsum(0)), device="cuda", dtype=torch.int32)
        Tq, Tk = int(cu_q[-1]), int(cu_k[-1])
        q = (torch.randn(Tq, nq, d, device="cuda") * 0.5).to(torch.bfloat16)
        k = (torch.randn(Tk, nkv, d, device="cuda") * 0.5).to(torch.bfloat16)
        v = (torch.randn(Tk, nkv, d, device="cuda") * 0.5).to(torch.bfloat16)
        do = (torch.randn(Tq, nq, d, device="cuda") * 0.5).to(torch.bfloat16)
        o, lse = C.attn_fwd(q, k, v, cu_q, cu_k, d ** -0.5, True, True)
        dq, dk, dv = C.attn_bwd(q, k, v, o, do, lse, cu_q, cu_k, d ** -0.5, True)
        torch.cuda.synchronize()
        qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
        ref_attn(qf, kf, vf, cu_q, cu_k).backward(do.float())
        errs = {n: ((a.float() - b.grad).norm() / (b.grad.norm() + 1e-12)).item() for n, a, b in (("dq", dq, qf), ("dk", dk, kf), ("dv", dv, vf))}
        good = all(e < 2e-2 for e in errs.values()) and all(bool(torch.isfinite(t.float()).all()) for t in (dq, dk, dv))
        ok &= good
        print(json.dumps({"check": "attn_bwd", "case": name, **{k_: round(v_, 5) for k_, v_ in errs.items()}, "ok": bool(good)}), flush=True)
    print(json.dumps({"all_checks_ok": bool(ok)}), flush=True)
    if not ok or "--check" in sys.argv:
        return 0 if ok else 1
    from flash_attn import flash_attn_varlen_func

    for seq_len, nseq in ((680, 24), (1280, 13), (4096, 4)):
        T = seq_len * nseq
        qkv = (torch.randn(T, (nq + 2 * nkv) * d, device="cuda") * 0.5).to(torch.bfloat16)
        q = qkv[:, : nq * d].view(T, nq, d)
        k = qkv[:, nq * d:(nq + nkv) * d].view(T, nkv, d)
        v = qkv[:, (nq + nkv) * d:].view(T, nkv, d)
        cu = torch.arange(0, T + 1, seq_len, device="cuda", dtype=torch.int32)
        t_own = bench(lambda: C.attn_fwd(q, k, v, cu, cu, d ** -0.5, True, True))
        t_fa = bench(lambda: flash_attn_varlen_func(q, k, v, cu, cu, seq_len, seq_len, 0.0, softmax_scale=d ** -0.5, causal=True))
        flops = 4.0 * nseq * nq * d * seq_len * seq_len / 2          # causal: half of the score matrix
        o, lse = C.attn_fwd(q, k, v, cu, cu, d ** -0.5, True, True)
        do = torch.randn_like(o)
        t_bwd = bench(lambda: C.attn_bwd(q, k, v, o, do, lse, cu, cu, d ** -0.5, True))
        qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        of = flash_attn_varlen_func(qg, kg, vg, cu, cu, seq_len, seq_len, 0.0, softmax_scale=d ** -0.5, causal=True)
        t_fab = bench(lambda: torch.autograd.grad(of, (qg, kg, vg), do, retain_graph=True))
        print(json.dumps({"op": "attn_bwd", "seq_len": seq_len, "nseq": nseq, "own_ms": round(t_bwd, 4), "flash_attn2_ms": round(t_fab, 4),
                          "own_tflops": round(2.5 * flops / t_bwd / 1e9, 1), "fa2_tflops": round(2.5 * flops / t_fab / 1e9, 1),
                          "own_vs_fa2": round(t_fab / t_bwd, 3)}), flush=True)
        print(json.dumps({"op": "attn_fwd", "seq_len": seq_len, "nseq": nseq, "tokens": T, "own_ms": round(t_own, 4), "flash_attn2_ms": round(t_fa, 4),
                          "own_tflops": round(flops / t_own / 1e9, 1), "fa2_tflops": round(flops / t_fa / 1e9, 1),
                          "own_vs_fa2": round(t_fa / t_own, 3)}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
