"""Numerics of every sm_100a kernel vs a plain PyTorch fp32 reference of the same op.

Mirrors the reference's Triton-vs-torch tier (3rdparty/verl/tests/utils/test_linear_cross_entropy.py).
"""
import math

import pytest
import torch

from polyrl_b200.ops import reference as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    from polyrl_b200.ops import ext

    return ext()


def _rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 0), (256, 512, 512, 0), (1000, 1032, 520, 0), (4096, 4096, 4096, 256),
    (256, 6144, 4096, 64), (256, 6144, 4096, 128), (333, 264, 776, 128), (64, 128, 128, 64),
])
def test_gemm_bf16(C, M, N, K, bn):
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    d = C.gemm_bf16(a, b, None, None, bn)
    want = a.float() @ b.float().t()
    assert _rel_err(d, want) < 5e-3


def test_gemm_bf16_residual_and_strided(C):
    torch.manual_seed(1)
    big = torch.randn(512, 1024 + 64, device="cuda", dtype=torch.bfloat16)
    a = big[:, :1024]                       # row pitch != K
    b = torch.randn(768, 1024, device="cuda", dtype=torch.bfloat16)
    r = torch.randn(512, 768, device="cuda", dtype=torch.bfloat16)
    d = C.gemm_bf16(a, b, r, None, 0)
    want = a.float() @ b.float().t() + r.float()
    assert _rel_err(d, want) < 5e-3


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K,bn,res", [
    (512, 512, 256, 256, False), (1000, 456, 208, 0, False), (2112, 1088, 1024, 0, True), (264, 136, 72, 128, False),
    (4096, 4096, 2048, 0, False),
])
def test_gemm2_cta_pair_layouts(C, a_mn, b_mn, M, N, K, bn, res):
    """CTA-pair (cta_group::2) GEMM family vs fp32 matmul: NT (forward), NN (dgrad), TN (wgrad) + the mixed layout;
    ragged M/N/K edges exercise TMA zero-fill on loads and clipping on the TMA store."""
    torch.manual_seed(4)
    if b_mn and N % 8:
        N -= N % 8
    a = (torch.randn((K, M) if a_mn else (M, K), device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.5).to(torch.bfloat16)
    r = torch.randn(M, N, device="cuda", dtype=torch.bfloat16) if res else None
    d = C.gemm2_bf16(a, b, a_mn, b_mn, r, None, bn)
    want = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t())
    if res:
        want = want + r.float()
    assert torch.isfinite(d.float()).all()
    assert _rel_err(d, want) < 5e-3


def test_gemm2_from_fresh_thread(C):
    """Tensor-map encode is a driver call: a thread that has made no CUDA runtime call yet (an autograd worker whose first op is
    one of our GEMMs writing into a cached buffer) must still be able to launch."""
    import threading

    a = torch.randn(160, 512, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(512, 1024, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(160, 1024, device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()
    err: list = []

    def run():
        try:
            C.gemm2_bf16(a, b, False, True, None, out, 0)
        except Exception as e:            # noqa: BLE001
            err.append(e)

    t = threading.Thread(target=run)
    t.start()
    t.join()
    torch.cuda.synchronize()
    assert not err, err
    assert _rel_err(out, a.float() @ b.float()) < 5e-3


def test_gemm2_accum_f32(C):
    torch.manual_seed(5)
    M, N, K = 520, 392, 1032
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(K, N, device="cuda", dtype=torch.bfloat16)
    out = torch.randn(M, N, device="cuda", dtype=torch.float32)
    want = out.clone() + a.float() @ b.float()
    C.gemm2_accum_f32_(out, a, b, False, True, 1.0)
    assert _rel_err(out, want) < 1e-5 + 2e-3
    C.gemm2_accum_f32_(out, a, b, False, True, 0.0)
    assert _rel_err(out, a.float() @ b.float()) < 2e-3


def test_gemm_fp8(C):
    torch.manual_seed(2)
    M, N, K = 256, 1024, 2048
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.05
    aq, a_s = C.quant_fp8_rowwise(a)
    wq, w_s = C.quant_fp8_rowwise(w)
    # quantiser vs torch
    want_q = (a.float() / a_s[:, None]).to(torch.float8_e4m3fn)
    assert (aq.float() - want_q.float()).abs().max().item() <= 32.0   # at most one e4m3 ulp at the top bin
    d = C.gemm_fp8(aq, wq, a_s, w_s, 0)
    want = (aq.float() * a_s[:, None]) @ (wq.float() * w_s[:, None]).t()
    assert _rel_err(d, want) < 5e-3
    assert _rel_err(d, a.float() @ w.float().t()) < 6e-2


def test_gemm_swiglu(C):
    torch.manual_seed(3)
    M, F, K = 300, 1024, 512
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    wg = torch.randn(F, K, device="cuda", dtype=torch.bfloat16) * 0.05
    wu = torch.randn(F, K, device="cuda", dtype=torch.bfloat16) * 0.05
    wgu = ref.interleave_gate_up(wg, wu).contiguous()
    h, gu = C.gemm_swiglu(x, wgu, True)
    g = x.float() @ wg.float().t()
    u = x.float() @ wu.float().t()
    want = torch.nn.functional.silu(g) * u
    assert _rel_err(h, want) < 1e-2
    g2, u2 = ref.split_gate_up(gu)
    assert _rel_err(g2, g) < 5e-3 and _rel_err(u2, u) < 5e-3
    # backward kernel
    dh = torch.randn(M, F, device="cuda", dtype=torch.bfloat16)
    dgu = C.swiglu_bwd(gu, dh)
    guf = gu.float().requires_grad_(True)
    gg, uu = ref.split_gate_up(guf)
    (torch.nn.functional.silu(gg) * uu * dh.float()).sum().backward()
    assert _rel_err(dgu, guf.grad) < 1e-2


@pytest.mark.parametrize("T,V,H,tau", [(256, 4096, 512, 1.0), (700, 32008, 1024, 0.7), (130, 1000, 256, 1.3)])
def test_linear_ce_fwd_bwd(C, T, V, H, tau):
    from polyrl_b200.ops import linear_cross_entropy

    torch.manual_seed(4)
    h = (torch.randn(T, H, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    w = (torch.randn(V, H, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    labels = torch.randint(0, V, (T,), device="cuda")
    lp, ent = linear_cross_entropy(h, w, labels, tau)
    hf = h.detach().float().requires_grad_(True)
    wf = w.detach().float().requires_grad_(True)
    lp_r, ent_r = ref.linear_cross_entropy(hf, wf, labels, tau)
    assert (lp - lp_r).abs().max().item() < 3e-2
    assert (ent - ent_r).abs().max().item() < 3e-2
    glp = torch.randn(T, device="cuda")
    gent = torch.randn(T, device="cuda") * 0.1
    (lp * glp + ent * gent).sum().backward()
    (lp_r * glp + ent_r * gent).sum().backward()
    assert _rel_err(h.grad, hf.grad) < 3e-2
    assert _rel_err(w.grad, wf.grad) < 3e-2


@pytest.mark.parametrize("T,H", [(257, 4096), (64, 768), (1000, 2048), (33, 8192)])
def test_rmsnorm(C, T, H):
    torch.manual_seed(5)
    x = torch.randn(T, H, device="cuda", dtype=torch.bfloat16)
    w = (1 + 0.1 * torch.randn(H, device="cuda")).to(torch.bfloat16)
    y, rstd = C.rmsnorm_fwd(x, w, 1e-6)
    assert _rel_err(y, ref.rmsnorm(x.float(), w.float(), 1e-6)) < 5e-3
    dy = torch.randn(T, H, device="cuda", dtype=torch.bfloat16)
    dx, dw = C.rmsnorm_bwd(x, w, dy, rstd)
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    (ref.rmsnorm(xf, wf, 1e-6) * dy.float()).sum().backward()
    assert _rel_err(dx, xf.grad) < 1e-2
    assert _rel_err(dw, wf.grad) < 1e-2


def test_rope_and_cache(C):
    torch.manual_seed(6)
    T, nq, nkv, d = 100, 8, 2, 128
    qkv = torch.randn(T, (nq + 2 * nkv) * d, device="cuda", dtype=torch.bfloat16)
    pos = torch.randint(0, 500, (T,), device="cuda")
    cos, sin = ref.rope_tables(512, d, 10000.0, device="cuda")
    want = ref.rope(qkv, pos, cos, sin, nq, nkv, d)
    kc = torch.zeros(16, nkv, 16, d, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    slots = torch.randperm(16 * 16, device="cuda")[:T]
    got = qkv.clone()
    C.rope_(got, pos, cos, sin, nq, nkv, d, 1.0, kc, vc, slots, 16)
    assert _rel_err(got, want) < 5e-3
    blk, off = slots // 16, slots % 16
    k_want = want[:, nq * d:(nq + nkv) * d].reshape(T, nkv, d)
    v_want = want[:, (nq + nkv) * d:].reshape(T, nkv, d)
    assert torch.equal(kc[blk, :, off], got[:, nq * d:(nq + nkv) * d].reshape(T, nkv, d))
    assert _rel_err(kc[blk, :, off], k_want) < 5e-3
    assert torch.equal(vc[blk, :, off], v_want)
    # inverse rotation
    back = got.clone()
    C.rope_(back, pos, cos, sin, nq, nkv, d, -1.0, None, None, None, 16)
    assert _rel_err(back, qkv) < 1e-2


def test_adamw_and_grad_accum(C):
    torch.manual_seed(7)
    n = 100003
    master = torch.randn(n, device="cuda")
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    acc = torch.zeros(n, device="cuda")
    p16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    rm, rmm, rv = master.clone(), m.clone(), v.clone()
    for step in range(1, 4):
        g16 = torch.randn(n, device="cuda", dtype=torch.bfloat16)
        g32 = torch.randn(n, device="cuda")
        ss = torch.zeros(1, device="cuda")
        C.grad_accum_(acc, g16, 0.5, None)
        C.grad_accum_(acc, g32, 0.25, ss)
        want_g = 0.5 * g16.float() + 0.25 * g32
        assert torch.allclose(acc, want_g, atol=1e-5)
        assert math.isclose(ss.item(), want_g.pow(2).sum().item(), rel_tol=1e-3)
        coef = min(1.0, 1.0 / (math.sqrt(ss.item()) + 1e-6))
        status = torch.zeros(1, device="cuda", dtype=torch.int32)
        C.adamw_(master, m, v, acc, p16, 1e-2, 0.9, 0.95, 1e-8, 0.1, step, 1.0, ss, status, True)
        ref.adamw_step(rm, rmm, rv, want_g, 1e-2, 0.9, 0.95, 1e-8, 0.1, step, coef)
        assert torch.allclose(master, rm, atol=2e-5, rtol=1e-4)
        assert torch.equal(p16, master.to(torch.bfloat16))
        assert acc.abs().max().item() == 0.0 and status.item() == 0
    # non-finite norm => skipped, grads zeroed
    acc.fill_(1.0)
    before = master.clone()
    ss = torch.full((1,), float("inf"), device="cuda")
    status = torch.zeros(1, device="cuda", dtype=torch.int32)
    C.adamw_(master, m, v, acc, p16, 1e-2, 0.9, 0.95, 1e-8, 0.1, 4, 1.0, ss, status, True)
    assert torch.equal(master, before) and status.item() == 1 and acc.abs().max().item() == 0.0


@pytest.mark.parametrize("nq,nkv,d,splits", [(32, 8, 128, 1), (32, 8, 128, 4), (28, 4, 128, 2), (12, 12, 64, 1), (64, 8, 128, 3)])
def test_paged_decode(C, nq, nkv, d, splits):
    torch.manual_seed(8)
    B, page, max_len = 9, 16, 700
    seq_lens = torch.randint(1, max_len, (B,), device="cuda", dtype=torch.int32)
    seq_lens[0] = 1
    seq_lens[1] = 16
    seq_lens[2] = 17
    nblk_per = (max_len + page - 1) // page
    num_blocks = B * nblk_per + 3
    kc = torch.randn(num_blocks, nkv, page, d, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(num_blocks, nkv, page, d, device="cuda", dtype=torch.bfloat16)
    perm = torch.randperm(num_blocks, device="cuda")[: B * nblk_per].to(torch.int32)
    bt = perm.reshape(B, nblk_per).contiguous()
    q = torch.randn(B, nq * d, device="cuda", dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(d)
    got = C.paged_decode(q, kc, vc, bt, seq_lens, nq, nkv, d, scale, splits, None)
    want = ref.paged_decode_attention(q, kc, vc, bt, seq_lens, nq, nkv, d, scale)
    assert _rel_err(got, want) < 1.5e-2


def test_sampling(C):
    torch.manual_seed(9)
    B, V = 64, 5000
    logits = torch.randn(B, V, device="cuda") * 2
    # greedy
    tok, lp = C.sample(logits, None, None, True)
    assert torch.equal(tok, logits.argmax(-1))
    want_lp = torch.log_softmax(logits, -1).gather(-1, tok[:, None]).squeeze(-1)
    assert torch.allclose(lp, want_lp, atol=1e-3)
    # sampled: log-prob consistent with temperature-scaled softmax, distribution roughly right
    inv_t = torch.full((B,), 1 / 0.8, device="cuda")
    small = torch.tensor([[2.0, 1.0, 0.0, -1.0]], device="cuda").repeat(20000, 1)
    seeds = torch.arange(20000, device="cuda", dtype=torch.int64) * 7919 + 13
    tok, lp = C.sample(small, torch.ones(20000, device="cuda"), seeds, False)
    freq = torch.bincount(tok, minlength=4).float() / 20000
    assert (freq - torch.softmax(small[0], -1)).abs().max().item() < 0.02
    seeds = torch.randint(0, 2 ** 40, (B,), device="cuda")
    tok, lp = C.sample(logits.to(torch.bfloat16), inv_t, seeds, False)
    want_lp = torch.log_softmax(logits.to(torch.bfloat16).float() / 0.8, -1).gather(-1, tok[:, None]).squeeze(-1)
    assert torch.allclose(lp, want_lp, atol=2e-3)


def _ref_attn(q, k, v, cu_q, cu_k):
    nq, nkv = q.shape[1], k.shape[1]
    out = torch.empty_like(q, dtype=torch.float32)
    cq, ck = cu_q.tolist(), cu_k.tolist()
    for i in range(len(cq) - 1):
        qs, ks, vs = q[cq[i]:cq[i + 1]].float(), k[ck[i]:ck[i + 1]].float(), v[ck[i]:ck[i + 1]].float()
        lq, lk = qs.shape[0], ks.shape[0]
        ks, vs = ks.repeat_interleave(nq // nkv, dim=1), vs.repeat_interleave(nq // nkv, dim=1)
        s = torch.einsum("qhd,khd->hqk", qs, ks) * (q.shape[-1] ** -0.5)
        mask = torch.arange(lk, device=q.device)[None, :] <= (torch.arange(lq, device=q.device)[:, None] + (lk - lq))
        out[cq[i]:cq[i + 1]] = torch.einsum("hqk,khd->qhd", torch.softmax(s.masked_fill(~mask[None], float("-inf")), -1), vs)
    return out


@pytest.mark.parametrize("lq,lk,nq,nkv", [
    ([640, 700, 33, 128, 1, 257, 129], None, 8, 2),           # packed self-attention, ragged tile edges, GQA 4:1
    ([64, 1, 200, 128], [364, 18, 328, 1152], 4, 4),          # prefill behind a cached prefix (bottom-right aligned causal)
    ([1500], None, 4, 1),
])
def test_tcgen05_flash_attention_fwd_bwd(lq, lk, nq, nkv):
    """Own tcgen05 flash-attention forward + backward vs an fp32 PyTorch reference (through the public op, autograd)."""
    from polyrl_b200.ops.attention import attention_varlen_prefix

    torch.manual_seed(30)
    lk = lk or lq
    cu_q = torch.tensor([0] + list(torch.tensor(lq).cumsum(0)), device="cuda", dtype=torch.int32)
    cu_k = torch.tensor([0] + list(torch.tensor(lk).cumsum(0)), device="cuda", dtype=torch.int32)
    Tq, Tk, d = int(cu_q[-1]), int(cu_k[-1]), 128
    q = (torch.randn(Tq, nq, d, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    k = (torch.randn(Tk, nkv, d, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    v = (torch.randn(Tk, nkv, d, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    do = (torch.randn(Tq, nq, d, device="cuda") * 0.5).to(torch.bfloat16)
    o = attention_varlen_prefix(q, k, v, cu_q, cu_k, max(lq), max(lk))
    o.backward(do)
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    want = _ref_attn(qf, kf, vf, cu_q, cu_k)
    want.backward(do.float())
    assert _rel_err(o, want) < 5e-3
    assert _rel_err(q.grad, qf.grad) < 1e-2 and _rel_err(k.grad, kf.grad) < 1e-2 and _rel_err(v.grad, vf.grad) < 1e-2
    with torch.no_grad():                      # inference path (no lse, strided views of a packed qkv row)
        qkv = torch.cat([q.detach().reshape(Tq, -1), k.detach()[:Tq].reshape(Tq, -1) if Tk >= Tq else k.detach().reshape(Tk, -1)], dim=1)
        o2 = attention_varlen_prefix(q.detach(), k.detach(), v.detach(), cu_q, cu_k, max(lq), max(lk))
        assert torch.equal(o2, o.detach()) and qkv.is_cuda


def _ref_filter(logits, temp, top_k, top_p):
    z = logits.float() / temp
    if top_k > 0:
        kth = torch.topk(z, min(top_k, z.shape[-1]), dim=-1).values[..., -1:]
        z = z.masked_fill(z < kth, float("-inf"))
    if top_p < 1.0:
        sz, si = torch.sort(z, dim=-1, descending=True)
        pr = torch.softmax(sz, dim=-1)
        drop = pr.cumsum(-1) - pr >= top_p
        sz = sz.masked_fill(drop, float("-inf"))
        kept = torch.full_like(z, float("-inf")).scatter(-1, si, sz)
        # ties at the nucleus boundary: a sort cuts a group of equal logits at an arbitrary member, the kernel keeps the whole
        # group (bf16 logits tie often) -> make the reference tie-inclusive
        thr = torch.where(torch.isfinite(kept), kept, torch.full_like(kept, float("inf"))).amin(-1, keepdim=True)
        z = z.masked_fill(z < thr, float("-inf"))
    return z


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("top_k,top_p", [(50, 1.0), (0, 0.9), (40, 0.8), (1, 1.0), (0, 0.2)])
def test_sampling_topk_topp_kernel(C, dtype, top_k, top_p):
    """Radix-select top-k / top-p sampler vs the sort-based reference: sampled token inside the reference kept set, log-prob equal
    to the renormalised log-softmax, greedy == argmax, and empirical frequencies on a small vocabulary."""
    torch.manual_seed(12)
    B, V, temp = 64, 32000, 0.7
    logits = (torch.randn(B, V, device="cuda") * 3).to(dtype)
    inv_t = torch.full((B,), 1.0 / temp, device="cuda")
    seeds = torch.randint(0, 2 ** 40, (B,), device="cuda")
    tok, lp = C.sample_topk_topp(logits, inv_t, seeds, False, top_k, top_p)
    z = _ref_filter(logits, temp, top_k, top_p)
    want_lp = torch.log_softmax(z, -1).gather(-1, tok[:, None]).squeeze(-1)
    assert torch.isfinite(want_lp).all(), "sampled a token outside the reference kept set"
    assert torch.allclose(lp, want_lp, atol=3e-3), (lp - want_lp).abs().max()
    # same seeds -> same tokens (tensor-parallel ranks must agree)
    tok2, _ = C.sample_topk_topp(logits, inv_t, seeds, False, top_k, top_p)
    assert torch.equal(tok, tok2)
    gtok, _ = C.sample_topk_topp(logits, inv_t, seeds, True, top_k, top_p)
    assert torch.equal(gtok, logits.float().argmax(-1))
    # frequencies: 6-token vocabulary, nucleus keeps a strict subset
    small = torch.tensor([[2.0, 1.5, 1.0, 0.0, -1.0, -3.0]], device="cuda").repeat(20000, 1).to(dtype)
    sd = torch.arange(20000, device="cuda", dtype=torch.int64) * 7919 + 3
    t2, _ = C.sample_topk_topp(small, torch.ones(20000, device="cuda"), sd, False, min(top_k, 4) if top_k else 0, top_p)
    zr = _ref_filter(small[:1], 1.0, min(top_k, 4) if top_k else 0, top_p)
    want = torch.softmax(zr[0], -1)
    freq = torch.bincount(t2, minlength=6).float() / 20000
    assert (freq - want).abs().max().item() < 0.02, (freq, want)


def test_functional_autograd_linear():
    from polyrl_b200 import ops

    torch.manual_seed(10)
    x = torch.randn(3, 50, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w = (torch.randn(512, 256, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    r = torch.randn(3, 50, 512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = ops.linear(x, w, residual=r)
    y.float().pow(2).sum().backward()
    xf, wf, rf = (t.detach().float().requires_grad_(True) for t in (x, w, r))
    (xf @ wf.t() + rf).pow(2).sum().backward()
    assert _rel_err(x.grad, xf.grad) < 2e-2 and _rel_err(w.grad, wf.grad) < 2e-2 and _rel_err(r.grad, rf.grad) < 2e-2


@pytest.mark.parametrize("B,N,K,splits", [(128, 4096, 4096, 0), (128, 6144, 4096, 4), (7, 512, 256, 1), (256, 1024, 14336, 0),
                                          (33, 1000, 4096, 3), (128, 128256, 1024, 0), (200, 4096, 4096, 1)])
def test_skinny_gemm(C, B, N, K, splits):
    torch.manual_seed(20)
    x = torch.randn(B, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    r = torch.randn(B, N, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    ws = torch.empty(24 * 1024 * 1024, device="cuda", dtype=torch.float32)
    ctr = torch.zeros(4096, device="cuda", dtype=torch.int32)
    want = x.float() @ w.float().t()
    got = C.skinny_gemm(x, w, None, None, None, None, ws, ctr, 0, splits)
    assert _rel_err(got, want) < 5e-3
    got = C.skinny_gemm(x, w, r, bias, None, None, ws, ctr, 0, splits)          # counters must have self-reset
    assert _rel_err(got, want + bias.float() + r.float()) < 5e-3
    assert int(ctr.abs().sum().item()) == 0


@pytest.mark.parametrize("B,N,K,mode,splits,pair", [
    (128, 4096, 4096, 0, 4, 0), (128, 6144, 4096, 0, 3, 0), (33, 1000, 4096, 0, 5, 0), (7, 512, 256, 0, 2, 0),
    (96, 2048, 1024, 1, 3, 0), (128, 4096, 4096, 0, 4, 1), (40, 512, 256, 0, 2, 1), (96, 2048, 1024, 1, 3, 1),
    (256, 4096, 4096, 0, 0, -1), (128, 28672, 1024, 1, 0, -1),
])
def test_stream_gemm_cluster_splitk(C, B, N, K, mode, splits, pair):
    """Decode stream GEMM (cluster split-K, DSMEM reduce-scatter, single CTAs and CTA pairs) vs fp32 matmul."""
    torch.manual_seed(22)
    x = torch.randn(B, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    z = x.float() @ w.float().t()
    if mode:
        got = C.stream_gemm(x, w, None, None, None, None, 1, splits, None, pair)
        want = ref.swiglu(z.to(torch.bfloat16)).float() if False else (lambda g, u: torch.nn.functional.silu(g) * u)(*ref.split_gate_up(z))
    else:
        r = torch.randn(B, N, device="cuda", dtype=torch.bfloat16)
        bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
        got = C.stream_gemm(x, w, r, bias, None, None, 0, splits, None, pair)
        want = z + bias.float() + r.float()
    assert _rel_err(got, want) < 6e-3


def test_stream_gemm_fp8(C):
    torch.manual_seed(23)
    B, N, K = 96, 2048, 1024
    x = torch.randn(B, K, device="cuda", dtype=torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    xq, xs = C.quant_fp8_rowwise(x)
    wq, wsc = C.quant_fp8_rowwise(w)
    got = C.stream_gemm(xq, wq, None, None, xs, wsc, 0, 2, None, 0)
    want = (xq.float() * xs[:, None]) @ (wq.float() * wsc[:, None]).t()
    assert _rel_err(got, want) < 6e-3


def test_skinny_gemm_swiglu_and_fp8(C):
    torch.manual_seed(21)
    B, F, K = 96, 2048, 1024
    x = torch.randn(B, K, device="cuda", dtype=torch.bfloat16)
    wg = (torch.randn(F, K, device="cuda") * 0.05).to(torch.bfloat16)
    wu = (torch.randn(F, K, device="cuda") * 0.05).to(torch.bfloat16)
    wgu = ref.interleave_gate_up(wg, wu).contiguous()
    ws = torch.empty(24 * 1024 * 1024, device="cuda", dtype=torch.float32)
    ctr = torch.zeros(4096, device="cuda", dtype=torch.int32)
    want = torch.nn.functional.silu(x.float() @ wg.float().t()) * (x.float() @ wu.float().t())
    for splits in (1, 0, 4):
        got = C.skinny_gemm(x, wgu, None, None, None, None, ws, ctr, 1, splits)
        assert got.shape == (B, F) and _rel_err(got, want) < 1e-2
    xq, xs = C.quant_fp8_rowwise(x)
    wq, wsc = C.quant_fp8_rowwise(wg)
    got = C.skinny_gemm(xq, wq, None, None, xs, wsc, ws, ctr, 0, 0)
    want8 = (xq.float() * xs[:, None]) @ (wq.float() * wsc[:, None]).t()
    assert _rel_err(got, want8) < 5e-3


def test_add_rmsnorm(C):
    torch.manual_seed(22)
    x = torch.randn(130, 4096, device="cuda", dtype=torch.bfloat16)
    r = torch.randn(130, 4096, device="cuda", dtype=torch.bfloat16)
    w = (1 + 0.1 * torch.randn(4096, device="cuda")).to(torch.bfloat16)
    xs = (x.float() + r.float()).to(torch.bfloat16)
    want = ref.rmsnorm(xs.float(), w.float(), 1e-5)
    y = C.add_rmsnorm_(x, r, w, 1e-5)
    assert torch.equal(x, xs) and _rel_err(y, want) < 5e-3
