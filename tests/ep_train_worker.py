"""Expert-parallel training equivalence (torchrun 4 ranks, gloo/CPU): EPTrainEngine (experts split over the group, dispatch /
combine all-to-all, dense part data-parallel) tracks a single-process oracle over two optimizer steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    cuda = torch.cuda.is_available() and os.environ.get("EP_TRAIN_DEVICE", "cpu") == "cuda"
    if cuda:                                   # GPU check (next-round verification): NCCL all_to_all_single, or EP_PEER=1 peer writes
        dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    from polyrl_b200.models import TransformerWeights, get_model_config, pack_padded
    from polyrl_b200.models import transformer as tf
    from polyrl_b200.ops import reference as ref
    from polyrl_b200.parallel.expert import EPTrainEngine, _exchange, local_expert_range

    # variable all-to-all: rank r sends (r + j + 1) rows to rank j
    send = [rank + j + 1 for j in range(world)]
    recv = [j + rank + 1 for j in range(world)]
    cfg = get_model_config(os.environ.get("EP_TRAIN_MODEL") or ("tiny-qwen3-moe-128" if cuda else "tiny-qwen3-moe"))
    if cuda and os.environ.get("EP_PEER") == "1":
        from polyrl_b200.parallel.expert import enable_peer_dispatch
        from polyrl_b200.parallel.symm import SymmetricHeap

        enable_peer_dispatch(dist.group.WORLD, SymmetricHeap(group=None, device=dev), 4096, cfg.hidden_size, 4)
    x = torch.cat([torch.full((n, 512), 100.0 * rank + j, device=dev) for j, n in enumerate(send)])
    y = _exchange(x, send, recv, dist.group.WORLD)
    want = torch.cat([torch.full((n, 512), 100.0 * j + rank, device=dev) for j, n in enumerate(recv)])
    assert torch.equal(y, want)

    assert local_expert_range(cfg, world, rank) == (2 * rank, 2 * rank + 2)
    lr = 1e-3 if cuda else 1e-2
    eng = EPTrainEngine(cfg, dev, dist.group.WORLD, lr=lr, weight_decay=0.0, seed=11)
    sl = next(i for i in range(cfg.num_layers) if i not in cfg.dense_layers)            # first sparse block
    assert eng.master[f"layers.{sl}.ewgu"].shape[0] == 2 and eng.master[f"layers.{sl}.router"].shape[0] == 8

    def batch(r: int, step: int):
        gen = torch.Generator().manual_seed(900 + 10 * step + r)
        ids = torch.randint(3, cfg.vocab_size, (2, 11 + 2 * r), generator=gen).to(dev)
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), generator=gen).to(dev)
        return inp, labels

    w_cpu = TransformerWeights(cfg, device="cpu", dtype=torch.float32, init=True, seed=11)
    w = TransformerWeights(cfg, device=dev, dtype=torch.float32, init=False)
    w.load_state_dict({n: t.to(dev) for n, t in w_cpu.state_dict().items()})
    params = {n: t.requires_grad_(True) for n, t in w.named_tensors()}
    # On the GPU the oracle runs the same bf16 kernels from a bf16 copy of its fp32 master weights (what the engine does), without
    # expert parallelism: the comparison then isolates dispatch / combine / gradient reduction from bf16-vs-fp32 rounding.
    cw = w
    if cuda:
        cw = TransformerWeights(cfg, device=dev, dtype=torch.bfloat16, init=False)
    cparams = {n: t.requires_grad_(True) for n, t in cw.named_tensors()} if cuda else params
    m = {n: torch.zeros_like(t) for n, t in params.items()}
    v = {n: torch.zeros_like(t) for n, t in params.items()}
    lp_err = gn_err = 0.0
    for step in range(2):
        inp, labels = batch(rank, step)
        lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
        (-(lp.mean()) + 0.01 * ent.mean()).backward()
        st = eng.step()
        for p in cparams.values():
            p.grad = None
        if cuda:
            with torch.no_grad():
                for n, p in cparams.items():
                    p.copy_(params[n])
        mine = None
        for r in range(world):
            ri, rl = batch(r, step)
            h = tf.forward_hidden(cfg, cw, ri)
            if cuda:
                from polyrl_b200 import ops
                rlp, rent = ops.linear_cross_entropy(h, tf.lm_head_weight(cfg, cw.head, cw.embed), rl, 1.0)
            else:
                rlp, rent = ref.linear_cross_entropy(h, tf.lm_head_weight(cfg, cw.head, cw.embed), rl, 1.0)
            ((-(rlp.mean()) + 0.01 * rent.mean()) / world).backward()
            if r == rank:
                mine = rlp.detach().float()
        # parameters that only steer discrete choices (the router's selection bias) receive no gradient
        grads = {n: (cparams[n].grad.float() if cparams[n].grad is not None else torch.zeros_like(params[n])) for n in params}
        gn = torch.sqrt(sum(g.pow(2).sum() for g in grads.values())).item()
        coef = min(1.0, 1.0 / (gn + 1e-6))
        with torch.no_grad():
            for n, p in params.items():
                ref.adamw_step(p, m[n], v[n], grads[n], lr, 0.9, 0.999, 1e-8, 0.0, step + 1, coef)
        lp_err = max(lp_err, (lp.detach().float() - mine).abs().max().item())
        if os.environ.get("EP_TRAIN_DEBUG") and rank == 0:
            for n in ("embed", "layers.0.wqkv", "layers.1.swgu", "layers.1.router"):
                if n in grads and n in eng.last_grads:
                    d = (eng.last_grads[n] - grads[n]).abs()
                    small = (grads[n].abs() < 1e-7).float().mean().item()
                    print(f"EP_TRAIN_DEBUG   {n}: max|dg|={d.max().item():.2e} max|g|={grads[n].abs().max().item():.2e} frac|g|<1e-7={small:.4f} nonzero={(grads[n] != 0).float().mean().item():.4f}", flush=True)
            print(f"EP_TRAIN_DEBUG step={step} lp_err={(lp.detach().float() - mine).abs().max().item():.3e} gn={gn:.5f} eng_gn={st['grad_norm']:.5f}", flush=True)
        gn_err = max(gn_err, abs(st["grad_norm"] - gn))
    got = eng.gather_full_state()
    want = {n: t.detach() for n, t in w.named_tensors()}
    assert set(got) == set(want)
    errs = {n: ((got[n].float() - want[n]).norm() / (want[n].norm() + 1e-9)).item() for n in want}
    worst = max(errs, key=errs.get)
    w_err = errs[worst]
    t = torch.tensor([lp_err, gn_err, w_err], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # GPU: same-precision oracle; what is left is summation order (atomics, all-reduce) and Adam amplifying it where |g| ~ noise
    tol = (0.2, 2e-2, 0.05) if cuda else (1e-4, 1e-4, 2e-3)
    if os.environ.get("EP_TRAIN_TOL"):
        tol = tuple(float(x) for x in os.environ["EP_TRAIN_TOL"].split(","))
    if rank == 0:
        print(f"EP_TRAIN lp_err={t[0]:.3e} gn_err={t[1]:.3e} w_err={t[2]:.3e} worst={worst}", flush=True)
    assert t[0] < tol[0] and t[1] < tol[1] and t[2] < tol[2], t
    if rank == 0:
        print(f"EP_TRAIN_OK lp_err={t[0]:.2e} gn_err={t[1]:.2e} w_err={t[2]:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
