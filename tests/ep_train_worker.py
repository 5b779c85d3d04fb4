"""Expert-parallel training equivalence (torchrun 4 ranks, gloo/CPU): EPTrainEngine (experts split over the group, dispatch /
combine all-to-all, dense part data-parallel) tracks a single-process oracle over two optimizer steps."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    from polyrl_b200.models import TransformerWeights, get_model_config, pack_padded
    from polyrl_b200.models import transformer as tf
    from polyrl_b200.ops import reference as ref
    from polyrl_b200.parallel.expert import EPTrainEngine, _exchange, local_expert_range

    # variable all-to-all: rank r sends (r + j + 1) rows to rank j
    send = [rank + j + 1 for j in range(world)]
    recv = [j + rank + 1 for j in range(world)]
    x = torch.cat([torch.full((n, 3), 100.0 * rank + j) for j, n in enumerate(send)])
    y = _exchange(x, send, recv, dist.group.WORLD)
    want = torch.cat([torch.full((n, 3), 100.0 * j + rank) for j, n in enumerate(recv)])
    assert torch.equal(y, want)

    cfg = get_model_config("tiny-qwen3-moe")
    assert local_expert_range(cfg, world, rank) == (2 * rank, 2 * rank + 2)
    eng = EPTrainEngine(cfg, "cpu", dist.group.WORLD, lr=1e-2, weight_decay=0.0, seed=11)
    assert eng.master["layers.0.ewgu"].shape[0] == 2 and eng.master["layers.0.router"].shape[0] == 8

    def batch(r: int, step: int):
        gen = torch.Generator().manual_seed(900 + 10 * step + r)
        ids = torch.randint(3, cfg.vocab_size, (2, 11 + 2 * r), generator=gen)
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), generator=gen)
        return inp, labels

    w = TransformerWeights(cfg, device="cpu", dtype=torch.float32, init=True, seed=11)
    params = {n: t.requires_grad_(True) for n, t in w.named_tensors()}
    m = {n: torch.zeros_like(t) for n, t in params.items()}
    v = {n: torch.zeros_like(t) for n, t in params.items()}
    lp_err = gn_err = 0.0
    for step in range(2):
        inp, labels = batch(rank, step)
        lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
        (-(lp.mean()) + 0.01 * ent.mean()).backward()
        st = eng.step()
        for p in params.values():
            p.grad = None
        mine = None
        for r in range(world):
            ri, rl = batch(r, step)
            h = tf.forward_hidden(cfg, w, ri)
            rlp, rent = ref.linear_cross_entropy(h, tf.lm_head_weight(cfg, w.head, w.embed), rl, 1.0)
            ((-(rlp.mean()) + 0.01 * rent.mean()) / world).backward()
            if r == rank:
                mine = rlp.detach()
        gn = torch.sqrt(sum(p.grad.pow(2).sum() for p in params.values())).item()
        coef = min(1.0, 1.0 / (gn + 1e-6))
        with torch.no_grad():
            for n, p in params.items():
                ref.adamw_step(p, m[n], v[n], p.grad, 1e-2, 0.9, 0.999, 1e-8, 0.0, step + 1, coef)
        lp_err = max(lp_err, (lp.detach() - mine).abs().max().item())
        gn_err = max(gn_err, abs(st["grad_norm"] - gn))
    got = eng.gather_full_state()
    want = {n: t.detach() for n, t in w.named_tensors()}
    assert set(got) == set(want)
    w_err = max(((got[n] - want[n]).norm() / (want[n].norm() + 1e-9)).item() for n in want)
    t = torch.tensor([lp_err, gn_err, w_err])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t[0] < 1e-4 and t[1] < 1e-4 and t[2] < 2e-3, t
    if rank == 0:
        print(f"EP_TRAIN_OK lp_err={t[0]:.2e} gn_err={t[1]:.2e} w_err={t[2]:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
