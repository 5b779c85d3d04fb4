"""GPU check of the pipeline-parallel training engine (torchrun, 4 ranks, NCCL, bf16 kernels): pp=2 x dp=2 and pp=4 with the 1F1B
schedule against a single-process fp32 reference.  Prints PP_GPU_OK ... on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from polyrl_b200.models import TransformerWeights, get_model_config, pack_padded
    from polyrl_b200.models import transformer as tf
    from polyrl_b200.ops import reference as ref
    from polyrl_b200.parallel.pipeline import PipelineEngine

    worst = [0.0, 0.0, 0.0]
    for name, pp, n_layers, n_micro in (("pp2xdp2", 2, 4, 4), ("pp4", 4, 4, 6)):
        if pp > world:
            continue
        dp = world // pp
        pp_group = dp_group = None
        for d in range(dp):
            g = dist.new_group(list(range(d * pp, (d + 1) * pp)))
            if d == rank // pp:
                pp_group = g
        for s in range(pp):
            g = dist.new_group(list(range(s, world, pp)))
            if s == rank % pp:
                dp_group = g
        cfg = get_model_config("tiny-llama-128", num_layers=n_layers)
        eng = PipelineEngine(cfg, dev, pp_group, dp_group if dp > 1 else None, lr=1e-3, weight_decay=0.0, seed=11)

        def micro(dp_rank, k, step, device):
            gen = torch.Generator().manual_seed(1000 * step + 50 * dp_rank + k)
            ids = torch.randint(3, cfg.vocab_size, (4, 40 + 8 * k), generator=gen).to(device)
            inp, _ = pack_padded(ids, torch.ones_like(ids))
            labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), generator=gen).to(device)
            return inp, labels

        w = TransformerWeights(cfg, device="cpu", dtype=torch.float32, init=True, seed=11)
        params = {n: t.requires_grad_(True) for n, t in w.named_tensors()}
        mbs = [micro(rank // pp, k, 0, dev) for k in range(n_micro)]
        losses = eng.train_batch(mbs, lambda lp, ent, k: (-(lp.mean()) + 0.01 * ent.mean()) / n_micro)
        st = eng.step()
        ref_losses = {}
        for d in range(dp):
            for k in range(n_micro):
                ri, rl = micro(d, k, 0, "cpu")
                h = tf.forward_hidden(cfg, w, ri)
                rlp, rent = ref.linear_cross_entropy(h, tf.lm_head_weight(cfg, w.head, w.embed), rl, 1.0)
                loss = (-(rlp.mean()) + 0.01 * rent.mean()) / n_micro
                (loss / dp).backward()
                ref_losses[(d, k)] = loss.item()
        gn = torch.sqrt(sum(p.grad.pow(2).sum() for p in params.values())).item()
        if eng.last:
            worst[0] = max(worst[0], max(abs(losses[k] - ref_losses[(rank // pp, k)]) for k in range(n_micro)))
        worst[1] = max(worst[1], abs(st["grad_norm"] - gn) / gn)
        assert eng.peak_in_flight <= min(pp - eng.stage, n_micro)
    t = torch.tensor(worst, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = t[0].item() < 0.05 and t[1].item() < 0.05
    if rank == 0:
        print(f"PP_GPU_{'OK' if ok else 'FAIL'} loss_err={t[0]:.3e} gn_rel_err={t[1]:.3e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
