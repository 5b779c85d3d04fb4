"""Vocab-parallel linear-CE (torchrun 2 ranks; gloo/CPU or nccl/GPU): TP result and grads vs the unsharded op."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    cuda = torch.cuda.is_available() and os.environ.get("CE_TP_DEVICE", "auto") != "cpu"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if cuda:
        dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    from polyrl_b200 import ops

    dt = torch.bfloat16 if cuda else torch.float32
    T, H, V = (300, 256, 4096) if cuda else (37, 32, 96)
    torch.manual_seed(3)
    h = (torch.randn(T, H, device=dev) * 0.5).to(dt)
    w = (torch.randn(V, H, device=dev) * 0.1).to(dt)
    labels = torch.randint(0, V, (T,), device=dev)
    g1, g2 = torch.randn(T, device=dev), torch.randn(T, device=dev) * 0.1
    tau = 0.7
    # oracle: fp32 torch on the full weight
    hf, wf = h.detach().float().clone().requires_grad_(True), w.detach().float().clone().requires_grad_(True)
    from polyrl_b200.ops import reference as ref

    lp0, e0 = ref.linear_cross_entropy(hf, wf, labels, tau)
    ((lp0 * g1).sum() + (e0 * g2).sum()).backward()
    vs = V // world
    hs = h.detach().clone().requires_grad_(True)
    ws = w[rank * vs:(rank + 1) * vs].detach().clone().requires_grad_(True)
    lp, ent = ops.linear_cross_entropy_tp(hs, ws, labels, tau, dist.group.WORLD, rank * vs)
    ((lp * g1).sum() + (ent * g2).sum()).backward()
    rel = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-9)).item()
    errs = dict(lp=(lp - lp0).abs().max().item(), ent=(ent - e0).abs().max().item(), dh=rel(hs.grad, hf.grad),
                dw=rel(ws.grad, wf.grad[rank * vs:(rank + 1) * vs]))
    t = torch.tensor([errs[k] for k in ("lp", "ent", "dh", "dw")], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("CE_TP_RESULTS " + " ".join(f"{k}={v:.3e}" for k, v in zip(("lp", "ent", "dh", "dw"), t.tolist())), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
