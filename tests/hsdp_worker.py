"""HSDP equivalence (torchrun, 4 ranks, gloo/CPU or nccl/GPU): fsdp_size=2 x 2 replicas vs plain ZeRO-3 over 4 ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    cuda = torch.cuda.is_available() and os.environ.get("HSDP_DEVICE", "auto") != "cpu"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if cuda:
        dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    from polyrl_b200.models import get_model_config, pack_padded
    from polyrl_b200.parallel.zero3 import Zero3Engine

    assert world % 2 == 0
    fs = 2
    shard_group = replica_group = None
    for b in range(world // fs):
        g = dist.new_group(list(range(b * fs, (b + 1) * fs)))
        if b == rank // fs:
            shard_group = g
    for i in range(fs):
        g = dist.new_group(list(range(i, world, fs)))
        if i == rank % fs:
            replica_group = g
    cfg = get_model_config("tiny-llama-128" if cuda else "tiny-llama")
    torch.manual_seed(100 + rank)                       # every rank trains on different data
    ids = torch.randint(3, cfg.vocab_size, (2, 17), device=dev)
    inp, _ = pack_padded(ids, torch.ones_like(ids))
    labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), device=dev)
    outs = {}
    for name, kw in (("zero", dict(group=None)), ("hsdp", dict(group=shard_group, replica_group=replica_group)),
                     ("resident", dict(group=None))):
        eng = Zero3Engine(cfg, dev, comm="nccl", lr=1e-2, weight_decay=0.0, seed=7, **kw)
        if name == "resident":              # full replica kept on every rank + activations saved (no recompute)
            eng.make_resident()
            eng.act_ckpt = "none"
        for _ in range(2):
            lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
            (-(lp.mean()) + 0.01 * ent.mean()).backward()
            st = eng.step()
        outs[name] = ({n: t.float().cpu().clone() for n, t in eng.named_full_tensors()}, st["grad_norm"])
    w_err = max(((outs["hsdp"][0][k] - outs["zero"][0][k]).norm() / (outs["zero"][0][k].norm() + 1e-9)).item()
                for k in outs["zero"][0])
    gn_err = abs(outs["hsdp"][1] - outs["zero"][1]) / (abs(outs["zero"][1]) + 1e-9)
    r_err = max(((outs["resident"][0][k] - outs["zero"][0][k]).norm() / (outs["zero"][0][k].norm() + 1e-9)).item()
                for k in outs["zero"][0])
    # LoRA adapters: frozen base gathered once into the resident replica, adapters re-gathered after every step
    from polyrl_b200.models.lora import LoraConfig

    lo = {}
    for name in ("lora_zero", "lora_resident"):
        eng = Zero3Engine(cfg, dev, comm="nccl", lr=1e-2, weight_decay=0.0, seed=7, lora=LoraConfig.from_model_cfg(4, 8, "all-linear"))
        if name == "lora_resident":
            eng.make_resident()
            eng.act_ckpt = "none"
        for _ in range(3):
            lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
            (-(lp.mean()) + 0.01 * ent.mean()).backward()
            eng.step()
        lo[name] = {n: t.float().cpu().clone() for n, t in eng.named_full_tensors() if ".lora." in n}
    l_err = max(((lo["lora_resident"][k] - lo["lora_zero"][k]).norm() / (lo["lora_zero"][k].norm() + 1e-9)).item() for k in lo["lora_zero"])
    if rank == 0:
        print(f"HSDP_RESULTS w_err={w_err:.3e} gn_err={gn_err:.3e} resident_err={r_err:.3e} lora_resident_err={l_err:.3e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
