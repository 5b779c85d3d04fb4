"""TP (x DP) training equivalence (torchrun 2 or 4 ranks; gloo/CPU or nccl/GPU): a (dp x tp=2) TPTrainEngine must track a
single-process reference (plain autograd over TransformerWeights) step by step — log-probs, grad norm and updated weights."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    cuda = torch.cuda.is_available() and os.environ.get("TP_TRAIN_DEVICE", "auto") != "cpu"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if cuda:
        dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    from polyrl_b200 import ops
    from polyrl_b200.models import TransformerWeights, get_model_config, pack_padded
    from polyrl_b200.models import transformer as tf
    from polyrl_b200.ops import reference as ref
    from polyrl_b200.parallel.tp_train import TPTrainEngine

    tp = 2
    dp = world // tp
    tp_group = dp_group = None
    for d in range(dp):
        g = dist.new_group(list(range(d * tp, (d + 1) * tp)))
        if d == rank // tp:
            tp_group = g
    for t in range(tp):
        g = dist.new_group(list(range(t, world, tp)))
        if t == rank % tp:
            dp_group = g
    cfg = get_model_config(os.environ.get("TP_TRAIN_MODEL") or ("tiny-llama-128" if cuda else "tiny-llama"))
    dtype = torch.bfloat16 if cuda else torch.float32
    sp_mode = os.environ.get("TP_TRAIN_SP", "0") == "1"                 # Megatron sequence parallelism inside the TP group
    eng = TPTrainEngine(cfg, dev, tp_group, dp_group if dp > 1 else None, lr=1e-2, weight_decay=0.0, seed=11, dtype=dtype,
                        sequence_parallel=sp_mode)

    def batch(dp_rank: int):
        gen = torch.Generator().manual_seed(500 + dp_rank)
        ids = torch.randint(3, cfg.vocab_size, (3 if sp_mode else 2, 19), generator=gen).to(dev)      # 57 tokens: odd, exercises the SP tail padding
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), generator=gen).to(dev)
        return inp, labels

    # ---- single-process reference on every rank: plain autograd over the full weights, mean gradient over the dp batches
    w_cpu = TransformerWeights(cfg, device="cpu", dtype=torch.float32, init=True, seed=11)     # same CPU generator as the engine
    w = TransformerWeights(cfg, device=dev, dtype=torch.float32, init=False)
    w.load_state_dict({n: t.to(dev) for n, t in w_cpu.state_dict().items()})
    params = {n: t.requires_grad_(True) for n, t in w.named_tensors()}
    m = {n: torch.zeros_like(t) for n, t in params.items()}
    v = {n: torch.zeros_like(t) for n, t in params.items()}
    lp_err = gn_err = 0.0
    for step in range(2):
        inp, labels = batch(rank // tp)
        lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
        (-(lp.mean()) + 0.01 * ent.mean()).backward()
        st = eng.step()
        # reference
        for p in params.values():
            p.grad = None
        ref_lps = []
        for d in range(dp):
            ri, rl = batch(d)
            h = tf.forward_hidden(cfg, w, ri)            # includes the final norm
            rlp, rent = ref.linear_cross_entropy(h, tf.lm_head_weight(cfg, w.head, w.embed), rl, 1.0)
            ((-(rlp.mean()) + 0.01 * rent.mean()) / dp).backward()
            ref_lps.append(rlp.detach())
        gn = torch.sqrt(sum(p.grad.pow(2).sum() for p in params.values())).item()
        coef = min(1.0, 1.0 / (gn + 1e-6))
        with torch.no_grad():
            for n, p in params.items():
                ref.adamw_step(p, m[n], v[n], p.grad, 1e-2, 0.9, 0.999, 1e-8, 0.0, step + 1, coef)
        if os.environ.get("TP_TRAIN_DEBUG") and rank == 0:
            print("step", step, "lp_err", (lp.detach().float() - ref_lps[rank // tp]).abs().max().item(), "gn", st["grad_norm"], gn, flush=True)
        lp_err = max(lp_err, (lp.detach().float() - ref_lps[rank // tp]).abs().max().item())
        gn_err = max(gn_err, abs(st["grad_norm"] - gn) / gn)
    full = eng.gather_full_state()
    w_err = max(((full[n].cpu() - params[n].detach().cpu()).norm() / (params[n].detach().cpu().norm() + 1e-9)).item() for n in full)
    t = torch.tensor([lp_err, gn_err, w_err], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"TP_TRAIN_RESULTS dp={dp} tp={tp} lp_err={t[0].item():.3e} gn_err={t[1].item():.3e} w_err={t[2].item():.3e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    if os.environ.get("TP_TRAIN_DEVICE", "auto") == "cpu" or not torch.cuda.is_available():
        # CPU SDPA picks different kernels with / without autograd and per head count; pin the math path so the single-process
        # oracle and the TP engine share numerics (Adam turns 1e-4 logit noise into O(lr) weight differences)
        from torch.nn.attention import SDPBackend, sdpa_kernel

        with sdpa_kernel(SDPBackend.MATH):
            main()
    else:
        main()
