"""Pipeline-parallel (x DP) training equivalence (torchrun 4 ranks, gloo/CPU): PipelineEngine with the 1F1B schedule must track a
single-process reference (plain autograd over TransformerWeights) — losses, grad norm, updated weights — for pp=2 x dp=2 and for
pp=4 with tied embeddings; at most pp - stage micro-batches are in flight on a stage."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cpu")
    dist.init_process_group("gloo")
    from polyrl_b200.models import TransformerWeights, get_model_config, pack_padded
    from polyrl_b200.models import transformer as tf
    from polyrl_b200.ops import reference as ref
    from polyrl_b200.parallel.pipeline import PipelineEngine, stage_layers

    assert stage_layers(7, 3) == [(0, 3), (3, 5), (5, 7)]
    worst = {}
    from polyrl_b200.parallel.pipeline import InterleavedPipelineEngine, build_interleaved_schedule

    # interleaved schedule: Megatron's tick count when the micro-batch count is a multiple of pp, and the warm-up activation cap
    for pp_, v_, m_ in ((2, 2, 4), (4, 2, 8), (4, 4, 16)):
        tab = build_interleaved_schedule(pp_, v_, m_)
        assert len(tab) == 2 * v_ * m_ + 2 * (pp_ - 1), (pp_, v_, m_, len(tab))
    for name, pp, tied, n_layers, n_micro, virt in (("pp2xdp2", 2, False, 4, 4, 1), ("pp4_tied", 4, True, 5, 6, 1),
                                                    ("pp4_few_micro", 4, False, 4, 2, 1), ("pp2v2xdp2", 2, False, 5, 4, 2),
                                                    ("pp4v2_tied", 4, True, 8, 6, 2)):
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
        cfg = get_model_config(os.environ.get("PP_TRAIN_MODEL", "tiny-llama"), num_layers=n_layers, tie_word_embeddings=tied)
        if virt > 1:
            eng = InterleavedPipelineEngine(cfg, dev, pp_group, dp_group if dp > 1 else None, virtual_stages=virt, lr=1e-2,
                                            weight_decay=0.0, seed=11)
            assert len(eng.chunks) == virt and all(hi > lo for lo, hi in eng.chunks)
        else:
            eng = PipelineEngine(cfg, dev, pp_group, dp_group if dp > 1 else None, lr=1e-2, weight_decay=0.0, seed=11)
            assert eng.hi - eng.lo >= 1
        assert ("embed" in eng.master) == (eng.first or (eng.last and tied))

        def micro(dp_rank: int, k: int, step: int):
            gen = torch.Generator().manual_seed(1000 * step + 50 * dp_rank + k)
            ids = torch.randint(3, cfg.vocab_size, (2, 9 + 3 * k), generator=gen)
            inp, _ = pack_padded(ids, torch.ones_like(ids))
            labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), generator=gen)
            return inp, labels

        w = TransformerWeights(cfg, device="cpu", dtype=torch.float32, init=True, seed=11)
        params = {n: t.requires_grad_(True) for n, t in w.named_tensors()}
        m = {n: torch.zeros_like(t) for n, t in params.items()}
        v = {n: torch.zeros_like(t) for n, t in params.items()}
        loss_err = gn_err = 0.0
        for step in range(2):
            mbs = [micro(rank // pp, k, step) for k in range(n_micro)]
            losses = eng.train_batch(mbs, lambda lp, ent, k: (-(lp.mean()) + 0.01 * ent.mean()) / n_micro)
            cap = min(pp - eng.stage, n_micro) if virt == 1 else min(2 * (pp - eng.stage - 1) + (virt - 1) * pp + 1, virt * n_micro)
            assert eng.peak_in_flight <= cap, (eng.peak_in_flight, eng.stage, cap)
            st = eng.step()
            for p in params.values():
                p.grad = None
            ref_losses = {}
            for d in range(dp):
                for k in range(n_micro):
                    ri, rl = micro(d, k, step)
                    h = tf.forward_hidden(cfg, w, ri)
                    rlp, rent = ref.linear_cross_entropy(h, tf.lm_head_weight(cfg, w.head, w.embed), rl, 1.0)
                    loss = (-(rlp.mean()) + 0.01 * rent.mean()) / n_micro
                    (loss / dp).backward()
                    ref_losses[(d, k)] = loss.item()
            for p in params.values():                 # selection-only parameters (router score bias) get no gradient
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            gn = torch.sqrt(sum(p.grad.pow(2).sum() for p in params.values())).item()
            coef = min(1.0, 1.0 / (gn + 1e-6))
            with torch.no_grad():
                for n, p in params.items():
                    ref.adamw_step(p, m[n], v[n], p.grad, 1e-2, 0.9, 0.999, 1e-8, 0.0, step + 1, coef)
            if eng.last:
                assert len(losses) == n_micro
                loss_err = max(loss_err, max(abs(losses[k] - ref_losses[(rank // pp, k)]) for k in range(n_micro)))
            else:
                assert losses == []
            gn_err = max(gn_err, abs(st["grad_norm"] - gn))
        got = eng.gather_full_state()
        want = {n: t.detach() for n, t in w.named_tensors()}
        assert set(got) == set(want), (sorted(got), sorted(want))
        w_err = max(((got[n] - want[n]).norm() / (want[n].norm() + 1e-9)).item() for n in want)
        # forward-only pass agrees with the reference log-probs on every stage (broadcast from the last)
        mbs = [micro(rank // pp, k, 7) for k in range(3)]
        fo = eng.forward_only(mbs)
        with torch.no_grad():
            for (ri, rl), (lp, ent) in zip(mbs, fo):
                h = tf.forward_hidden(cfg, w, ri)
                rlp, rent = ref.linear_cross_entropy(h, tf.lm_head_weight(cfg, w.head, w.embed), rl, 1.0)
                assert (lp - rlp).abs().max() < 1e-4 and (ent - rent).abs().max() < 1e-4
        worst[name] = (loss_err, gn_err, w_err)
        assert loss_err < 1e-5 and gn_err < 1e-4 and w_err < 2e-3, (name, loss_err, gn_err, w_err)
    # ---- BaseEngine adapter: PPO loss through the pipeline == the ZeRO engine on one rank
    import tempfile

    from polyrl_b200.workers.engine_api import get_engine_cls, ppo_loss

    pp_group = None
    for d in range(world // 2):
        g = dist.new_group([2 * d, 2 * d + 1])
        if d == rank // 2:
            pp_group = g
    solo = [dist.new_group([r]) for r in range(world)][rank]
    gen = torch.Generator().manual_seed(77)                      # same batch on every rank (dp = 1 here: two independent pipelines)
    B, S, R = 6, 20, 8
    ids = torch.randint(3, 500, (B, S), generator=gen)
    mask = torch.ones(B, S, dtype=torch.long)
    mask[0, :4] = 0
    batch = {"input_ids": ids, "attention_mask": mask, "responses": ids[:, -R:], "response_mask": mask[:, -R:],
             "advantages": torch.randn(B, R, generator=gen), "position_ids": (mask.cumsum(-1) - 1).clamp_min(0)}
    ppe = get_engine_cls("language_model", "pp")("tiny-llama", "cpu", pp_group, None, lr=1e-2, micro_batch_size=2, weight_decay=0.0, seed=5)
    ze = get_engine_cls("language_model", "zero")("tiny-llama", "cpu", group=solo, lr=1e-2, micro_batch_size=2, weight_decay=0.0, seed=5)
    for e in (ppe, ze):
        e.initialize()
    full = {n: t.clone() for n, t in ze.engine.named_full_tensors()}          # the engines seed differently: share the ZeRO weights
    cfg_t = ppe.model_cfg
    ppe.engine.load_full({"embed": full["embed"]}, [{k.split(".", 2)[2]: v for k, v in full.items() if k.startswith(f"layers.{i}.")}
                                                    for i in range(cfg_t.num_layers)],
                         {k: full[k] for k in ("norm_w", "lm_head") if k in full})
    a, b = ppe.infer_batch(batch), ze.infer_batch(batch)
    assert (a["log_probs"] - b["log_probs"]).abs().max() < 1e-4 and a["log_probs"].shape == (B, R)
    batch["old_log_probs"] = b["log_probs"] + 0.05 * torch.randn(B, R, generator=gen)
    with ppe.train_mode(), ze.train_mode():
        ma, mb_ = ppe.train_batch(batch, ppo_loss()), ze.train_batch(batch, ppo_loss())
        sa, sb = ppe.optimizer_step(), ze.optimizer_step()
    assert abs(ma["actor/pg_loss"] - mb_["actor/pg_loss"]) < 1e-5 and abs(sa["grad_norm"] - sb["grad_norm"]) < 1e-4
    full_a, full_b = ppe.engine.gather_full_state(), dict(ze.engine.named_full_tensors())
    assert max(((full_a[n] - full_b[n]).norm() / (full_b[n].norm() + 1e-9)).item() for n in full_b) < 2e-3
    # role worker on top of either backend: identical PPO update (two mini-batches -> two optimizer steps)
    from polyrl_b200.protocol import DataProto
    from polyrl_b200.workers.roles import ActorWorker

    dp_ = DataProto.from_dict({k: v for k, v in batch.items()})
    dp_.meta_info["temperature"] = 1.0
    ra, rb = ActorWorker(ppe, ppo_mini_batch_size=3), ActorWorker(ze, ppo_mini_batch_size=3)
    la, lb = ra.compute_log_prob(dp_).batch["old_log_probs"], rb.compute_log_prob(dp_).batch["old_log_probs"]
    assert (la - lb).abs().max() < 1e-4
    ma2, mb2 = ra.update_actor(dp_), rb.update_actor(dp_)
    assert abs(ma2["actor/pg_loss"] - mb2["actor/pg_loss"]) < 1e-4 and abs(ma2["actor/grad_norm"] - mb2["actor/grad_norm"]) < 1e-3
    assert ppe.engine.opt_step == ze.engine.opt_step == 3
    d = tempfile.gettempdir() + f"/pp_ckpt_{os.environ.get('MASTER_PORT', '0')}_{rank // 2}"
    ppe.save_checkpoint(d, global_step=3)
    other = get_engine_cls("language_model", "pp")("tiny-llama", "cpu", pp_group, None, lr=1e-2, micro_batch_size=2, weight_decay=0.0, seed=99)
    other.initialize()
    assert other.load_checkpoint(d)["global_step"] == 3
    assert all(torch.equal(other.engine.master[k], ppe.engine.master[k]) for k in ppe.engine.master)
    t = torch.tensor([max(x[i] for x in worst.values()) for i in range(3)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"PP_TRAIN_OK loss_err={t[0]:.2e} gn_err={t[1]:.2e} w_err={t[2]:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
