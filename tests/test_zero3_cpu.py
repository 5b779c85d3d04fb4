"""Direct tests of the ZeRO engine on CPU (world = 1): activation-checkpoint modes agree, state dict round trip, gradient
accumulation across micro-batches == one big batch, skipped step on non-finite gradients, parameter count."""
import math

import torch

from polyrl_b200.models import get_model_config, pack_padded
from polyrl_b200.parallel.zero3 import Zero3Engine


def _data(cfg, seed, rows=4, length=21):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, cfg.vocab_size, (rows, length), generator=g)
    inp, _ = pack_padded(ids, torch.ones_like(ids))
    labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), generator=g)
    return ids, inp, labels


def _loss(eng, inp, labels, scale=1.0):
    lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
    return (-(lp.mean()) + 0.01 * ent.mean()) * scale


def test_activation_checkpoint_modes_are_equivalent():
    cfg = get_model_config("tiny-llama")
    _, inp, labels = _data(cfg, 1)
    outs = {}
    for mode in ("full", "none", "1", "auto"):
        eng = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=4)
        eng.act_ckpt = mode
        for _ in range(2):
            _loss(eng, inp, labels).backward()
            st = eng.step()
        outs[mode] = ({n: t.clone() for n, t in eng.named_full_tensors()}, st["grad_norm"], eng.saved_layers_last)
    assert outs["full"][2] == 0 and outs["none"][2] == cfg.num_layers and outs["1"][2] == 1 and outs["auto"][2] == cfg.num_layers
    for mode in ("none", "1", "auto"):
        assert abs(outs[mode][1] - outs["full"][1]) < 1e-5
        assert all(torch.allclose(outs[mode][0][n], outs["full"][0][n], atol=1e-6) for n in outs["full"][0])


def test_micro_batch_accumulation_equals_one_batch_and_state_dict_roundtrip():
    cfg = get_model_config("tiny-llama")
    ids, inp, labels = _data(cfg, 2)
    a = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.01, seed=5)
    _loss(a, inp, labels).backward()
    sa = a.step()
    b = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.01, seed=5)
    # the same 4 sequences as two micro-batches of 2; the token-mean loss needs the per-micro-batch weight 1/2
    n_half = inp.input_ids.numel() // 2
    for part in (0, 1):
        i2, _ = pack_padded(ids[2 * part:2 * part + 2], torch.ones(2, ids.shape[1], dtype=torch.long))
        _loss(b, i2, labels[part * n_half:(part + 1) * n_half], scale=0.5).backward()
    sb = b.step()
    assert abs(sa["grad_norm"] - sb["grad_norm"]) < 1e-4
    fa, fb = dict(a.named_full_tensors()), dict(b.named_full_tensors())
    # AdamW's first step is sign-like: elements whose gradient is ~0 may flip under fp32 re-association, so compare in norm
    assert max(((fa[n] - fb[n]).norm() / (fa[n].norm() + 1e-9)).item() for n in fa) < 2e-3
    sd = a.state_dict()
    c = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.01, seed=99)
    c.load_state_dict(sd)
    assert c.opt_step == a.opt_step
    _loss(a, inp, labels).backward(); _loss(c, inp, labels).backward()
    a.step(); c.step()
    fa, fc = dict(a.named_full_tensors()), dict(c.named_full_tensors())
    assert all(torch.equal(fa[n], fc[n]) for n in fa)                 # moments + master restored bit for bit


def test_non_finite_gradients_skip_the_step_and_counts():
    cfg = get_model_config("tiny-llama")
    _, inp, labels = _data(cfg, 3)
    eng = Zero3Engine(cfg, "cpu", lr=1e-2, seed=6)
    before = {n: t.clone() for n, t in eng.named_full_tensors()}
    (_loss(eng, inp, labels) * float("nan")).backward()
    st = eng.step()
    assert st["skipped"] and not math.isfinite(st["grad_norm"]) and eng.opt_step == 0
    after = dict(eng.named_full_tensors())
    assert all(torch.equal(before[n], after[n]) for n in before)
    _loss(eng, inp, labels).backward()
    assert not eng.step()["skipped"] and eng.opt_step == 1           # gradients of the skipped step were dropped
    h, f, v, L = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.num_layers
    qkv = (cfg.num_heads + 2 * cfg.num_kv_heads) * cfg.head_dim
    assert eng.num_params() == L * (h * qkv + cfg.num_heads * cfg.head_dim * h + 3 * h * f + 2 * h) + 2 * v * h + h
    eng.arm_norm_fold()                                               # CPU / library path: a no-op that must not break step()
    _loss(eng, inp, labels).backward()
    assert eng.step()["grad_norm"] > 0


def test_activation_offload_hooks_keep_gradients_exact():
    """Saved-activation mode with host offload: activations above the size threshold are copied out at forward and restored at
    backward; parameter views stay put; the update is bit-identical to the run without offload."""
    cfg = get_model_config("tiny-llama")
    _, inp, labels = _data(cfg, 7)
    res = {}
    for off in (False, True):
        eng = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=8)
        eng.act_ckpt = "none"
        eng.act_offload, eng._offload_force, eng.act_offload_min_bytes = off, True, 1024
        _loss(eng, inp, labels).backward()
        st = eng.step()
        res[off] = ({n: t.clone() for n, t in eng.named_full_tensors()}, st["grad_norm"], dict(eng.offload_stats))
    assert res[False][2] == {"bytes": 0, "tensors": 0}
    assert res[True][2]["tensors"] >= cfg.num_layers * 4 and res[True][2]["bytes"] > 0
    assert res[True][1] == res[False][1]
    assert all(torch.equal(res[True][0][n], res[False][0][n]) for n in res[False][0])
    # weights never go through the hooks: the offloaded volume scales with the token count only
    _, inp2, labels2 = _data(cfg, 7, rows=8)
    eng = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=8)
    eng.act_ckpt = "none"
    eng.act_offload, eng._offload_force, eng.act_offload_min_bytes = True, True, 1024
    _loss(eng, inp2, labels2).backward()
    assert 1.8 < eng.offload_stats["bytes"] / res[True][2]["bytes"] < 2.2


def test_reference_policy_follows_the_actor_on_request():
    """trainer.ref_update_freq: ``update_reference_policy`` copies the trained actor into the frozen reference engine."""
    from polyrl_b200.config.schema import load_config
    from polyrl_b200.workers.hybrid_worker import ActorRolloutRefWorker

    cfg = load_config(overrides=["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-llama", "actor_rollout_ref.actor.use_kl_loss=True",
                                 "actor_rollout_ref.actor.use_dynamic_bsz=True", "actor_rollout_ref.rollout.cuda_graph=False",
                                 "actor_rollout_ref.rollout.max_num_seqs=8", "data.max_prompt_length=16", "data.max_response_length=8"])
    w = ActorRolloutRefWorker(cfg, role="actor_rollout_ref", device="cpu")
    w.init_model()
    with torch.no_grad():
        for u in w.actor_engine.base_units:
            u.shard.add_(0.01)
    a = {n: t.clone() for n, t in w.actor_engine.named_full_tensors()}
    r0 = {n: t.clone() for n, t in w.ref_engine.named_full_tensors()}
    assert any(not torch.equal(a[n], r0[n]) for n in a)
    w.update_reference_policy()
    r1 = dict(w.ref_engine.named_full_tensors())
    assert all(torch.equal(a[n].to(r1[n].dtype), r1[n]) for n in a)


def test_async_checkpoint_snapshot_is_isolated_from_later_updates(tmp_path):
    """checkpoint.async_save: the state is copied to host memory at save() time and written in the background — weights changed
    right after save() do not leak into the file; load() joins the writer first."""
    from polyrl_b200.models import get_model_config
    from polyrl_b200.parallel.zero3 import Zero3Engine
    from polyrl_b200.utils.checkpoint import ShardedCheckpointManager

    cfg = get_model_config("tiny-llama")
    eng = Zero3Engine(cfg, "cpu", seed=4)
    mgr = ShardedCheckpointManager(eng, cfg, async_save=True)
    want = {n: t.clone() for n, t in eng.named_full_tensors()}
    mgr.save(str(tmp_path / "global_step_1" / "actor"), 1, None, extra={"lr_step": 7})
    with torch.no_grad():
        for u in eng.units:
            u.shard.add_(1.0)
            if u.master is not None:
                u.master.add_(1.0)
    extra = mgr.load(str(tmp_path / "global_step_1" / "actor"))
    got = dict(eng.named_full_tensors())
    assert extra["lr_step"] == 7 and all(torch.equal(got[n], want[n]) for n in want)
    mgr.wait()


def test_sandwich_norm_family_trains_through_the_sharded_engine():
    """GLM-4 blocks (RMSNorm on the attention and MLP outputs, partial rotary): the engine's loss and every gradient-driven update
    equal plain autograd over the functional model with the same weights, in every activation-checkpoint mode."""
    from polyrl_b200.models import transformer as tf
    from polyrl_b200.ops import reference as ref

    cfg = get_model_config("tiny-glm4")
    _, inp, labels = _data(cfg, 3)
    eng = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=6)
    names = [n for n, _ in eng.named_full_tensors()]
    assert any(n.endswith("ln_attn") for n in names) and any(n.endswith("ln_mlp") for n in names)
    w = tf.TransformerWeights(cfg, "cpu", torch.float32, init=False)
    w.load_state_dict({n: t.clone() for n, t in eng.named_full_tensors()})
    w.requires_grad_(True)
    h = tf.forward_hidden(cfg, w, inp)
    logits = h @ tf.lm_head_weight(cfg, w.head, w.embed).t()
    lp_ref = torch.log_softmax(logits.float(), -1).gather(1, labels[:, None]).squeeze(1)
    lp, _ = eng.forward_logprob_entropy(inp, labels, 1.0)
    assert torch.allclose(lp, lp_ref, atol=1e-5)
    (-lp_ref.mean()).backward()
    (-lp.mean()).backward()
    gn = math.sqrt(sum(float(t.grad.pow(2).sum()) for _, t in w.named_tensors() if t.grad is not None))
    st = eng.step()
    assert abs(st["grad_norm"] - gn) < 1e-4 * max(1.0, gn)
    for mode in ("none", "1"):
        e2 = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=6)
        e2.act_ckpt = mode
        (-e2.forward_logprob_entropy(inp, labels, 1.0)[0].mean()).backward()
        assert abs(e2.step()["grad_norm"] - gn) < 1e-4 * max(1.0, gn)
