"""Mixture-of-experts family (Qwen3-MoE layout) on CPU: logits equal HF, HF checkpoint round trip, the ZeRO engine trains router
and experts like plain autograd, the rollout engine decodes it, the pipeline engine's stage function runs it."""
import math

import pytest
import torch

from polyrl_b200.models import hf_io
from polyrl_b200.models import transformer as tf
from polyrl_b200.models.config import ModelConfig, get_model_config
from polyrl_b200.models.transformer import TransformerWeights, pack_padded


def _hf_tiny():
    from transformers import Qwen3MoeConfig, Qwen3MoeForCausalLM

    torch.manual_seed(0)
    c = Qwen3MoeConfig(vocab_size=512, hidden_size=128, intermediate_size=256, moe_intermediate_size=128, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=2, head_dim=32, num_experts=8, num_experts_per_tok=2, norm_topk_prob=True,
                       max_position_embeddings=256, decoder_sparse_step=1, mlp_only_layers=[], tie_word_embeddings=False,
                       pad_token_id=0, eos_token_id=1, bos_token_id=2)
    c._attn_implementation = "eager"
    m = Qwen3MoeForCausalLM(c).eval().float()
    with torch.no_grad():
        for n, p in m.named_parameters():          # default init leaves the router flat and the stacked experts empty-ish
            if "experts" in n or "mlp.gate" in n:
                p.normal_(0.0, 0.05)
            elif p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return c, m


def test_moe_logits_match_hf_and_checkpoint_round_trip(tmp_path):
    from transformers import Qwen3MoeForCausalLM

    hc, model = _hf_tiny()
    cfg = ModelConfig.from_hf_config(hc.to_dict())
    assert cfg.num_experts == 8 and cfg.num_experts_per_tok == 2 and cfg.moe_intermediate_size == 128 and cfg.qk_norm and cfg.norm_topk_prob
    w = tf.TransformerWeights(cfg, init=False)
    w.load_state_dict(tf.convert_hf_state_dict(cfg, model.state_dict()))
    assert w.layers[0]["ewgu"].shape == (8, 256, 128) and w.layers[0]["router"].shape == (8, 128)
    ids = torch.randint(3, 500, (2, 23), generator=torch.Generator().manual_seed(1))
    inp, _ = tf.pack_padded(ids, torch.ones_like(ids))
    with torch.no_grad():
        ref_logits = model(ids).logits.reshape(-1, 512)
        mine = tf.logits_forward(cfg, w.head, w.embed, tf.forward_hidden(cfg, w, inp))
    assert (mine - ref_logits).abs().max() < 5e-4
    # several experts actually fire and tokens disagree on their choice
    h = torch.randn(46, 128)
    top = torch.topk(torch.softmax(h @ w.layers[0]["router"].t(), -1), 2).indices
    assert top.unique().numel() >= 4
    # export -> HF reloads it; load back into a fresh engine
    hf_io.save_hf(w.state_dict(), cfg, tmp_path / "ck", dtype=torch.float32)
    m2 = Qwen3MoeForCausalLM.from_pretrained(tmp_path / "ck", dtype=torch.float32).eval()
    with torch.no_grad():
        assert (m2(ids).logits.reshape(-1, 512) - ref_logits).abs().max() < 1e-4
    cfg2 = get_model_config(str(tmp_path / "ck"))
    assert cfg2.num_experts == 8 and cfg2.moe_intermediate_size == 128
    sd = hf_io.load_state_dict(cfg2, tmp_path / "ck")
    assert all(torch.equal(sd[k], v) for k, v in w.state_dict().items())
    dense = get_model_config("tiny-qwen3-moe", num_experts=0, num_experts_per_tok=0, moe_intermediate_size=0)
    moe = get_model_config("tiny-qwen3-moe")
    assert moe.num_params() == sum(t.numel() for _, t in tf.TransformerWeights(moe).named_tensors())
    assert moe.num_active_params() == moe.num_params() - 2 * 6 * 3 * 128 * 128 and dense.num_active_params() == dense.num_params()


def test_zero_engine_trains_router_and_experts_like_autograd():
    from polyrl_b200.ops import reference as ref
    from polyrl_b200.parallel.zero3 import Zero3Engine

    cfg = get_model_config("tiny-qwen3-moe")
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(3, 500, (3, 17), generator=g)
    inp, _ = tf.pack_padded(ids, torch.ones_like(ids))
    labels = torch.randint(0, 500, (inp.input_ids.numel(),), generator=g)
    for mode in ("full", "none"):
        eng = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=3, max_grad_norm=1e9)
        eng.act_ckpt = mode
        full = {n: t.clone() for n, t in eng.named_full_tensors()}
        w = tf.TransformerWeights(cfg, init=False)
        w.load_state_dict(full)
        w.requires_grad_(True)
        lp_o, _ = ref.linear_cross_entropy(tf.forward_hidden(cfg, w, inp), w.head["lm_head"], labels)
        (-lp_o.mean()).backward()
        lp, _ = eng.forward_logprob_entropy(inp, labels, 1.0)
        assert (lp - lp_o.detach()).abs().max() < 1e-5
        (-lp.mean()).backward()
        for u in eng.units:
            u.wait_grads()
            for (n, shp), numel, so in zip(u.specs, u.numels, u.shard_offsets):
                name = n if u in (eng.embed, eng.head) else f"{u.name}.{n}"
                want = dict(w.named_tensors())[name].grad
                got = u.grad_acc[so:so + numel].view(shp)
                assert (got - want).abs().max() < 1e-5 * max(1.0, want.abs().max().item()), (mode, name)
        gr = dict(w.named_tensors())
        assert gr["layers.0.router"].grad.abs().max() > 0 and gr["layers.1.ewgu"].grad.abs().max() > 0
        assert not eng.step()["skipped"]


def test_rollout_engine_decodes_moe_like_full_forward():
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams

    cfg = get_model_config("tiny-qwen3-moe")
    w = tf.TransformerWeights(cfg, seed=5)
    for lw in w.layers:
        lw["router"].normal_(0.0, 0.5, generator=torch.Generator().manual_seed(1))      # decisive routing
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=8, max_model_len=96, kv_cache_tokens=8 * 96, max_prefill_tokens=16)
    eng.set_weights(w.embed, w.layers, w.head)
    prompts = [[5, 6, 7, 8, 9, 10, 11], [20, 21, 22], list(range(30, 60))]
    out = eng.generate(prompts, SamplingParams(max_new_tokens=10, temperature=0.0, ignore_eos=True))
    for p, grp in zip(prompts, out):
        seq = list(p)
        for tok in grp[0].output_ids:                     # greedy oracle: full forward over the growing sequence
            t = torch.tensor(seq)
            inp = tf.PackedInput(t, torch.arange(len(seq)), torch.tensor([0, len(seq)], dtype=torch.int32), len(seq))
            with torch.no_grad():
                nxt = int(tf.logits_forward(cfg, w.head, w.embed, tf.forward_hidden(cfg, w, inp))[-1].argmax())
            assert nxt == tok
            seq.append(tok)


def test_router_aux_loss_matches_hf_and_trains_through_zero_units():
    from transformers.models.qwen3_moe.modeling_qwen3_moe import load_balancing_loss_func

    from polyrl_b200.ops import reference as ref
    from polyrl_b200.parallel.zero3 import Zero3Engine

    cfg = get_model_config("tiny-qwen3-moe")
    g = torch.Generator().manual_seed(4)
    logits = torch.randn(37, 8, generator=g)
    probs = torch.softmax(logits, -1)
    top_e = torch.topk(probs, 2).indices
    assert abs(tf.moe_load_balance_loss(cfg, probs, top_e).item() - load_balancing_loss_func((logits,), 8, 2).item()) < 1e-5
    ids = torch.randint(3, 500, (3, 15), generator=g)
    inp, _ = tf.pack_padded(ids, torch.ones_like(ids))
    labels = torch.randint(0, 500, (inp.input_ids.numel(),), generator=g)
    for mode in ("full", "none"):
        eng = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=3, max_grad_norm=1e9)
        eng.act_ckpt, eng.moe_aux = mode, True
        full = {n: t.clone() for n, t in eng.named_full_tensors()}
        w = tf.TransformerWeights(cfg, init=False)
        w.load_state_dict(full)
        w.requires_grad_(True)
        aux: list = []
        x = tf.embed_forward(cfg, w.embed, inp)
        for lw in w.layers:
            x = tf.layer_forward(cfg, lw, x, inp, aux_out=aux)
        lp_o, _ = ref.linear_cross_entropy(tf.final_norm(cfg, w.head, x), w.head["lm_head"], labels)
        aux_o = torch.stack(aux).mean()
        (-lp_o.mean() + 0.5 * aux_o).backward()
        lp, _ = eng.forward_logprob_entropy(inp, labels, 1.0)
        assert abs(eng.last_aux_loss.item() - aux_o.item()) < 1e-6 and eng.last_aux_loss.item() > 0.9      # >= 1 by Cauchy-Schwarz-ish
        (-lp.mean() + 0.5 * eng.last_aux_loss).backward()
        grads = dict(w.named_tensors())
        for u in eng.units:
            u.wait_grads()
            for (n, shp), numel, so in zip(u.specs, u.numels, u.shard_offsets):
                name = n if u in (eng.embed, eng.head) else f"{u.name}.{n}"
                want = grads[name].grad
                assert (u.grad_acc[so:so + numel].view(shp) - want).abs().max() < 1e-5 * max(1.0, want.abs().max().item()), (mode, name)
    # the PPO actor adds the term when the coefficient is set
    from polyrl_b200.config.schema import load_config
    from polyrl_b200.workers.hybrid_worker import ActorRolloutRefWorker

    c = load_config(None, ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-qwen3-moe", "data.max_prompt_length=12",
                           "data.max_response_length=6", "actor_rollout_ref.rollout.max_num_seqs=8", "actor_rollout_ref.rollout.n=2",
                           "actor_rollout_ref.rollout.cuda_graph=False", "actor_rollout_ref.actor.ppo_mini_batch_size=4",
                           "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=2", "actor_rollout_ref.actor.router_aux_loss_coef=0.01"])
    wk = ActorRolloutRefWorker(c, role="actor_rollout", device="cpu")
    wk.init_model()
    assert wk.actor_engine.moe_aux
    import numpy as np

    from polyrl_b200.protocol import DataProto

    pids = torch.randint(3, 200, (2, 12), generator=g)
    dp = DataProto.from_dict({"input_ids": pids, "attention_mask": torch.ones_like(pids), "position_ids": torch.arange(12).expand(2, -1).contiguous()},
                             non_tensors={"uid": np.array(["a", "b"], dtype=object)})
    dp.meta_info.update(temperature=1.0, ignore_eos=True)
    out = wk.generate_sequences(dp)
    out = out.union(wk.compute_log_prob(out))
    out.batch["advantages"] = torch.randn(4, 6, generator=g) * out.batch["response_mask"]
    out.meta_info.update({"is_opt_step": True})
    m = wk.update_actor_stream(out).meta_info["metrics"]
    assert "actor/router_aux_loss" in m and min(m["actor/router_aux_loss"]) > 0.9


def test_static_shape_expert_path_matches_the_grouped_loop(monkeypatch):
    """Graph-capturable MoE decode: the fixed-capacity batched-GEMM path equals the per-expert loop (also when every token picks the
    same experts), and the engine produces the same greedy tokens with it."""
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams

    cfg = get_model_config("tiny-qwen3-moe")
    w = tf.TransformerWeights(cfg, seed=5)
    g = torch.Generator().manual_seed(0)
    for lw in w.layers:
        lw["router"].normal_(0.0, 0.5, generator=g)
    h = torch.randn(37, cfg.hidden_size, generator=g)
    a, b = tf.moe_mlp(cfg, w.layers[0], h), tf.moe_mlp_static(cfg, w.layers[0], h)
    assert (a - b).abs().max() < 1e-5
    same = h[:1].expand(9, -1).contiguous()                       # all rows route identically: one expert receives all 9 tokens
    assert (tf.moe_mlp(cfg, w.layers[0], same) - tf.moe_mlp_static(cfg, w.layers[0], same)).abs().max() < 1e-5
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("POLYRL_MOE_STATIC_DECODE", flag)
        eng = RolloutEngine(cfg, "cpu", max_num_seqs=8, max_model_len=64, kv_cache_tokens=8 * 64)
        eng.set_weights(w.embed, w.layers, w.head)
        outs[flag] = [q[0].output_ids for q in eng.generate([[5, 6, 7, 8], [20, 21, 22]], SamplingParams(max_new_tokens=8, temperature=0.0, ignore_eos=True))]
    assert outs["0"] == outs["1"]


def test_tp_expert_sharding_partials_sum_to_the_full_moe():
    """rollout/tp.py::moe_partial: every TP rank computes the assignments routed to ITS experts; the partial sums (dynamic grouping
    and the static-shape graph-capturable variant with its spill row) add up to the single-device mixture."""
    from types import SimpleNamespace

    from polyrl_b200.models import TransformerWeights, get_model_config
    from polyrl_b200.models import transformer as tf
    from polyrl_b200.rollout.tp import moe_partial

    from polyrl_b200.rollout.tp import TPContext

    for preset, li in (("tiny-qwen3-moe", 0), ("tiny-glm4-moe", 1), ("tiny-qwen2-moe", 0)):
        cfg = get_model_config(preset)
        w = TransformerWeights(cfg, "cpu", torch.float32, seed=2).layers[li]
        if "router_bias" in w:
            w["router_bias"].normal_(0.0, 0.2, generator=torch.Generator().manual_seed(4))
        h = torch.randn(13, cfg.hidden_size)
        want = tf.moe_mlp(cfg, w, h)
        for world in (2, 4):              # 4 ranks: the 256-wide shared experts no longer split into whole 128-row blocks -> rank 0 runs them
            for static in (False, True):
                total = torch.zeros_like(want)
                for r in range(world):
                    tp = SimpleNamespace(rank=r, world=world, cfg=cfg)
                    wl = TPContext.shard_layer(tp, w)                 # the engine's own slicing (experts whole, shared expert column / row)
                    wl = {k: v for k, v in wl.items() if k in ("router", "router_bias", "ewgu", "ewd", "swgu", "swd", "sgate")}
                    total += moe_partial(tp, cfg, wl, h, static)
                assert (total - want).abs().max() < 1e-5, (preset, world, static)


def test_mixtral_logits_match_hf_and_legacy_checkpoint_names_load(tmp_path):
    """Mixtral family: llama attention + top-2 of N SwiGLU experts.  Logits equal HF MixtralForCausalLM; a checkpoint directory in
    the on-disk naming of the released weights (block_sparse_moe.experts.{e}.w1/w2/w3) loads through the lazy reader."""
    import transformers
    from safetensors.torch import save_file

    from polyrl_b200.models import TransformerWeights, forward_hidden, lm_head_weight, pack_padded
    from polyrl_b200.models import transformer as tf
    from polyrl_b200.models.config import ModelConfig
    from polyrl_b200.models.hf_io import hf_config_dict, load_state_dict

    torch.manual_seed(0)
    hc = transformers.MixtralConfig(vocab_size=320, hidden_size=128, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                    num_key_value_heads=2, max_position_embeddings=256, num_local_experts=4, num_experts_per_tok=2,
                                    sliding_window=None, rms_norm_eps=1e-5, tie_word_embeddings=False)
    m = transformers.MixtralForCausalLM(hc).eval().float()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "experts" in n:
                p.normal_(0, 0.05)
            elif p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    cfg = ModelConfig.from_hf_config(hc.to_dict())
    assert cfg.num_experts == 4 and cfg.moe_intermediate_size == 128 and cfg.norm_topk_prob and not cfg.qk_norm
    sd = tf.convert_hf_state_dict(cfg, m.state_dict())
    w = TransformerWeights(cfg, "cpu", torch.float32, init=False)
    w.load_state_dict(sd)
    ids = torch.randint(0, 320, (2, 19))
    with torch.no_grad():
        want = m(ids).logits
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        got = forward_hidden(cfg, w, inp) @ lm_head_weight(cfg, w.head, w.embed).t()
    assert (got - want.reshape(-1, 320)).abs().max() < 3e-4
    assert hf_config_dict(cfg)["model_type"] == "mixtral" and hf_config_dict(cfg)["num_local_experts"] == 4
    # released-checkpoint naming on disk
    legacy = {}
    for k, v in m.state_dict().items():
        if ".mlp.experts.gate_up_proj" in k:
            pre = k.replace("mlp.experts.gate_up_proj", "block_sparse_moe.experts.")
            for e in range(4):
                legacy[f"{pre}{e}.w1.weight"], legacy[f"{pre}{e}.w3.weight"] = v[e, :128].clone(), v[e, 128:].clone()
        elif ".mlp.experts.down_proj" in k:
            pre = k.replace("mlp.experts.down_proj", "block_sparse_moe.experts.")
            for e in range(4):
                legacy[f"{pre}{e}.w2.weight"] = v[e].clone()
        elif ".mlp.gate.weight" in k:
            legacy[k.replace("mlp.gate.weight", "block_sparse_moe.gate.weight")] = v.clone()
        else:
            legacy[k] = v.clone()
    d = tmp_path / "mixtral"
    d.mkdir()
    save_file(legacy, str(d / "model.safetensors"))
    (d / "config.json").write_text(__import__("json").dumps(hc.to_dict()))
    back = load_state_dict(cfg, d)
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def _randomise_(model, skip=()):
    """Move HF's all-ones / zero / uninitialised tensors to informative values (stacked experts come back uninitialised)."""
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p_ in model.named_parameters():
            if "experts.gate_up_proj" in n or "experts.down_proj" in n:
                p_.copy_(0.05 * torch.randn(p_.shape, generator=g))
            elif p_.dim() == 1:
                p_.add_(0.1 * torch.randn(p_.shape, generator=g))
        for n, b in model.named_buffers():
            if "e_score_correction_bias" in n:
                b.copy_(0.2 * torch.randn(b.shape, generator=g))


@pytest.mark.parametrize("family", ["qwen2_moe", "glm4_moe", "qwen3_moe_mixed"])
def test_moe_variants_match_hf(family, tmp_path):
    """Qwen2-MoE (shared expert behind a per-token sigmoid gate, qkv bias, a dense block via mlp_only_layers), GLM-4.5-style MoE
    (sigmoid scores, bias-corrected group-limited selection, scaled weights, always-on shared experts, first block dense, partial
    rotary, per-head QK norm) and a Qwen3-MoE stack with decoder_sparse_step = 2: logits equal HF, greedy decoding through the paged
    engine equals the full forward, ZeRO gradients equal plain autograd, and the HF checkpoint directory round-trips."""
    import transformers

    from polyrl_b200.models import hf_io
    from polyrl_b200.parallel.zero3 import Zero3Engine
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams

    torch.manual_seed(0)
    common = dict(vocab_size=320, hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
                  max_position_embeddings=128, rms_norm_eps=1e-5, tie_word_embeddings=False, pad_token_id=0)
    if family == "qwen2_moe":
        hc = transformers.Qwen2MoeConfig(**common, num_experts=8, num_experts_per_tok=2, moe_intermediate_size=128,
                                         shared_expert_intermediate_size=256, norm_topk_prob=False, mlp_only_layers=[1], decoder_sparse_step=1)
        m = transformers.Qwen2MoeForCausalLM(hc)
    elif family == "glm4_moe":
        hc = transformers.Glm4MoeConfig(**common, head_dim=32, n_routed_experts=8, num_experts_per_tok=2, moe_intermediate_size=128,
                                        n_shared_experts=1, n_group=4, topk_group=2, routed_scaling_factor=1.5, norm_topk_prob=True,
                                        first_k_dense_replace=1, use_qk_norm=True, attention_bias=True)
        m = transformers.Glm4MoeForCausalLM(hc)
    else:
        hc = transformers.Qwen3MoeConfig(**common, head_dim=32, num_experts=8, num_experts_per_tok=2, moe_intermediate_size=128,
                                         norm_topk_prob=True, mlp_only_layers=[], decoder_sparse_step=2)
        m = transformers.Qwen3MoeForCausalLM(hc)
    m = m.eval().float()
    _randomise_(m)
    cfg = ModelConfig.from_hf_config(m.config.to_dict())
    sparse = [tf.is_sparse_layer(cfg, i) for i in range(3)]
    assert sparse == {"qwen2_moe": [True, False, True], "glm4_moe": [False, True, True], "qwen3_moe_mixed": [False, True, False]}[family]
    w = TransformerWeights(cfg, "cpu", torch.float32, init=False)
    sd = m.state_dict()
    w.load_state_dict(tf.convert_hf_state_dict(cfg, sd))
    assert sum(t.numel() for _, t in w.named_tensors()) == cfg.num_params()
    ids = torch.randint(3, 320, (2, 33))
    with torch.no_grad():
        want = m(ids).logits
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        got = tf.forward_hidden(cfg, w, inp) @ tf.lm_head_weight(cfg, w.head, w.embed).t()
    assert (got - want.reshape(-1, 320)).abs().max().item() < 3e-4
    # paged engine (dynamic and static expert paths) vs the full forward
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=4, max_model_len=64, kv_cache_tokens=4 * 64)
    eng.set_weights(w.embed, w.layers, w.head)
    prompt = ids[0, :9].tolist()
    out = eng.generate([prompt], SamplingParams(max_new_tokens=6, temperature=0.0, ignore_eos=True))[0][0].output_ids
    with torch.no_grad():
        hf_out = m.generate(torch.tensor([prompt]), max_new_tokens=6, do_sample=False, eos_token_id=None, pad_token_id=0)[0, 9:].tolist()
    assert out == hf_out
    h = torch.randn(5, cfg.hidden_size)
    lw = w.layers[sparse.index(True)]
    assert (tf.moe_mlp_static(cfg, lw, h) - tf.moe_mlp(cfg, lw, h)).abs().max() < 1e-5
    # ZeRO engine: same log-probs and gradient norm as plain autograd
    labels = torch.randint(0, 320, (inp.input_ids.numel(),))
    e = Zero3Engine(cfg, "cpu", lr=1e-3, weight_decay=0.0, seed=2, max_grad_norm=1e9)
    for u, src in [(e.embed, w.embed), (e.head, w.head)] + list(zip(e.layers, w.layers)):
        u.load_full_state(src)
    w.requires_grad_(True)
    lp_o = torch.log_softmax((tf.forward_hidden(cfg, w, inp) @ tf.lm_head_weight(cfg, w.head, w.embed).t()).float(), -1).gather(1, labels[:, None]).squeeze(1)
    (-lp_o.mean()).backward()
    gn = math.sqrt(sum(float(t.grad.pow(2).sum()) for _, t in w.named_tensors() if t.grad is not None))
    lp, _ = e.forward_logprob_entropy(inp, labels, 1.0)
    assert (lp - lp_o.detach()).abs().max() < 1e-5
    (-lp.mean()).backward()
    assert abs(e.step()["grad_norm"] - gn) < 1e-4 * max(1.0, gn)
    # checkpoint directory out and back, and HF reloads it with the same logits
    w.requires_grad_(False)
    e2 = Zero3Engine(cfg, "cpu", seed=5)
    for u, src in [(e2.embed, w.embed), (e2.head, w.head)] + list(zip(e2.layers, w.layers)):
        u.load_full_state(src)
    hf_io.save_hf(e2, cfg, tmp_path / "ck", dtype=torch.float32)
    m2 = type(m).from_pretrained(tmp_path / "ck", dtype=torch.float32).eval()
    with torch.no_grad():
        assert (m2(ids).logits - want).abs().max() < 1e-4
    cfg2 = get_model_config(str(tmp_path / "ck"))
    assert cfg2.dense_layers == cfg.dense_layers and cfg2.moe_router == cfg.moe_router and cfg2.shared_expert_intermediate_size == cfg.shared_expert_intermediate_size
    back = hf_io.load_state_dict(cfg2, tmp_path / "ck")
    mine = w.state_dict()
    assert set(back) == set(mine) and all(torch.allclose(back[k], mine[k]) for k in mine)


@pytest.mark.parametrize("variant", ["q_lora_interleaved", "plain_q_yarn", "dense", "v2_group_limited", "v2_lite_greedy"])
def test_deepseek_v3_mla_matches_hf(variant, tmp_path):
    """DeepSeek-V3 architecture (multi-head latent attention + sigmoid group-limited routing + shared experts; Moonlight ships under
    the same model_type): logits equal HF ``DeepseekV3ForCausalLM`` with the query bottleneck and interleaved rotary weights, with a
    plain query projection under YaRN (mscale folded into the softmax scale), and as a dense MLA stack; greedy decoding through the
    paged engine — which caches only the latent and decodes in the absorbed form — equals HF ``generate``, also with chunked prefill
    behind a cached prefix; ZeRO gradients equal plain autograd; the HF checkpoint directory round-trips."""
    import transformers

    from polyrl_b200.models import hf_io
    from polyrl_b200.parallel.zero3 import Zero3Engine
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams

    torch.manual_seed(0)
    kw = dict(q_lora_rank=48, rope_interleave=True)
    if variant == "plain_q_yarn":
        kw = dict(q_lora_rank=None, rope_interleave=False,
                  rope_parameters=dict(rope_type="yarn", factor=4.0, original_max_position_embeddings=32, beta_fast=32.0, beta_slow=1.0,
                                       mscale=1.0, mscale_all_dim=0.8, rope_theta=10000.0))
    n_layers = 3
    v2 = variant.startswith("v2")
    if v2:        # DeepSeek-V2 / V2-Lite: the same attention with complex-pair rotary, softmax routing (optionally group-limited), scaled weights
        hc = transformers.DeepseekV2Config(
            vocab_size=320, hidden_size=128, intermediate_size=256, moe_intermediate_size=128, num_hidden_layers=n_layers, num_attention_heads=4,
            num_key_value_heads=4, n_shared_experts=2, n_routed_experts=8, routed_scaling_factor=16.0 if variant == "v2_group_limited" else 1.0,
            kv_lora_rank=32, q_lora_rank=48 if variant == "v2_group_limited" else None, qk_rope_head_dim=16, v_head_dim=32, qk_nope_head_dim=32,
            topk_method="group_limited_greedy" if variant == "v2_group_limited" else "greedy", n_group=4, topk_group=2, num_experts_per_tok=2,
            first_k_dense_replace=1, max_position_embeddings=128, rms_norm_eps=1e-6, pad_token_id=0, tie_word_embeddings=False)
    hc = hc if v2 else transformers.DeepseekV3Config(
        vocab_size=320, hidden_size=128, intermediate_size=256, moe_intermediate_size=128, num_hidden_layers=n_layers, num_attention_heads=4,
        num_key_value_heads=4, n_shared_experts=1, n_routed_experts=8, routed_scaling_factor=2.5, kv_lora_rank=32, qk_rope_head_dim=16,
        v_head_dim=32, qk_nope_head_dim=32, n_group=4, topk_group=2, num_experts_per_tok=2,
        first_k_dense_replace=n_layers if variant == "dense" else 1, norm_topk_prob=True, max_position_embeddings=128, rms_norm_eps=1e-6,
        pad_token_id=0, tie_word_embeddings=False, **kw)
    m = (transformers.DeepseekV2ForCausalLM if v2 else transformers.DeepseekV3ForCausalLM)(hc).eval().float()
    for mod in m.modules():          # transformers 5.5: DeepseekV2Moe's group-limited branch reads an attribute its __init__ never sets
        if "topk_method" in vars(mod):
            mod.num_experts = hc.n_routed_experts
    _randomise_(m)
    cfg = ModelConfig.from_hf_config(m.config.to_dict())
    if v2:
        assert cfg.moe_router == "softmax" and cfg.shared_expert_intermediate_size == 256 and cfg.n_group == (4 if variant == "v2_group_limited" else 1)
    assert cfg.attn_type == "mla" and cfg.head_dim == 48 and cfg.kv_cache_layout() == (1, 48, False)
    assert bool(cfg.num_experts) == (variant != "dense") and cfg.dense_layers == (() if variant == "dense" else (0,))
    w = TransformerWeights(cfg, "cpu", torch.float32, init=False)
    w.load_state_dict(tf.convert_hf_state_dict(cfg, m.state_dict()))
    assert sum(t.numel() for _, t in w.named_tensors()) == cfg.num_params()
    ids = torch.randint(3, 320, (2, 33))
    with torch.no_grad():
        want = m(ids).logits
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        got = tf.forward_hidden(cfg, w, inp) @ tf.lm_head_weight(cfg, w.head, w.embed).t()
    assert (got - want.reshape(-1, 320)).abs().max().item() < 3e-4
    # paged engine over the latent pool: one entry of kv_lora_rank + rope numbers per token and layer, no value pool
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=4, max_model_len=64, kv_cache_tokens=4 * 64)
    eng.set_weights(w.embed, w.layers, w.head)
    eng.init_kv_cache()
    assert eng.k_cache[0].shape[1:] == (1, eng.k_cache[0].shape[2], 48) and eng.v_cache[0].data_ptr() == eng.k_cache[0].data_ptr()
    assert eng.kv_bytes_per_token() == n_layers * 48 * 4
    prompt = ids[0, :21].tolist()
    with torch.no_grad():
        hf_out = m.generate(torch.tensor([prompt]), max_new_tokens=6, do_sample=False, eos_token_id=None, pad_token_id=0)[0, 21:].tolist()
    sp = SamplingParams(max_new_tokens=6, temperature=0.0, ignore_eos=True)
    assert eng.generate([prompt], sp)[0][0].output_ids == hf_out
    # chunked prefill (later chunks attend over the cached latent prefix) and the n-samples-share-a-prefill path
    eng2 = RolloutEngine(cfg, "cpu", max_num_seqs=4, max_model_len=64, kv_cache_tokens=4 * 64, max_prefill_tokens=8)
    eng2.set_weights(w.embed, w.layers, w.head)
    outs = eng2.generate([prompt], SamplingParams(max_new_tokens=6, temperature=0.0, ignore_eos=True, n=2))[0]
    assert [o.output_ids for o in outs] == [hf_out, hf_out]
    # ZeRO engine: same log-probs and gradient norm as plain autograd
    labels = torch.randint(0, 320, (inp.input_ids.numel(),))
    e = Zero3Engine(cfg, "cpu", lr=1e-3, weight_decay=0.0, seed=2, max_grad_norm=1e9)
    for u, src in [(e.embed, w.embed), (e.head, w.head)] + list(zip(e.layers, w.layers)):
        u.load_full_state(src)
    w.requires_grad_(True)
    lp_o = torch.log_softmax((tf.forward_hidden(cfg, w, inp) @ tf.lm_head_weight(cfg, w.head, w.embed).t()).float(), -1).gather(1, labels[:, None]).squeeze(1)
    (-lp_o.mean()).backward()
    gn = math.sqrt(sum(float(t.grad.pow(2).sum()) for _, t in w.named_tensors() if t.grad is not None))
    lp, _ = e.forward_logprob_entropy(inp, labels, 1.0)
    assert (lp - lp_o.detach()).abs().max() < 1e-5
    (-lp.mean()).backward()
    assert abs(e.step()["grad_norm"] - gn) < 1e-4 * max(1.0, gn)
    # checkpoint directory out and back, and HF reloads it with the same logits
    w.requires_grad_(False)
    e2 = Zero3Engine(cfg, "cpu", seed=5)
    for u, src in [(e2.embed, w.embed), (e2.head, w.head)] + list(zip(e2.layers, w.layers)):
        u.load_full_state(src)
    hf_io.save_hf(e2, cfg, tmp_path / "ck", dtype=torch.float32)
    m2 = type(m).from_pretrained(tmp_path / "ck", dtype=torch.float32).eval()
    for mod in m2.modules():
        if "topk_method" in vars(mod):
            mod.num_experts = cfg.num_experts
    with torch.no_grad():
        assert (m2(ids).logits - want).abs().max() < 1e-4
    cfg2 = get_model_config(str(tmp_path / "ck"))
    assert (cfg2.moe_router, cfg2.n_group, cfg2.routed_scaling_factor) == (cfg.moe_router, cfg.n_group, cfg.routed_scaling_factor)
    assert (cfg2.attn_type, cfg2.q_lora_rank, cfg2.kv_lora_rank, cfg2.rope_interleaved, cfg2.rope_scaling, cfg2.dense_layers, cfg2.num_experts) == \
        (cfg.attn_type, cfg.q_lora_rank, cfg.kv_lora_rank, cfg.rope_interleaved, cfg.rope_scaling, cfg.dense_layers, cfg.num_experts)
    back = hf_io.load_state_dict(cfg2, tmp_path / "ck")
    mine = w.state_dict()
    assert set(back) == set(mine) and all(torch.allclose(back[k], mine[k]) for k in mine)


def test_deepseek_v3_worker_rollout_logprob_update():
    """The MLA + MoE preset through the hybrid worker: rollout over the latent cache, old log-probs, one streamed PPO update;
    engine log-probs agree with the trainer's recomputation (same weights, fp32)."""
    import numpy as np

    from polyrl_b200.config.schema import load_config
    from polyrl_b200.protocol import DataProto
    from polyrl_b200.workers.hybrid_worker import ActorRolloutRefWorker

    g = torch.Generator().manual_seed(3)
    c = load_config(None, ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-deepseek-v3", "data.max_prompt_length=12",
                           "data.max_response_length=6", "actor_rollout_ref.rollout.max_num_seqs=8", "actor_rollout_ref.rollout.n=2",
                           "actor_rollout_ref.rollout.cuda_graph=False", "actor_rollout_ref.actor.ppo_mini_batch_size=4",
                           "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=2", "actor_rollout_ref.actor.router_aux_loss_coef=0.01"])
    wk = ActorRolloutRefWorker(c, role="actor_rollout", device="cpu")
    wk.init_model()
    pids = torch.randint(3, 200, (2, 12), generator=g)
    dp = DataProto.from_dict({"input_ids": pids, "attention_mask": torch.ones_like(pids), "position_ids": torch.arange(12).expand(2, -1).contiguous()},
                             non_tensors={"uid": np.array(["a", "b"], dtype=object)})
    dp.meta_info.update(temperature=1.0, ignore_eos=True)
    out = wk.generate_sequences(dp)
    out = out.union(wk.compute_log_prob(out))
    mask = out.batch["response_mask"].bool()
    assert (out.batch["rollout_log_probs"] - out.batch["old_log_probs"])[mask].abs().max() < 1e-4
    out.batch["advantages"] = torch.randn(4, 6, generator=g) * out.batch["response_mask"]
    out.meta_info.update({"is_opt_step": True})
    m = wk.update_actor_stream(out).meta_info["metrics"]
    assert np.isfinite(m["actor/grad_norm"]).all() and "actor/router_aux_loss" in m
