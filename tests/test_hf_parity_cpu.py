"""Architecture parity with HuggingFace reference implementations (transformers, CPU, random tiny models): after
``convert_hf_state_dict`` our packed forward must reproduce the HF logits for Llama, Qwen2 (qkv bias) and Qwen3 (per-head q/k
RMSNorm, tied embeddings) — this pins the RoPE convention, GQA head mapping, fused-qkv / interleaved gate-up layouts and the
norm placement that checkpoint import / export rely on (reference models: the families PolyRL trains through HF AutoModel,
3rdparty/verl/verl/workers/fsdp_workers.py:269-577; example workload Qwen3-1.7B)."""
import pytest
import torch

from polyrl_b200.models import TransformerWeights, forward_hidden, lm_head_weight, pack_padded
from polyrl_b200.models import transformer as tf
from polyrl_b200.models.config import ModelConfig, get_model_config

transformers = pytest.importorskip("transformers")


def _hf(name):
    common = dict(vocab_size=320, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                  max_position_embeddings=256, rope_theta=10000.0)
    if name == "llama":
        return transformers.LlamaForCausalLM(transformers.LlamaConfig(rms_norm_eps=1e-5, tie_word_embeddings=False, attention_bias=False, **common))
    if name == "mistral":
        return transformers.MistralForCausalLM(transformers.MistralConfig(rms_norm_eps=1e-5, tie_word_embeddings=False, sliding_window=64,
                                                                          **common))
    if name == "qwen2":
        return transformers.Qwen2ForCausalLM(transformers.Qwen2Config(rms_norm_eps=1e-6, tie_word_embeddings=False, **common))
    return transformers.Qwen3ForCausalLM(transformers.Qwen3Config(rms_norm_eps=1e-6, head_dim=32, tie_word_embeddings=True, attention_bias=False, **common))


@pytest.mark.parametrize("name", ["llama", "mistral", "qwen2", "qwen3"])
def test_logits_match_hf(name):
    torch.manual_seed(0)
    m = _hf(name).eval().float()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n or n.endswith(".bias"):
                p.add_(torch.randn_like(p) * 0.1)               # gains / biases away from their trivial init
    cfg = ModelConfig.from_hf_config(m.config.to_dict())
    assert cfg.qkv_bias == (name == "qwen2") and cfg.qk_norm == (name == "qwen3")
    sd = tf.convert_hf_state_dict(cfg, m.state_dict())
    w = TransformerWeights(cfg, "cpu", torch.float32, init=False)
    w.load_state_dict(sd)
    ids = torch.randint(0, 320, (3, 23))
    mask = torch.ones_like(ids)
    mask[1, 15:] = 0                                               # right padding: packed path drops it
    with torch.no_grad():
        want = m(ids, attention_mask=mask).logits
        inp, idx = pack_padded(ids, mask)
        got = forward_hidden(cfg, w, inp) @ lm_head_weight(cfg, w.head, w.embed).t()
    keep = mask.bool()
    assert (got - want[keep]).abs().max().item() < 2e-4
    if name == "mistral":                                           # the window is recorded, exported, and enforced
        from polyrl_b200.models.hf_io import hf_config_dict

        out = hf_config_dict(cfg)
        assert cfg.sliding_window == 64 and out["model_type"] == "mistral" and out["sliding_window"] == 64
        long_ids = torch.randint(0, 320, (1, 80))
        with pytest.raises(NotImplementedError, match="attention window"):
            forward_hidden(cfg, w, pack_padded(long_ids, torch.ones_like(long_ids))[0])
    back = tf.export_hf_state_dict(cfg, sd)                         # export restores every HF tensor bit for bit
    hf_sd = m.state_dict()
    assert all(torch.equal(back[k], hf_sd[k]) for k in back if k in hf_sd) and len([k for k in back if k in hf_sd]) >= len(hf_sd) - 1


def test_qk_norm_engine_matches_cacheless_oracle():
    """Qwen3-style block through the paged engine (prefill + decode paths) vs the cache-less oracle."""
    from polyrl_b200.models import get_model_config
    from polyrl_b200.rollout.registry import get_rollout_cls

    cfg = get_model_config("tiny-qwen3")
    w = TransformerWeights(cfg, "cpu", torch.float32, seed=5)
    for lw in w.layers:
        lw["q_norm"].add_(torch.randn(cfg.head_dim) * 0.2)
        lw["k_norm"].add_(torch.randn(cfg.head_dim) * 0.2)
    prompts = [[5, 9, 33, 7, 2], [12, 4, 4, 8, 19, 3, 77, 41]]
    outs = {}
    for name in ("native", "naive"):
        eng = get_rollout_cls(name)(cfg, "cpu", max_model_len=64, seed=0, **({"max_num_seqs": 4, "use_cuda_graph": False} if name == "native" else {}))
        eng.set_weights(w.embed, w.layers, w.head)
        g = eng.generate(prompts, {"max_new_tokens": 8, "temperature": 0.0, "ignore_eos": True})
        outs[name] = [(r.output_ids, r.logprobs) for grp in g for r in grp]
    for (a, la), (b, lb) in zip(outs["native"], outs["naive"]):
        assert a == b and max(abs(x - y) for x, y in zip(la, lb)) < 1e-4


def test_gpt2_logits_match_hf():
    """BASELINE config 1 trains a GPT-2-family PPO model: the gpt2 block (LayerNorm, learned positions, fused c_attn, GELU-tanh
    MLP, biases, tied head) must equal HF GPT2LMHeadModel."""
    torch.manual_seed(1)
    hc = transformers.GPT2Config(vocab_size=300, n_positions=64, n_embd=96, n_layer=2, n_head=4, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    m = transformers.GPT2LMHeadModel(hc).eval().float()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "ln_" in n or n.endswith(".bias"):
                p.add_(torch.randn_like(p) * 0.1)
    cfg = ModelConfig.from_hf_config(hc.to_dict())
    assert cfg.family == "gpt2" and cfg.num_heads == 4
    sd = tf.convert_hf_state_dict(cfg, m.state_dict())
    w = TransformerWeights(cfg, "cpu", torch.float32, init=False)
    w.load_state_dict(sd)
    ids = torch.randint(0, 300, (2, 19))
    with torch.no_grad():
        want = m(ids).logits
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        got = (forward_hidden(cfg, w, inp) @ lm_head_weight(cfg, w.head, w.embed).t()).view(2, 19, -1)
    assert (got - want).abs().max().item() < 2e-4
    back = tf.export_hf_state_dict(cfg, sd)
    hf_sd = m.state_dict()
    assert all(torch.equal(back[k], hf_sd[k]) for k in back if k in hf_sd)


def test_hf_checkpoint_dir_roundtrip_through_engine(tmp_path):
    """save_pretrained dir (safetensors) -> ZeRO engine (unit-by-unit load) -> same log-probs as HF; engine -> save_hf ->
    AutoModel loads it back and agrees."""
    from polyrl_b200.models import get_model_config
    from polyrl_b200.models.hf_io import HFCheckpoint, load_into_engine, load_state_dict, save_hf
    from polyrl_b200.parallel.zero3 import Zero3Engine

    torch.manual_seed(2)
    m = _hf("qwen3").eval().float()
    m.save_pretrained(tmp_path / "hf")
    cfg = get_model_config(str(tmp_path / "hf"))
    assert cfg.qk_norm and cfg.tie_word_embeddings and HFCheckpoint(tmp_path / "hf").has_weights()
    eng = Zero3Engine(cfg, "cpu", seed=0, trainable=False)
    assert load_into_engine(eng, tmp_path / "hf") > 10
    ids = torch.randint(0, 320, (2, 21))
    inp, _ = pack_padded(ids, torch.ones_like(ids))
    labels = torch.roll(inp.input_ids, -1)
    with torch.no_grad():
        lp, _ = eng.forward_logprob_entropy(inp, labels, 1.0)
        want = torch.log_softmax(m(ids).logits.float(), -1).reshape(-1, 320).gather(-1, labels[:, None]).squeeze(-1)
    assert (lp - want).abs().max().item() < 2e-4
    sd = load_state_dict(cfg, tmp_path / "hf")
    assert set(sd) == {n for n, _ in eng.named_full_tensors()}
    names = save_hf(eng, cfg, tmp_path / "export", dtype=torch.float32)
    assert names == ["model.safetensors"]
    m2 = transformers.AutoModelForCausalLM.from_pretrained(tmp_path / "export").eval().float()
    with torch.no_grad():
        assert (m2(ids).logits - m(ids).logits).abs().max().item() < 1e-5
    # sharded export + index
    assert len(save_hf(sd, cfg, tmp_path / "export2", max_shard_bytes=200_000, dtype=torch.float32)) > 1
    assert (tmp_path / "export2" / "model.safetensors.index.json").exists()
    sd2 = load_state_dict(cfg, tmp_path / "export2")
    assert all(torch.equal(sd[k].float(), sd2[k].float()) for k in sd)


def test_worker_starts_from_hf_dir_and_checkpoint_writes_hf_model(tmp_path):
    from polyrl_b200.config.schema import load_config
    from polyrl_b200.utils.checkpoint import ShardedCheckpointManager
    from polyrl_b200.workers.hybrid_worker import ActorRolloutRefWorker

    torch.manual_seed(3)
    m = _hf("llama").eval().float()
    m.save_pretrained(tmp_path / "hf")
    cfg = load_config(None, ["trainer.device=cpu", f"actor_rollout_ref.model.path={tmp_path / 'hf'}", "data.max_prompt_length=8",
                             "data.max_response_length=4", "actor_rollout_ref.rollout.max_num_seqs=4", "actor_rollout_ref.rollout.cuda_graph=False"])
    wk = ActorRolloutRefWorker(cfg, role="actor_rollout", device="cpu")
    wk.init_model()
    got = dict(wk.actor_engine.named_full_tensors())
    want = tf.convert_hf_state_dict(wk.model_cfg, m.state_dict())
    assert all(torch.allclose(got[k].float(), want[k].float(), atol=1e-6) for k in want)
    # the colocated engine decodes with the loaded weights: greedy first token == HF argmax
    ids = [5, 9, 33, 7, 2]
    g = wk.rollout.generate([ids], {"max_new_tokens": 1, "temperature": 0.0, "ignore_eos": True})
    with torch.no_grad():
        assert g[0][0].output_ids[0] == int(m(torch.tensor([ids])).logits[0, -1].argmax())
    mgr = ShardedCheckpointManager(wk.actor_engine, wk.model_cfg, save_contents=("model", "hf_model"))
    mgr.save(str(tmp_path / "ck"), global_step=1)
    m2 = transformers.AutoModelForCausalLM.from_pretrained(tmp_path / "ck" / "huggingface", torch_dtype=torch.float32).eval()
    x = torch.randint(0, 320, (1, 9))
    with torch.no_grad():
        assert (m2(x).logits - m(x).logits).abs().max().item() < 2e-2          # exported in bf16


def test_critic_loads_backbone_from_hf_dir_and_keeps_value_head(tmp_path):
    from polyrl_b200.config.schema import load_config
    from polyrl_b200.workers.hybrid_worker import CriticWorker

    torch.manual_seed(4)
    m = _hf("llama").eval().float()
    m.save_pretrained(tmp_path / "hf")
    cfg = load_config(None, ["trainer.device=cpu", f"actor_rollout_ref.model.path={tmp_path / 'hf'}", f"critic.model.path={tmp_path / 'hf'}",
                             "algorithm.adv_estimator=gae"])
    cw = CriticWorker(cfg, device="cpu")
    cw.init_model()
    got = dict(cw.engine.named_full_tensors())
    want = tf.convert_hf_state_dict(cw.model_cfg, m.state_dict())
    assert torch.allclose(got["layers.1.wd"].float(), want["layers.1.wd"].float(), atol=1e-6) and "v_head" in got and "lm_head" not in got
    assert got["v_head"][0].abs().sum() > 0


@pytest.mark.parametrize("scaling", [
    dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=64),
    dict(rope_type="linear", factor=4.0),
    dict(rope_type="yarn", factor=4.0, original_max_position_embeddings=64),
])
def test_long_context_rope_scaling_matches_hf(scaling):
    """Llama-3.1-style / linear / YaRN rotary scaling declared in the HF config: logits equal HF on sequences longer than the
    original context, and the exported config carries the scaling back."""
    torch.manual_seed(0)
    hc = transformers.LlamaConfig(vocab_size=320, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                                  num_key_value_heads=2, max_position_embeddings=256, rms_norm_eps=1e-5, tie_word_embeddings=False,
                                  rope_parameters={**scaling, "rope_theta": 10000.0})
    m = transformers.LlamaForCausalLM(hc).eval().float()
    cfg = ModelConfig.from_hf_config(m.config.to_dict())
    assert cfg.rope_scaling["rope_type"] == scaling["rope_type"]
    w = TransformerWeights(cfg, "cpu", torch.float32, init=False)
    w.load_state_dict(tf.convert_hf_state_dict(cfg, m.state_dict()))
    ids = torch.randint(0, 320, (2, 150))
    with torch.no_grad():
        want = m(ids).logits
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        got = forward_hidden(cfg, w, inp) @ lm_head_weight(cfg, w.head, w.embed).t()
    assert (got - want.reshape(-1, 320)).abs().max().item() < 3e-4
    plain = ModelConfig.from_hf_config({**m.config.to_dict(), "rope_parameters": {"rope_type": "default", "rope_theta": 10000.0}})
    with torch.no_grad():
        off = forward_hidden(plain, w, inp) @ lm_head_weight(cfg, w.head, w.embed).t()
    assert (off - want.reshape(-1, 320)).abs().max().item() > 1e-3            # the scaling really matters at these lengths
    from polyrl_b200.models.hf_io import hf_config_dict

    assert ModelConfig.from_hf_config(hf_config_dict(cfg)).rope_scaling == cfg.rope_scaling


@pytest.mark.parametrize("family", ["glm", "glm4"])
def test_glm_families_match_hf(family, tmp_path):
    """GLM-4-9B-style (``glm``: partial interleaved rotary, fused gate_up projection, qkv bias) and GLM-4-0414-style (``glm4``: the
    same plus RMSNorms on the attention and MLP outputs): logits equal HF, greedy generation through the paged engine equals HF
    generate, and an exported checkpoint loads back into HF with identical logits."""
    from polyrl_b200.models import hf_io
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams

    torch.manual_seed(0)
    cls, ccls = ((transformers.GlmForCausalLM, transformers.GlmConfig) if family == "glm"
                 else (transformers.Glm4ForCausalLM, transformers.Glm4Config))
    hc = ccls(vocab_size=320, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
              head_dim=32, max_position_embeddings=128, rms_norm_eps=1e-5, attention_bias=True, pad_token_id=0, eos_token_id=[1, 2],
              tie_word_embeddings=False)
    m = cls(hc).eval().float()
    with torch.no_grad():                                   # norms away from their all-ones init so a missing one shows
        for n, p_ in m.named_parameters():
            if "layernorm" in n or n.endswith("norm.weight"):
                p_.add_(0.1 * torch.randn_like(p_))
    cfg = ModelConfig.from_hf_config(m.config.to_dict())
    assert cfg.partial_rotary_factor == 0.5 and cfg.rope_interleaved and cfg.sandwich_norm == (family == "glm4")
    w = TransformerWeights(cfg, "cpu", torch.float32, init=False)
    w.load_state_dict(tf.convert_hf_state_dict(cfg, m.state_dict()))
    ids = torch.randint(3, 320, (2, 40))
    with torch.no_grad():
        want = m(ids).logits
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        got = forward_hidden(cfg, w, inp) @ lm_head_weight(cfg, w.head, w.embed).t()
    assert (got - want.reshape(-1, 320)).abs().max().item() < 3e-4
    # decode path (paged KV, per-token rope) against HF greedy generation
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=2, max_model_len=96, kv_cache_tokens=2 * 96)
    eng.set_weights(w.embed, w.layers, w.head)
    prompt = ids[0, :17].tolist()
    out = eng.generate([prompt], SamplingParams(max_new_tokens=8, temperature=0.0, ignore_eos=True))[0][0]
    with torch.no_grad():
        hf_out = m.generate(torch.tensor([prompt]), max_new_tokens=8, do_sample=False, eos_token_id=None, pad_token_id=0)[0, 17:].tolist()
    assert out.output_ids == hf_out
    # export -> HF
    sd = tf.export_hf_state_dict(cfg, w.state_dict())
    m2 = cls(ccls(**{k: v for k, v in hf_io.hf_config_dict(cfg).items() if k not in ("architectures", "torch_dtype", "model_type")})).eval().float()
    missing, unexpected = m2.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in k for k in missing), (missing, unexpected)
    with torch.no_grad():
        assert (m2(ids).logits - want).abs().max().item() < 1e-5


def test_value_head_round_trips_through_hf_token_classification(tmp_path):
    """Critics / reward models: an HF ``...ForTokenClassification`` checkpoint with one label (``score`` weight + bias) loads into the
    value-head engine and gives the same per-token values; a value-head engine saved in HF format loads back into HF with equal
    outputs; a plain LM checkpoint still initialises a critic (fresh head, loaded trunk)."""
    from polyrl_b200.models import hf_io
    from polyrl_b200.parallel.zero3 import Zero3Engine

    torch.manual_seed(0)
    hc = transformers.Qwen2Config(vocab_size=320, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                                  num_key_value_heads=2, max_position_embeddings=128, rms_norm_eps=1e-6, tie_word_embeddings=False,
                                  num_labels=1, classifier_dropout=0.0)
    m = transformers.Qwen2ForTokenClassification(hc).eval().float()
    with torch.no_grad():
        m.score.bias.fill_(0.37)
    m.save_pretrained(tmp_path / "rm", safe_serialization=True)
    cfg = get_model_config(str(tmp_path / "rm"), value_head=True)
    eng = Zero3Engine(cfg, "cpu", seed=3)
    assert hf_io.load_into_engine(eng, tmp_path / "rm") > 0
    ids = torch.randint(3, 320, (2, 19))
    inp, _ = pack_padded(ids, torch.ones_like(ids))
    with torch.no_grad():
        want = m(ids).logits[..., 0].reshape(-1)
        got = eng.forward_values(inp)
    assert (got - want).abs().max() < 2e-4 and want.abs().max() > 0.1
    # engine -> HF directory -> HF model
    hf_io.save_hf(eng, cfg, tmp_path / "out", dtype=torch.float32)
    m2 = transformers.AutoModelForTokenClassification.from_pretrained(tmp_path / "out", dtype=torch.float32).eval()
    assert type(m2).__name__ == "Qwen2ForTokenClassification"
    with torch.no_grad():
        assert (m2(ids).logits[..., 0].reshape(-1) - want).abs().max() < 2e-4
    # a language-model checkpoint: trunk loaded, head left as initialised
    lm = transformers.Qwen2ForCausalLM(hc).eval().float()
    lm.save_pretrained(tmp_path / "lm", safe_serialization=True)
    crit = Zero3Engine(get_model_config(str(tmp_path / "lm"), value_head=True), "cpu", seed=4)
    head0 = {n: t.clone() for n, t in crit.named_full_tensors() if n in ("v_head", "v_head_b")}
    hf_io.load_into_engine(crit, tmp_path / "lm")
    after = dict(crit.named_full_tensors())
    assert all(torch.equal(after[n], head0[n]) for n in head0)
    assert torch.equal(after["layers.0.wo"], lm.state_dict()["model.layers.0.self_attn.o_proj.weight"])


def test_apertus_family_matches_hf(tmp_path):
    """Apertus (reference: 3rdparty/verl/verl/models/transformers/apertus.py): Qwen3-style attention with per-head QK norm, llama3 rotary
    scaling by default, and a NON-gated MLP down(xIELU(up(x))) with two learned activation scalars per block: logits equal HF,
    greedy generation through the paged engine equals HF generate, ZeRO gradients (incl. the activation scalars) equal plain
    autograd, and the HF checkpoint directory round-trips."""
    import math

    from polyrl_b200.models import hf_io
    from polyrl_b200.parallel.zero3 import Zero3Engine
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams

    torch.manual_seed(0)
    hc = transformers.ApertusConfig(vocab_size=320, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                                    num_key_value_heads=2, max_position_embeddings=128, rms_norm_eps=1e-5, tie_word_embeddings=False,
                                    pad_token_id=0, hidden_act="xielu", dtype="float32")
    m = transformers.ApertusForCausalLM(hc).eval().float()
    with torch.no_grad():
        for n, p_ in m.named_parameters():
            if p_.dim() == 1:                               # norms and the activation scalars away from their init
                p_.add_(0.1 * torch.randn_like(p_))
    cfg = ModelConfig.from_hf_config(m.config.to_dict())
    assert cfg.mlp_act == "xielu" and cfg.qk_norm and cfg.rope_scaling["rope_type"] == "llama3"
    w = TransformerWeights(cfg, "cpu", torch.float32, init=False)
    w.load_state_dict(tf.convert_hf_state_dict(cfg, m.state_dict()))
    assert sum(t.numel() for _, t in w.named_tensors()) == cfg.num_params() and "wup" in w.layers[0] and "wgu" not in w.layers[0]
    fresh = TransformerWeights(cfg, "cpu", torch.float32, seed=1).layers[0]            # our init == HF's activation init
    assert abs(fresh["act_ap"].item() - math.log(math.expm1(0.8))) < 1e-6 and abs(fresh["act_an"].item() - math.log(math.expm1(0.3))) < 1e-6
    ids = torch.randint(3, 320, (2, 40))
    with torch.no_grad():
        want = m(ids).logits
        inp, _ = pack_padded(ids, torch.ones_like(ids))
        got = forward_hidden(cfg, w, inp) @ lm_head_weight(cfg, w.head, w.embed).t()
    assert (got - want.reshape(-1, 320)).abs().max().item() < 3e-4
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=2, max_model_len=96, kv_cache_tokens=2 * 96)
    eng.set_weights(w.embed, w.layers, w.head)
    prompt = ids[0, :17].tolist()
    out = eng.generate([prompt], SamplingParams(max_new_tokens=8, temperature=0.0, ignore_eos=True))[0][0]
    with torch.no_grad():
        hf_out = m.generate(torch.tensor([prompt]), max_new_tokens=8, do_sample=False, eos_token_id=None, pad_token_id=0)[0, 17:].tolist()
    assert out.output_ids == hf_out
    # ZeRO engine == plain autograd
    labels = torch.randint(0, 320, (inp.input_ids.numel(),))
    e = Zero3Engine(cfg, "cpu", lr=1e-3, weight_decay=0.0, seed=2, max_grad_norm=1e9)
    for u, src in [(e.embed, w.embed), (e.head, w.head)] + list(zip(e.layers, w.layers)):
        u.load_full_state(src)
    w.requires_grad_(True)
    lp_o = torch.log_softmax((forward_hidden(cfg, w, inp) @ lm_head_weight(cfg, w.head, w.embed).t()).float(), -1).gather(1, labels[:, None]).squeeze(1)
    (-lp_o.mean()).backward()
    assert w.layers[0]["act_ap"].grad.abs().item() > 0 and w.layers[1]["act_an"].grad.abs().item() > 0
    gn = math.sqrt(sum(float(t.grad.pow(2).sum()) for _, t in w.named_tensors() if t.grad is not None))
    lp, _ = e.forward_logprob_entropy(inp, labels, 1.0)
    assert (lp - lp_o.detach()).abs().max() < 1e-5
    (-lp.mean()).backward()
    assert abs(e.step()["grad_norm"] - gn) < 1e-4 * max(1.0, gn)
    # checkpoint directory out and back
    w.requires_grad_(False)
    e2 = Zero3Engine(cfg, "cpu", seed=5)
    for u, src in [(e2.embed, w.embed), (e2.head, w.head)] + list(zip(e2.layers, w.layers)):
        u.load_full_state(src)
    hf_io.save_hf(e2, cfg, tmp_path / "ck", dtype=torch.float32)
    m2 = transformers.ApertusForCausalLM.from_pretrained(tmp_path / "ck", dtype=torch.float32).eval()
    with torch.no_grad():
        assert (m2(ids).logits - want).abs().max() < 1e-4
    cfg2 = get_model_config(str(tmp_path / "ck"))
    assert cfg2.mlp_act == "xielu" and cfg2.rope_scaling == cfg.rope_scaling
    back = hf_io.load_state_dict(cfg2, tmp_path / "ck")
    mine = w.state_dict()
    assert set(back) == set(mine) and all(torch.allclose(back[k], mine[k]) for k in mine)
