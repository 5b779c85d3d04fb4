"""Architecture parity with HuggingFace reference implementations (transformers, CPU, random tiny models): after
``convert_hf_state_dict`` our packed forward must reproduce the HF logits for Llama, Qwen2 (qkv bias) and Qwen3 (per-head q/k
RMSNorm, tied embeddings) — this pins the RoPE convention, GQA head mapping, fused-qkv / interleaved gate-up layouts and the
norm placement that checkpoint import / export rely on (reference models: the families PolyRL trains through HF AutoModel,
3rdparty/verl/verl/workers/fsdp_workers.py:269-577; example workload Qwen3-1.7B)."""
import pytest
import torch

from polyrl_b200.models import TransformerWeights, forward_hidden, lm_head_weight, pack_padded
from polyrl_b200.models import transformer as tf
from polyrl_b200.models.config import ModelConfig

transformers = pytest.importorskip("transformers")


def _hf(name):
    common = dict(vocab_size=320, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                  max_position_embeddings=256, rope_theta=10000.0)
    if name == "llama":
        return transformers.LlamaForCausalLM(transformers.LlamaConfig(rms_norm_eps=1e-5, tie_word_embeddings=False, attention_bias=False, **common))
    if name == "qwen2":
        return transformers.Qwen2ForCausalLM(transformers.Qwen2Config(rms_norm_eps=1e-6, tie_word_embeddings=False, **common))
    return transformers.Qwen3ForCausalLM(transformers.Qwen3Config(rms_norm_eps=1e-6, head_dim=32, tie_word_embeddings=True, attention_bias=False, **common))


@pytest.mark.parametrize("name", ["llama", "qwen2", "qwen3"])
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
