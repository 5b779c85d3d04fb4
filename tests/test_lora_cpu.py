"""LoRA: frozen base + trainable adapter units (reference capability: peft LoRA through FSDP, fsdp_workers.py:269-577)."""
import torch

from polyrl_b200.models import TransformerWeights, get_model_config, pack_padded
from polyrl_b200.models import transformer as tf
from polyrl_b200.models.lora import LoraConfig, export_peft_state_dict, lora_param_specs, merge_layer_
from polyrl_b200.parallel.zero3 import Zero3Engine


def _batch(cfg):
    torch.manual_seed(0)
    ids = torch.randint(3, cfg.vocab_size, (3, 19))
    mask = torch.ones_like(ids)
    mask[2, 11:] = 0
    inp, _ = pack_padded(ids, mask)
    labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),))
    return inp, labels


def test_lora_trains_adapters_only_and_merges():
    cfg = get_model_config("tiny-llama")
    lcfg = LoraConfig.from_model_cfg(4, 8, "all-linear")
    eng = Zero3Engine(cfg, "cpu", lr=5e-2, weight_decay=0.0, seed=3, lora=lcfg)
    assert all(u.master is None for u in eng.base_units) and all(u.master is not None for u in eng.lora_layers)
    assert len(eng.opt_units) == cfg.num_layers
    inp, labels = _batch(cfg)
    base0 = {n: t.clone() for n, t in eng.named_full_tensors() if ".lora." not in n}
    lp0, _ = eng.forward_logprob_entropy(inp, labels, 1.0)
    # B = 0 at init: the adapter path is a no-op
    plain = Zero3Engine(cfg, "cpu", seed=3, trainable=False)
    lp_plain, _ = plain.forward_logprob_entropy(inp, labels, 1.0)
    assert torch.allclose(lp0.detach(), lp_plain, atol=1e-5)
    losses = []
    for _ in range(4):
        lp, _ = eng.forward_logprob_entropy(inp, labels, 1.0)
        loss = -lp.mean()
        loss.backward()
        st = eng.step()
        assert st["grad_norm"] > 0 and not st["skipped"]
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 1e-3                                     # adapters learn
    named = dict((n, t.clone()) for n, t in eng.named_full_tensors())
    for n, t in base0.items():
        assert torch.equal(named[n], t), n                                   # base is frozen
    assert any(t.abs().sum() > 0 for n, t in named.items() if n.endswith("_B"))
    # merged weights reproduce the adapter forward
    w = TransformerWeights(cfg, "cpu", torch.float32, init=False)
    w.load_state_dict({n: t for n, t in named.items() if ".lora." not in n})
    for i in range(cfg.num_layers):
        merge_layer_(cfg, w.layers[i], {k.split(".lora.")[1]: v for k, v in named.items() if k.startswith(f"layers.{i}.lora.")}, lcfg.scale)
    x = tf.forward_hidden(cfg, w, inp)
    from polyrl_b200 import ops

    lp_merged, _ = ops.linear_cross_entropy(tf.final_norm(cfg, w.head, x), tf.lm_head_weight(cfg, w.head, w.embed), labels, 1.0)
    with torch.no_grad():
        lp_adapter, _ = eng.forward_logprob_entropy(inp, labels, 1.0)
    assert torch.allclose(lp_merged, lp_adapter, atol=2e-4)
    peft = export_peft_state_dict(named)
    assert len(peft) == cfg.num_layers * len(lora_param_specs(cfg, lcfg))
    assert "base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight" in peft


def test_lora_partial_targets_and_recompute_modes_agree():
    cfg = get_model_config("tiny-llama")
    lcfg = LoraConfig.from_model_cfg(2, 4, "q_proj,v_proj,down_proj")
    inp, labels = _batch(cfg)
    outs = []
    for mode in ("full", "none"):
        eng = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=5, lora=lcfg)
        eng.act_ckpt = mode
        for _ in range(2):
            lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
            (-(lp.mean()) + 0.01 * ent.mean()).backward()
            eng.step()
        outs.append({n: t.clone() for n, t in eng.named_full_tensors() if ".lora." in n})
    assert set(outs[0]) == set(outs[1]) and len(outs[0]) == cfg.num_layers * 6
    for n in outs[0]:
        assert torch.allclose(outs[0][n], outs[1][n], atol=1e-6), n


def test_lora_colocated_grpo_loop():
    """End to end: LoRA actor, adapters merged into the colocated rollout engine's replica every step."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    ov = ["trainer.device=cpu", "data.synthetic=True", "data.synthetic_num_samples=32", "data.max_prompt_length=24",
          "data.synthetic_prompt_len_range=[6,16]", "data.max_response_length=8", "data.train_batch_size=4",
          "actor_rollout_ref.rollout.n=2", "actor_rollout_ref.rollout.min_stream_batch_size=4",
          "actor_rollout_ref.actor.ppo_mini_batch_size=2", "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=2",
          "actor_rollout_ref.rollout.log_prob_micro_batch_size_per_gpu=4", "actor_rollout_ref.rollout.max_num_seqs=8",
          "actor_rollout_ref.actor.optim.lr=1e-3", "trainer.logger=[console]", "trainer.total_epochs=1",
          "actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "trainer.n_gpus_per_node=1",
          "trainer.total_training_steps=2", "actor_rollout_ref.model.lora_rank=4", "actor_rollout_ref.model.lora_alpha=8"]
    env = {**os.environ, "PYTHONPATH": str(root), "OMP_NUM_THREADS": "2"}
    r = subprocess.run([sys.executable, "-m", "polyrl_b200.trainer.main_stream", *ov], capture_output=True, text=True, timeout=600,
                       env=env, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:2" in r.stdout and "actor/pg_loss" in r.stdout


def test_engine_api_sft_ppo_value_and_checkpoint(tmp_path):
    """BaseEngine surface (reference: verl/workers/engine): infer / train_batch with pluggable losses, optimizer + LR step,
    checkpoint round trip."""
    from polyrl_b200.workers.engine_api import get_engine_cls, ppo_loss, sft_loss, value_loss

    torch.manual_seed(0)
    b, p, r = 4, 6, 5
    ids = torch.randint(3, 200, (b, p + r))
    batch = {"input_ids": ids, "attention_mask": torch.ones_like(ids), "responses": ids[:, p:], "response_mask": torch.ones(b, r)}
    eng = get_engine_cls("language_model", "fsdp")("tiny-llama", "cpu", lr=5e-3, lr_schedule=lambda s: 5e-3 * 0.5 ** s, micro_batch_size=2,
                                                   weight_decay=0.0)
    eng.initialize()
    with eng.eval_mode():
        out0 = eng.infer_batch(batch)
    assert out0["log_probs"].shape == (b, r) and out0["entropy"].shape == (b, r)
    with eng.train_mode():
        l0 = eng.train_batch(batch, sft_loss)["sft/loss"]
        st = eng.optimizer_step()
        assert st["lr"] == 5e-3 and st["grad_norm"] > 0 and eng.lr_scheduler_step() == 2.5e-3
        l1 = eng.train_batch(batch, sft_loss)["sft/loss"]
        eng.optimizer_zero_grad()
        assert l1 < l0
        pb = {**batch, "old_log_probs": out0["log_probs"], "advantages": torch.randn(b, r)}
        m = eng.train_batch(pb, ppo_loss(clip_ratio=0.2))
        assert "actor/pg_loss" in m and "actor/ppo_kl" in m
        eng.optimizer_step()
    eng.save_checkpoint(str(tmp_path / "ck"), global_step=3)
    with eng.eval_mode():
        ref_lp = eng.infer_batch(batch)["log_probs"]
    eng2 = get_engine_cls("language_model")("tiny-llama", "cpu", micro_batch_size=4, seed=123)
    eng2.initialize()
    eng2.load_checkpoint(str(tmp_path / "ck"))
    with eng2.eval_mode():
        assert torch.allclose(eng2.infer_batch(batch)["log_probs"], ref_lp, atol=1e-5) and eng2.lr_step == 1
    veng = get_engine_cls("value_model")("tiny-llama", "cpu", lr=1e-3, micro_batch_size=4)
    veng.initialize()
    with veng.eval_mode():
        vals = veng.infer_batch(batch)["values"]
    assert vals.shape == (b, r)
    with veng.train_mode():
        m = veng.train_batch({**batch, "values": vals, "returns": vals + 1.0}, value_loss(0.5))
        assert m["critic/vf_loss"] > 0


def test_role_workers_over_base_engine_match_the_streaming_actor():
    """ActorWorker / CriticWorker (roles over BaseEngine): log-probs equal the hybrid worker's actor on the same weights, an update
    with two mini-batches takes two optimizer steps and lowers the surrogate loss; the critic role fits returns."""
    from polyrl_b200.protocol import DataProto
    from polyrl_b200.workers.engine_api import get_engine_cls
    from polyrl_b200.workers.roles import ActorWorker, CriticWorker

    g = torch.Generator().manual_seed(0)
    b, p, r = 8, 6, 5
    ids = torch.randint(3, 200, (b, p + r), generator=g)
    mask = torch.ones_like(ids)
    mask[1, :2] = 0
    data = DataProto.from_dict({"input_ids": ids, "attention_mask": mask, "position_ids": (mask.cumsum(-1) - 1).clamp_min(0),
                                "responses": ids[:, p:], "response_mask": mask[:, p:]})
    data.meta_info["temperature"] = 1.0
    actor = ActorWorker(get_engine_cls("language_model", "zero")("tiny-llama", "cpu", lr=1e-2, micro_batch_size=2, weight_decay=0.0, seed=4),
                        ppo_mini_batch_size=4, clip_ratio=0.2)
    actor.init_model()
    old = actor.compute_log_prob(data)
    assert old.batch["old_log_probs"].shape == (b, r) and old.batch["entropys"].min() > 0
    # oracle: the streaming actor on the same weights
    from polyrl_b200.workers.actor import ActorConfig, DataParallelPPOActor

    dpa = DataParallelPPOActor(actor.engine.engine, ActorConfig(ppo_mini_batch_size=8, ppo_micro_batch_size_per_gpu=4))
    lp2, _ = dpa.compute_log_prob(data)
    assert (lp2 - old.batch["old_log_probs"]).abs().max() < 1e-5
    data = data.union(old)
    data.batch["advantages"] = torch.randn(b, r, generator=g)
    step0 = actor.engine.engine.opt_step
    m = actor.update_actor(data)
    assert actor.engine.engine.opt_step == step0 + 2 and m["actor/grad_norm"] > 0 and "actor/pg_loss" in m
    new = actor.compute_log_prob(data).batch["old_log_probs"]
    adv, msk = data.batch["advantages"], data.batch["response_mask"].float()
    surrogate = lambda lp: -((torch.exp(lp - data.batch["old_log_probs"]) * adv) * msk).sum() / msk.sum()      # noqa: E731
    assert surrogate(new) < surrogate(data.batch["old_log_probs"])
    critic = CriticWorker(get_engine_cls("value_model", "zero")("tiny-llama", "cpu", lr=5e-4, micro_batch_size=4, weight_decay=0.0, seed=5),
                          ppo_mini_batch_size=8, ppo_epochs=4)
    critic.init_model()
    v0 = critic.compute_values(data).batch["values"]
    data.batch["values"], data.batch["returns"] = v0, torch.ones(b, r)
    m = critic.update_critic(data)
    v1 = critic.compute_values(data).batch["values"]
    assert "critic/vf_loss" in m and ((v1 - 1) ** 2).mean() < ((v0 - 1) ** 2).mean()


def test_peft_adapter_directory_round_trip(tmp_path):
    """model.lora_adapter_path: an exported peft adapter (names with and without the ``.default`` adapter segment) loads back into
    the engine's adapter units; a rank mismatch is rejected."""
    import json

    import pytest
    from safetensors.torch import save_file

    from polyrl_b200.models import get_model_config
    from polyrl_b200.models.lora import LoraConfig, export_peft_state_dict, import_peft_state_dict, load_peft_adapter, peft_adapter_config
    from polyrl_b200.parallel.zero3 import Zero3Engine

    cfg = get_model_config("tiny-llama")
    lc = LoraConfig.from_model_cfg(4, 8, "all-linear")
    src = Zero3Engine(cfg, "cpu", seed=1, lora=lc)
    with torch.no_grad():
        for u in src.lora_layers:
            u.shard.normal_(0, 0.1)
            if u.master is not None:
                u.master.copy_(u.shard.float())
    named = {n: t.clone() for n, t in src.named_full_tensors() if ".lora." in n}
    peft = export_peft_state_dict(named)
    assert import_peft_state_dict(peft).keys() == named.keys()
    renamed = {k.replace(".lora_A.weight", ".lora_A.default.weight").replace(".lora_B.weight", ".lora_B.default.weight"): v for k, v in peft.items()}
    assert import_peft_state_dict(renamed).keys() == named.keys()
    d = tmp_path / "adapter"
    d.mkdir()
    save_file({k: v.contiguous() for k, v in renamed.items()}, str(d / "adapter_model.safetensors"))
    (d / "adapter_config.json").write_text(json.dumps(peft_adapter_config(cfg, lc)))
    dst = Zero3Engine(cfg, "cpu", seed=2, lora=lc)
    assert load_peft_adapter(dst, str(d)) == len(named)
    got = {n: t for n, t in dst.named_full_tensors() if ".lora." in n}
    assert all(torch.allclose(got[n].float(), named[n].float(), atol=1e-6) for n in named)
    (d / "adapter_config.json").write_text(json.dumps({**peft_adapter_config(cfg, lc), "r": 16}))
    with pytest.raises(ValueError, match="rank"):
        load_peft_adapter(dst, str(d))
