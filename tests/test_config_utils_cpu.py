"""Config system, seqlen balancing, checkpoint manager, scorers, metrics (CPU)."""
import json

import numpy as np
import pytest
import torch

from polyrl_b200.config.schema import PPOConfig, apply_overrides, load_config, validate_config, need_critic
from polyrl_b200.data.reward_score import default_compute_score, gsm8k, math as math_score
from polyrl_b200.utils.seqlen_balancing import get_seqlen_balanced_partitions, rearrange_micro_batches


def test_overrides_yaml_and_validation(tmp_path):
    y = tmp_path / "c.yaml"
    y.write_text("algorithm:\n  adv_estimator: grpo\nactor_rollout_ref:\n  actor:\n    optim:\n      lr: 3.0e-6\n")
    cfg = load_config(str(y), ["actor_rollout_ref.rollout.n=8", "data.train_batch_size=16", "trainer.logger=[console,file]",
                               "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=2", "actor_rollout_ref.rollout.name=sglang-disaggregated",
                               "actor_rollout_ref.actor.clip_ratio_high=0.28", "trainer.n_gpus_per_node=2",
                               "+actor_rollout_ref.model.override_config.num_layers=4"])
    assert cfg.algorithm.adv_estimator == "grpo" and cfg.actor_rollout_ref.actor.optim.lr == 3.0e-6
    assert cfg.actor_rollout_ref.rollout.n == 8 and cfg.trainer.logger == ["console", "file"]
    assert cfg.actor_rollout_ref.rollout.name == "native-disaggregated"
    assert cfg.actor_rollout_ref.model.override_config == {"num_layers": 4}
    assert not need_critic(cfg)
    validate_config(cfg)
    with pytest.raises(KeyError):
        apply_overrides(cfg, ["actor_rollout_ref.actor.nonexistent=1"])
    cfg.actor_rollout_ref.rollout.min_stream_batch_size = 12          # not divisible by n=8
    with pytest.raises(ValueError):
        validate_config(cfg)
    with pytest.raises(NotImplementedError):
        apply_overrides(PPOConfig(), ["actor_rollout_ref.rollout.pipeline_model_parallel_size=2"])
    assert cfg["trainer"]["total_epochs"] == 1 and "data" in cfg


def test_seqlen_balancing():
    rng = np.random.default_rng(0)
    lens = rng.integers(10, 2000, size=64).tolist()
    parts = get_seqlen_balanced_partitions(lens, 8, equal_size=True)
    sums = [sum(lens[i] for i in p) for p in parts]
    assert all(len(p) == 8 for p in parts) and sorted(i for p in parts for i in p) == list(range(64))
    assert max(sums) - min(sums) < 0.05 * np.mean(sums)
    mbs = rearrange_micro_batches(lens, 8192)
    assert all(sum(lens[i] for i in p) <= 8192 for p in mbs) and sorted(i for p in mbs for i in p) == list(range(64))
    with pytest.raises(ValueError):
        rearrange_micro_batches([100, 9000], 8192)


def test_scorers():
    assert gsm8k.compute_score("so the answer is #### 1,234", "1234") == 1.0
    assert gsm8k.compute_score("#### 12 then #### 13", "12") == 0.0
    assert gsm8k.compute_score("no answer", "1") == 0.0
    assert math_score.compute_score("thus \\boxed{\\frac{1}{2}}", "0.5") == 1.0
    assert math_score.compute_score("\\boxed{3}", "4") == 0.0
    assert default_compute_score("openai/gsm8k", "#### 7", "7") == 1.0
    assert default_compute_score("open-r1/OpenR1-Math-220k", "\\boxed{x+1}", "x+1") == 1.0
    with pytest.raises(NotImplementedError):
        default_compute_score("unknown/ds", "", "")
    f = default_compute_score
    assert f("math_dapo", "blah\nAnswer: $\\frac{1}{2}$", "\\frac12") == 1.0 and f("aime24", "no answer here", "3") == -1.0
    assert f("numina_cn_k12", "so we get \\boxed{0.5}", "1/2") == 1.0 and f("numina_olympiads", "the result is 3.14159", "3.1416") == 1.0
    assert abs(f("hiyouga/geometry3k", "<think>x</think> so \\boxed{12}", "12") - 1.0) < 1e-9
    assert abs(f("hiyouga/geometry3k", "\\boxed{12}", "12") - 0.9) < 1e-9
    assert f("searchR1_nq", "<answer>The Eiffel Tower!</answer>", {"target": ["eiffel tower"]}) == 1.0
    assert f("searchR1_nq", "<answer>London</answer>", {"target": ["Paris"]}) == 0.0


def test_checkpoint_roundtrip_and_rotation(tmp_path):
    from polyrl_b200.models import get_model_config, pack_padded
    from polyrl_b200.parallel.zero3 import Zero3Engine
    from polyrl_b200.utils.checkpoint import ShardedCheckpointManager, find_latest_ckpt_path, write_latest

    cfg = get_model_config("tiny-llama")
    eng = Zero3Engine(cfg, "cpu", lr=1e-2)
    ids = torch.randint(3, cfg.vocab_size, (2, 8))
    inp, _ = pack_padded(ids, torch.ones_like(ids))
    labels = torch.randint(0, cfg.vocab_size, (16,))
    lp, _ = eng.forward_logprob_entropy(inp, labels)
    (-lp.mean()).backward()
    eng.step()
    mgr = ShardedCheckpointManager(eng, cfg)
    for s in (1, 2, 3):
        mgr.save(str(tmp_path / f"global_step_{s}" / "actor"), s, max_ckpt_to_keep=2, extra={"lr_step": s})
        write_latest(tmp_path, s)
    assert not (tmp_path / "global_step_1" / "actor").exists() and (tmp_path / "global_step_3" / "actor" / "config.json").exists()
    assert find_latest_ckpt_path(tmp_path).endswith("global_step_3")
    want = [u.master.clone() for u in eng.units]
    lp2, _ = eng.forward_logprob_entropy(inp, labels)
    (-lp2.mean()).backward()
    eng.step()
    assert not torch.equal(eng.units[1].master, want[1])
    extra = ShardedCheckpointManager(eng, cfg).load(str(tmp_path / "global_step_3" / "actor"))
    assert extra == {"lr_step": 3} and eng.opt_step == 1
    assert all(torch.equal(u.master, w) for u, w in zip(eng.units, want))


def test_metrics_and_reward_manager():
    from polyrl_b200.data.reward_manager import get_reward_manager_cls
    from polyrl_b200.protocol import DataProto
    from polyrl_b200.utils.metrics import compute_data_metrics, process_validation_metrics, reduce_metrics

    b, p, r = 4, 3, 5
    mask = torch.ones(b, p + r, dtype=torch.long)
    mask[1, -2:] = 0
    d = DataProto.from_dict({"responses": torch.randint(0, 9, (b, r)), "attention_mask": mask, "response_mask": mask[:, -r:],
                             "token_level_scores": torch.rand(b, r), "token_level_rewards": torch.rand(b, r),
                             "advantages": torch.randn(b, r), "returns": torch.randn(b, r), "values": torch.randn(b, r)},
                            {"data_source": ["synthetic/x"] * b, "reward_model": [{"ground_truth": "1"}] * b})
    m = compute_data_metrics(d, use_critic=True)
    assert m["response_length/max"] == 5 and m["response_length/min"] == 3 and "critic/vf_explained_var" in m
    assert reduce_metrics({"a": [1.0, 3.0], "x/max": [1, 5]}) == {"a": 2.0, "x/max": 5.0}
    rm = get_reward_manager_cls("naive")(tokenizer=None)
    t = rm(d)
    assert t.shape == (b, r) and (t != 0).sum(-1).max() <= 1
    dapo = get_reward_manager_cls("dapo")(tokenizer=None, max_resp_len=5, overlong_buffer_len=2, penalty_factor=1.0)
    assert dapo(d).shape == (b, r)
    v = process_validation_metrics(["a", "a", "a", "a"], ["u1", "u1", "u2", "u2"], {"reward": [1.0, 0.0, 1.0, 1.0]})
    assert v["a"]["reward"]["mean@2"] == pytest.approx(0.75)


def test_memory_utils(tmp_path):
    import time

    from polyrl_b200.utils.memory import MemorySnapshotSampler, aggressive_empty_cache, log_gpu_memory_usage

    assert "allocated_gb" in aggressive_empty_cache()
    s = MemorySnapshotSampler(str(tmp_path), interval_s=0.02).start()
    time.sleep(0.1)
    s.stop()
    assert s.dump("t").exists() and len(s.samples) >= 2
    assert "free_gb" in log_gpu_memory_usage("test", rank=None)


def test_hbm_sizing_planner():
    from polyrl_b200.utils.hbm_sizing import count_params, plan
    from polyrl_b200.models import get_model_config

    assert abs(count_params(get_model_config("llama3-8b")) / 1e9 - 8.03) < 0.01
    p1 = plan("llama3-8b", n_gpus=1, max_num_seqs=128, max_model_len=768)
    assert p1.decisions["fits_training"] and 0 < p1.saved_blocks < 32 and p1.resident_replica == 0
    p8 = plan("llama3-8b", n_gpus=8, max_num_seqs=128, max_model_len=768)
    assert p8.saved_blocks == 32 and p8.resident_replica > 14 and p8.decisions["fits_training"]
    p70 = plan("llama3-70b", n_gpus=8, rollout_tp=8, critic="llama3-8b", max_model_len=4096)
    d = p70.decisions
    assert p70.resident_replica == 0 and p70.saved_blocks == 0 and d["optimizer_offload_during_rollout"]
    assert d["release_rollout_weights_during_training"] and d["fits_rollout"] and d["fits_training"]
    assert d["max_tokens_per_microbatch"] < 16384


def test_reference_launch_script_keys_are_accepted():
    """Every dotted override used by the reference stack's example / recipe launch scripts (PPO tree; collected from
    3rdparty/verl/examples/*/*.sh and recipe/*/*.sh) must parse, and the ones with a counterpart must take effect."""
    from polyrl_b200.config.schema import load_config

    keys = """actor_rollout_ref.actor.entropy_checkpointing=True actor_rollout_ref.actor.entropy_from_logits_with_chunking=True
    actor_rollout_ref.actor.fsdp_config.mixed_precision.param_dtype=bf16 actor_rollout_ref.actor.fsdp_config.use_orig_params=True
    actor_rollout_ref.actor.megatron.context_parallel_size=1 actor_rollout_ref.actor.megatron.expert_tensor_parallel_size=1
    actor_rollout_ref.actor.megatron.grad_offload=True actor_rollout_ref.actor.megatron.param_offload=True
    actor_rollout_ref.actor.megatron.override_transformer_config.recompute_granularity=full
    actor_rollout_ref.actor.megatron.override_transformer_config.moe_grouped_gemm=True
    actor_rollout_ref.actor.megatron.use_dist_checkpointing=True actor_rollout_ref.actor.megatron.use_mbridge=True
    actor_rollout_ref.actor.optim.lr_decay_style=constant actor_rollout_ref.actor.optim.override_optimizer_config.optimizer_offload_fraction=1
    actor_rollout_ref.actor.profiler.tool_config.npu.discrete=True actor_rollout_ref.actor.use_torch_compile=False
    actor_rollout_ref.model.exclude_modules=.*visual.* actor_rollout_ref.model.fused_kernel_options.impl_backend=torch
    actor_rollout_ref.model.lora_adapter_path=/x actor_rollout_ref.model.override_config.model_config.max_position_embeddings=4096
    actor_rollout_ref.model.trust_remote_code=True actor_rollout_ref.ref.entropy_checkpointing=True
    actor_rollout_ref.ref.log_prob_micro_batch_size=32 actor_rollout_ref.ref.megatron.pipeline_model_parallel_size=3
    actor_rollout_ref.ref.profiler.enable=True actor_rollout_ref.ref.ulysses_sequence_parallel_size=1
    actor_rollout_ref.rollout.engine_kwargs.sglang.attention_backend=triton actor_rollout_ref.rollout.layered_summon=True
    actor_rollout_ref.rollout.multi_stage_wake_up=True actor_rollout_ref.rollout.over_sample_rate=0.1
    actor_rollout_ref.rollout.trace.backend=mlflow actor_rollout_ref.rollout.update_weights_bucket_megabytes=512
    actor_rollout_ref.rollout.agent.agent_loop_config_path=/x actor_rollout_ref.rollout.multi_turn.num_repeat_rollouts=3
    algorithm.use_pf_ppo=True algorithm.pf_ppo.reweight_method=pow algorithm.pf_ppo.weight_pow=2.0
    critic.megatron.tensor_model_parallel_size=1 critic.model.trust_remote_code=True critic.ulysses_sequence_parallel_size=2
    data.apply_chat_template_kwargs.reasoning_effort=medium data.custom_cls.name=RLHFDataset data.custom_cls.path=x.py
    data.trust_remote_code=True data.filter_accuracy=True data.accuracy_lower_bound=0.2 data.oversample_factor=4
    global_profiler.global_tool_config.torch_memory.stack_depth=32 reward_model.forward_max_token_len_per_gpu=98304
    reward_model.megatron.tensor_model_parallel_size=4 reward_model.model.fsdp_config.param_offload=True
    reward_model.model.input_tokenizer=x reward_model.param_offload=False reward_model.profiler.enable=True
    reward_model.sandbox_fusion.max_concurrent=128 reward_model.use_dynamic_bsz=True reward_model.enable_resource_pool=True
    reward_model.rollout.name=sglang reward_model.model.beta_train=0.05 reward_model.model.optim.grad_clip=10.0
    reward_model.model.update=before reward_model.mini_batch_size=64 trainer.use_legacy_worker_impl=auto trainer.log_freq=1
    trainer.ref_update_freq=1 trainer.rl_logging_board_dir=/x trainer.num_data_controllers=1""".split()
    for kv in keys:
        load_config(overrides=[kv])
    c = load_config(overrides=["actor_rollout_ref.rollout.log_prob_micro_batch_size=64", "trainer.n_gpus_per_node=8", "rollout.n_gpus_per_node=4",
                               "rollout.nnodes=1", "actor_rollout_ref.rollout.enforce_eager=True", "critic.model.fsdp_config.param_offload=True",
                               "actor_rollout_ref.actor.loss_mode=kl_cov", "actor_rollout_ref.actor.clip_cov_ub=4",
                               "reward_model.reward_kwargs.overlong_buffer_cfg.enable=True", "reward_model.reward_kwargs.overlong_buffer_cfg.len=128",
                               "actor_rollout_ref.model.enable_activation_offload=True", "trainer.max_ckpt_to_keep=3",
                               "actor_rollout_ref.actor.optim.lr_scheduler_type=cosine", "actor_rollout_ref.actor.freeze_vision_tower=True",
                               "actor_rollout_ref.actor.optim.min_lr=1e-7", "actor_rollout_ref.actor.optim.lr=1e-6"])
    a = c.actor_rollout_ref
    assert c.trainer.n_rollout_gpus == 4 and a.rollout.log_prob_micro_batch_size_per_gpu == 16        # 64 over the 4 trainer GPUs
    assert a.rollout.cuda_graph is False and c.critic.fsdp_config.param_offload is True
    assert a.actor.policy_loss.loss_mode == "kl_cov" and a.actor.policy_loss.clip_cov_ub == 4.0
    assert c.reward_model.overlong_buffer.enable and c.reward_model.overlong_buffer.len == 128
    assert a.actor.fsdp_config.activation_offload and c.trainer.max_actor_ckpt_to_keep == 3 and c.trainer.max_critic_ckpt_to_keep == 3
    assert a.actor.optim.warmup_style == "cosine" and abs(a.actor.optim.min_lr_ratio - 0.1) < 1e-9 and a.model.freeze_vision_tower


def test_reference_sft_script_keys_are_accepted():
    """The SFT trainer takes the reference's sft_trainer / fsdp_sft_trainer override names."""
    from polyrl_b200.config.schema import apply_overrides
    from polyrl_b200.trainer.sft import SFTConfig, finalize_sft_config

    keys = """checkpoint.save_contents=[model,optimizer,extra] data.max_length=16384 data.max_token_len_per_gpu=65536 data.messages_key=messages
    data.micro_batch_size=64 data.multiturn.enable=true data.multiturn.tools_key=tools data.pad_mode=no_padding data.use_dynamic_bsz=True
    engine.fsdp_size=8 engine.strategy=fsdp engine.ulysses_sequence_parallel_size=2 engine.use_mbridge=False model.partial_pretrain=tiny-llama
    model.strategy=fsdp model.use_liger=False model.use_remove_padding=True optim.betas=[0.9,0.95] optim.clip_grad=1.0 optim.lr_decay_style=cosine
    optim.lr_warmup_init=0 optim.min_lr=2e-6 optim.weight_decay=0.1 trainer.experiment_name=x trainer.max_ckpt_to_keep=5 trainer.project_name=p
    trainer.resume_mode=auto trainer.test_freq=10 trainer.total_training_steps=50 ulysses_sequence_parallel_size=2 use_remove_padding=true""".split()
    for kv in keys:
        apply_overrides(SFTConfig(), [kv])
    c = SFTConfig()
    apply_overrides(c, ["model.partial_pretrain=tiny-llama", "data.multiturn.enable=true", "data.multiturn.messages_key=msgs", "data.micro_batch_size=16",
                        "trainer.total_training_steps=7", "trainer.test_freq=2", "engine.ulysses_sequence_parallel_size=2"])
    finalize_sft_config(c, world=4)
    assert (c.model.path, c.data.multiturn_messages_key, c.data.micro_batch_size_per_gpu) == ("tiny-llama", "msgs", 4)
    assert (c.trainer.total_steps, c.trainer.eval_freq, c.trainer.ulysses_sequence_parallel_size) == (7, 2, 2)
    import pytest

    bad = SFTConfig()
    apply_overrides(bad, ["engine.tensor_model_parallel_size=2"])
    with pytest.raises(NotImplementedError):
        finalize_sft_config(bad)


def test_hydra_style_yaml_files_load(tmp_path):
    """YAML files written for the reference stack (Hydra ``defaults`` / ``hydra`` blocks, ``_target_`` markers, ``${...}`` and
    ``oc.select`` interpolation, ``???`` mandatory markers, a few keys without counterpart) load with their values applied."""
    from polyrl_b200.config import schema as S

    y = tmp_path / "ppo_stream_trainer.yaml"
    y.write_text("""
hydra:
  searchpath: [file://somewhere]
defaults:
  - ppo_trainer
  - _self_
data:
  train_batch_size: 64
  max_prompt_length: 256
  tokenizer: null
actor_rollout_ref:
  model:
    path: ???
  actor:
    _target_: some.module.ActorConfig
    ppo_mini_batch_size: ${data.train_batch_size}
    optim: {lr: 2e-6, total_training_steps: "${oc.select:trainer.total_training_steps,-1}"}
    unknown_knob: 3
  rollout:
    _target_: x.RolloutConfig
    name: sglang-disaggregated
    n: 4
    prompt_length: ${data.max_prompt_length}
    rollout_manager: {port: 5001, endpoint: null}
    min_stream_batch_size: 8
trainer:
  total_training_steps: 12
  logger: ["console"]
""")
    c = S.load_config(str(y), overrides=["actor_rollout_ref.rollout.n=8"])
    a = c.actor_rollout_ref
    assert c.data.train_batch_size == 64 and a.actor.ppo_mini_batch_size == 64 and a.rollout.prompt_length == 256
    assert a.actor.optim.lr == 2e-6 and a.actor.optim.total_training_steps == 12 and c.trainer.total_training_steps == 12
    assert a.rollout.name == "native-disaggregated" and a.rollout.rollout_manager.port == 5001 and a.rollout.n == 8
    assert a.model.path == "tiny-llama"                                  # "???" keeps the default
    assert S._UNKNOWN_YAML_KEYS == ["actor_rollout_ref.actor.unknown_knob"]
    strict = tmp_path / "own.yaml"
    strict.write_text("data: {train_batch_size: 8, nonsense: 1}\n")
    import pytest

    with pytest.raises(KeyError):
        S.load_config(str(strict))                                       # the framework's own files stay strict
