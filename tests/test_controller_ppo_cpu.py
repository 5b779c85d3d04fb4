"""Single-controller GRPO on CPU: one driver process, two worker processes (gloo) behind MPWorkerGroup, every phase dispatched
through the @register protocol (reference: RayPPOTrainer.fit over RayWorkerGroup, 3rdparty/verl/verl/trainer/ppo/ray_trainer.py:968).
The in-process backend is also checked against the SPMD colocated trainer fed the same prompts (same weights after one step)."""
import pytest
import torch

from polyrl_b200.config.schema import load_config
from polyrl_b200.data.synthetic import SyntheticPromptDataset
from polyrl_b200.models import get_model_config
from polyrl_b200.trainer.controller_ppo import ControllerPPOTrainer, make_worker_group

OV = ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "actor_rollout_ref.rollout.n=4",
      "data.max_prompt_length=32", "data.max_response_length=12", "data.train_batch_size=4", "actor_rollout_ref.actor.ppo_mini_batch_size=4",
      "actor_rollout_ref.actor.use_dynamic_bsz=True", "actor_rollout_ref.actor.ppo_max_token_len_per_gpu=256",
      "actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True", "actor_rollout_ref.rollout.log_prob_max_token_len_per_gpu=256",
      "actor_rollout_ref.rollout.max_num_seqs=16", "actor_rollout_ref.rollout.cuda_graph=False", "trainer.total_epochs=1"]


def _run(backend: str, world: int, steps: int, tmp_path=None, extra=()):
    cfg = load_config(overrides=OV + [f"trainer.n_gpus_per_node={world}", *extra])
    mc = get_model_config("tiny-llama")
    ds = SyntheticPromptDataset(mc.vocab_size, 32, (8, 24), 32)
    wg = make_worker_group(cfg, backend, world)
    tr = ControllerPPOTrainer(cfg, wg, ds, val_dataset=SyntheticPromptDataset(mc.vocab_size, 4, (8, 24), 32, seed=3))
    try:
        hist = tr.fit(max_steps=steps)
        val = tr.validate()
        if tmp_path is not None:
            wg.save_checkpoint(str(tmp_path / "ck"), tr.global_step, None)
        return hist, val, wg.weights_checksum()
    finally:
        tr.shutdown()


def test_controller_grpo_two_worker_processes(tmp_path):
    hist, val, ck = _run("mp", 2, 2, tmp_path)
    assert len(hist) == 2
    for m in hist:
        assert m["num_samples"] == 16 and m["actor/grad_norm"] > 0 and m["actor/grad_norm"] == m["actor/grad_norm"]
        assert {"timing_s/gen", "timing_s/old_log_prob", "timing_s/update_actor", "timing_s/update_weight"} <= set(m)
    assert "val-core/score/mean@1" in val and ck > 0
    assert (tmp_path / "ck" / "actor").exists()


def test_controller_in_process_backend_with_critic():
    hist, _, _ = _run("local", 1, 1, extra=["algorithm.adv_estimator=gae", "critic.model.path=tiny-llama",
                                            "critic.ppo_mini_batch_size=4", "critic.use_dynamic_bsz=True",
                                            "critic.ppo_max_token_len_per_gpu=256"])
    assert hist[0]["num_samples"] == 16 and "critic/vf_loss" in hist[0]


def test_controller_matches_spmd_colocated_step():
    from polyrl_b200.protocol import DataProto, collate_fn
    from polyrl_b200.trainer.colocated import ColocatedPPO

    _, _, ck = _run("local", 1, 1)
    cfg = load_config(overrides=OV + ["trainer.n_gpus_per_node=1"])
    mc = get_model_config("tiny-llama")
    ds = SyntheticPromptDataset(mc.vocab_size, 32, (8, 24), 32)
    ppo = ColocatedPPO(cfg, device="cpu")
    ppo.step(DataProto.from_single_dict(collate_fn([ds[i] for i in range(4)])))
    ref = sum(float(t.double().abs().sum()) for _, t in ppo.worker.actor_engine.named_full_tensors())
    assert abs(ck - ref) <= 1e-6 * ref, (ck, ref)


PP2 = ["actor_rollout_ref.actor.megatron.pipeline_model_parallel_size=2"]


@pytest.mark.parametrize("world,extra", [(2, PP2), (4, PP2 + ["actor_rollout_ref.actor.megatron.virtual_pipeline_model_parallel_size=2",
                                                             "actor_rollout_ref.model.override_config={num_layers: 4}"]),
                                         (2, ["actor_rollout_ref.actor.megatron.tensor_model_parallel_size=2"]),
                                         (2, PP2 + ["actor_rollout_ref.actor.use_kl_loss=True", "actor_rollout_ref.actor.kl_loss_coef=0.01",
                                                    "actor_rollout_ref.actor.entropy_coeff=0.001", "actor_rollout_ref.actor.optim.lr=1e-3"]),
                                         (2, ["actor_rollout_ref.actor.megatron.expert_model_parallel_size=2",
                                              "actor_rollout_ref.model.path=tiny-qwen3-moe"])])
def test_controller_drives_a_pipeline_parallel_actor(world, extra):
    """actor.strategy=megatron: GRPO through the single controller with a pipeline-parallel actor (pp=2; 4 workers = 2 pipelines
    with interleaved virtual stages), data-parallel rollout engines refreshed from the gathered pipeline state.  The log-probs the
    pipeline computes for the sampled responses must equal the rollout engines' own (same weights on both sides)."""
    cfg = load_config(overrides=OV + [f"trainer.n_gpus_per_node={world}", "actor_rollout_ref.actor.strategy=megatron",
                                      "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=4", *extra])
    tp, ep = "tensor_model_parallel_size=2" in extra[0], "expert_model_parallel_size=2" in extra[0]
    mc = get_model_config(cfg.actor_rollout_ref.model.path, **cfg.actor_rollout_ref.model.override_config)
    ds = SyntheticPromptDataset(mc.vocab_size, 32, (8, 24), 32)
    wg = make_worker_group(cfg, "mp", world)
    tr = ControllerPPOTrainer(cfg, wg, ds)
    try:
        assert (tr.model_info["ep"] if ep else tr.model_info["tp"] if tp else tr.model_info["pp"]) == 2
        assert tr.model_info["dp"] == (world if ep else world // 2)
        assert tr.model_info["stage_params"] < mc.num_params()
        from polyrl_b200.protocol import DataProto, collate_fn

        batch = wg.generate_sequences(DataProto.from_single_dict(collate_fn([ds[i] for i in range(4)])))
        lp = wg.compute_log_prob(batch).batch["old_log_probs"]
        mask = batch.batch["response_mask"].bool()
        assert (lp - batch.batch["rollout_log_probs"])[mask].abs().max() < 1e-3           # pipeline forward == engine decode
        before = wg.weights_checksum()
        hist = tr.fit(max_steps=2)
        assert len(hist) == 2 and all(m["num_samples"] == 16 and m["actor/grad_norm"] > 0 for m in hist)
        assert wg.weights_checksum() != before                                            # the update reached the rollout engines
        if any("use_kl_loss=True" in e for e in extra):                                   # reference policy + KL / entropy terms
            assert all("actor/kl_loss" in m and "actor/entropy" in m and "timing_s/ref" in m for m in hist)
            assert hist[0]["actor/kl_loss"] < 1e-7 < hist[1]["actor/kl_loss"]             # identical at step 1, drifting after the update
    finally:
        tr.shutdown()


def test_bypass_mode_uses_engine_log_probs_as_old_policy():
    """algorithm.rollout_correction.bypass_mode (= async_training.use_rollout_log_probs): no recompute pass; at step 1 the ratio is
    the engine / trainer mismatch only, so the clip fraction is zero and ppo_kl tiny."""
    hist, _, _ = _run("local", 1, 1, extra=["async_training.use_rollout_log_probs=True"])
    m = hist[0]
    assert m["actor/grad_norm"] > 0 and m["actor/entropy_old"] != m["actor/entropy_old"]          # NaN: no entropy pass was run
    assert m["actor/pg_clipfrac"] == 0.0 and abs(m["actor/ppo_kl"]) < 1e-3


def test_model_based_reward_through_the_controller():
    """reward_model.enable: the workers' value-head reward model scores the rollouts (no rule-based scorer), GRPO trains on it."""
    hist, _, _ = _run("mp", 2, 1, extra=["reward_model.enable=True", "reward_model.model.path=tiny-llama", "reward_model.micro_batch_size_per_gpu=4"])
    m = hist[0]
    assert m["num_samples"] == 16 and m["actor/grad_norm"] > 0 and m["critic/score/mean"] != 0.0


def test_skip_rollout_replays_dumped_generations(tmp_path):
    """rollout.skip_rollout: the first run dumps its generations, a second run with the same dump dir replays them (identical
    responses, the engine is not asked to generate)."""
    extra = ["actor_rollout_ref.rollout.skip_rollout=True", f"actor_rollout_ref.rollout.skip_dump_dir={tmp_path}/dump"]
    h1, _, ck1 = _run("local", 1, 1, extra=extra)
    assert list((tmp_path / "dump" / "rank0").glob("*_step0.pt"))
    h2, _, ck2 = _run("local", 1, 1, extra=extra + ["data.seed=1"])
    assert h1[0]["response_length/mean"] == h2[0]["response_length/mean"] and abs(ck1 - ck2) <= 1e-9 * ck1
