"""Recipe trainers on CPU (in-process worker backend, tiny model): DAPO dynamic sampling re-generates until enough informative
groups exist, SPPO / SPIN / GSPO / entropy-mechanism run their losses end to end.  Reference: 3rdparty/verl/recipe/*."""
import numpy as np
import pytest
import torch

from polyrl_b200.config.schema import load_config
from polyrl_b200.data.synthetic import SyntheticPromptDataset
from polyrl_b200.models import get_model_config
from polyrl_b200.protocol import DataProto
from polyrl_b200.recipes import RECIPES, dapo, entropy, gspo, spin, sppo
from polyrl_b200.trainer.controller_ppo import make_worker_group

BASE = ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-llama", "data.max_prompt_length=32", "data.max_response_length=10",
        "actor_rollout_ref.actor.use_dynamic_bsz=True", "actor_rollout_ref.actor.ppo_max_token_len_per_gpu=256",
        "actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True", "actor_rollout_ref.rollout.log_prob_max_token_len_per_gpu=256",
        "actor_rollout_ref.rollout.max_num_seqs=32", "actor_rollout_ref.rollout.cuda_graph=False", "trainer.n_gpus_per_node=1",
        "reward_model.reward_manager=naive"]


def _trainer(cls, defaults, extra, reward_fn=None):
    cfg = load_config(overrides=defaults + BASE + extra)
    mc = get_model_config("tiny-llama")
    ds = SyntheticPromptDataset(mc.vocab_size, 64, (8, 24), 32)
    return cls(cfg, make_worker_group(cfg, "local", 1), ds, None, reward_fn)


def test_keep_informative_groups_and_win_probability():
    b = DataProto.from_dict({"token_level_scores": torch.tensor([[1.0], [1.0], [0.0], [1.0], [0.5], [0.5]]),
                             "response_mask": torch.ones(6, 1)},
                            non_tensors={"uid": np.array(["a", "a", "b", "b", "c", "c"], dtype=object)})
    kept, k, g = dapo.keep_informative_groups(b, "seq_reward")
    assert (k, g) == (1, 3) and list(kept.non_tensor_batch["uid"]) == ["b", "b"]
    p = sppo.win_probability(torch.tensor([1.0, 1.0, 0.0, 2.0, 1.0]), ["a", "a", "b", "b", "b"])
    assert p.tolist() == [0.5, 0.5, 0.0, 1.0, 0.5]


def test_dapo_dynamic_sampling_regenerates():
    calls = {"n": 0}

    def reward_fn(batch):                      # every other prompt group gets a constant reward -> uninformative, must be dropped
        calls["n"] += 1
        resp = batch.batch["responses"]
        uid = batch.non_tensor_batch["uid"]
        order = {u: i for i, u in enumerate(dict.fromkeys(uid))}
        s = torch.tensor([float(int(resp[i, 0]) % 2) if order[uid[i]] % 2 == 0 else 0.25 for i in range(len(batch))])
        return s

    tr = _trainer(dapo.DAPOTrainer, dapo.DEFAULTS, ["actor_rollout_ref.rollout.n=4", "data.train_batch_size=4", "data.gen_batch_size=4",
                                                    "actor_rollout_ref.actor.ppo_mini_batch_size=4"], reward_fn)
    try:
        hist = tr.fit(max_steps=1)
    finally:
        tr.shutdown()
    m = hist[0]
    assert m["num_samples"] == 16                       # exactly train_batch_size groups x n after alignment
    assert m["dapo/num_gen_batches"] >= 2 and calls["n"] == m["dapo/num_gen_batches"]
    assert m["dapo/kept_group_ratio"] <= 0.5 + 1e-6 and m["actor/grad_norm"] > 0


def test_dapo_gives_up_after_max_gen_batches():
    tr = _trainer(dapo.DAPOTrainer, dapo.DEFAULTS, ["actor_rollout_ref.rollout.n=2", "data.train_batch_size=2", "data.gen_batch_size=2",
                                                    "actor_rollout_ref.actor.ppo_mini_batch_size=2",
                                                    "algorithm.filter_groups.max_num_gen_batches=2"],
                  lambda b: torch.zeros(len(b)))
    try:
        with pytest.raises(RuntimeError, match="dynamic sampling"):
            tr.fit(max_steps=1)
    finally:
        tr.shutdown()


@pytest.mark.parametrize("name,mod,extra", [
    ("sppo", sppo, ["actor_rollout_ref.rollout.n=4"]),
    ("spin", spin, []),
    ("gspo", gspo, ["actor_rollout_ref.rollout.n=4"]),
    ("entropy", entropy, ["actor_rollout_ref.rollout.n=4", "actor_rollout_ref.actor.policy_loss.loss_mode=kl_cov", "data.gen_batch_size=4"]),
])
def test_recipe_runs_one_step(name, mod, extra):
    tr = _trainer(RECIPES[name], mod.DEFAULTS, ["data.train_batch_size=4", "actor_rollout_ref.actor.ppo_mini_batch_size=4", *extra])
    try:
        hist = tr.fit(max_steps=1)
    finally:
        tr.shutdown()
    m = hist[0]
    assert m["num_samples"] in (8, 16) and np.isfinite(m["actor/grad_norm"]) and m["actor/grad_norm"] > 0, m
    if name == "sppo":
        assert 0.0 <= m["sppo/win_prob_mean"] <= 1.0 and "actor/sppo_log_ratio" in m
    if name == "spin":
        assert "spin/pair_margin" in m
