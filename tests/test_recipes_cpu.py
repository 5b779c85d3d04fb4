"""Recipe trainers on CPU (in-process worker backend, tiny model): DAPO dynamic sampling re-generates until enough informative
groups exist, SPPO / SPIN / GSPO / entropy-mechanism run their losses end to end.  Reference: 3rdparty/verl/recipe/*."""
import numpy as np
import pytest
import torch

from polyrl_b200.config.schema import load_config
from polyrl_b200.data.synthetic import SyntheticPromptDataset
from polyrl_b200.models import get_model_config
from polyrl_b200.protocol import DataProto
from polyrl_b200.recipes import RECIPES, dapo, entropy, flowrl, gspo, prime, spin, sppo
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
    ("flowrl", flowrl, ["actor_rollout_ref.rollout.n=4"]),
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
    if name == "flowrl":
        assert {"actor/log_z", "actor/flowrl_residual", "actor/importance_weight", "actor/ref_kl"} <= set(m) and "timing_s/ref" in m


def test_flowrl_objective_and_head_learning(tmp_path):
    """Trajectory balance: zero loss exactly when log Z + mean log pi = beta r + mean log pi_ref; the head's gradient moves log Z
    towards the target; tokens outside the trust region leave the averages; the head survives a worker checkpoint."""
    from polyrl_b200.algos.flowrl import LogZHead, flowrl_objective

    torch.manual_seed(0)
    b, r = 4, 6
    lp = -torch.rand(b, r)
    ref = -torch.rand(b, r)
    adv = torch.randn(b, 1).expand(b, r).contiguous()
    mask = torch.ones(b, r)
    mask[0, 4:] = 0
    m = mask
    target = 2.0 * ((adv * m).sum(-1) / m.sum(-1)) + (ref * m).sum(-1) / m.sum(-1) - (lp * m).sum(-1) / m.sum(-1)
    loss, met = flowrl_objective(lp, ref, lp.clone(), target, adv, mask, 0.2, 0.2, beta=2.0)
    assert loss.item() < 1e-10 and abs(met["actor/importance_weight"] - 1.0) < 1e-6 and met["actor/cispo_mask_ratio"] == 0.0
    z = torch.zeros(b, requires_grad=True)
    loss, _ = flowrl_objective(lp, ref, lp.clone(), z, adv, mask, 0.2, 0.2, beta=2.0)
    loss.backward()
    assert torch.allclose(z.grad, -2 * target / b, atol=1e-6)          # d/dz mean((z - target)^2)
    old = lp.clone()
    old[1, 0] = lp[1, 0] - 1.0                                        # ratio e^1 > 1 + clip with a positive advantage -> dropped
    adv2 = adv.abs()
    _, met = flowrl_objective(lp, ref, old, target, adv2, mask, 0.2, 0.2, beta=2.0)
    assert met["actor/cispo_mask_ratio"] > 0
    head = LogZHead(16, num_layers=3, dropout=0.0)
    assert head(torch.randn(5, 16)).shape == (5,)
    # end to end through a worker: one update changes the head, a checkpoint restores it
    from polyrl_b200.config.schema import load_config, need_reference_policy
    from polyrl_b200.workers.actor import FlowRLActor
    from polyrl_b200.workers.hybrid_worker import ActorRolloutRefWorker

    cfg = load_config(overrides=flowrl.DEFAULTS + BASE + ["actor_rollout_ref.rollout.n=2", "data.train_batch_size=2",
                                                          "actor_rollout_ref.actor.ppo_mini_batch_size=2"])
    assert need_reference_policy(cfg)
    w = ActorRolloutRefWorker(cfg, role="actor_rollout_ref", device="cpu")
    w.init_model()
    assert isinstance(w.actor, FlowRLActor)
    before = {k: v.clone() for k, v in w.actor.log_z.state_dict().items()}
    from polyrl_b200.data.synthetic import SyntheticPromptDataset
    from polyrl_b200.protocol import DataProto, collate_fn

    ds = SyntheticPromptDataset(w.model_cfg.vocab_size, 8, (8, 16), 32)
    batch = w.generate_sequences(DataProto.from_single_dict(collate_fn([ds[0], ds[1]])))
    batch.meta_info["temperature"] = 1.0
    batch.batch["old_log_probs"] = w.compute_log_prob(batch).batch["old_log_probs"]
    batch.batch["ref_log_prob"] = w.compute_ref_log_prob(batch).batch["ref_log_prob"]
    batch.batch["advantages"] = torch.randn(len(batch), 1).expand(-1, batch.batch["responses"].shape[1]) * batch.batch["response_mask"]
    batch.meta_info.update(is_opt_step=True, is_lr_step=True)
    w.update_actor_stream(batch)
    after = w.actor.log_z.state_dict()
    assert any(not torch.equal(after[k], before[k]) for k in before)
    w.save_checkpoint(str(tmp_path / "ck"), None, 1, None)
    w2 = ActorRolloutRefWorker(cfg, role="actor_rollout_ref", device="cpu")
    w2.init_model()
    w2.load_checkpoint(str(tmp_path / "ck"))
    assert all(torch.equal(w2.actor.log_z.state_dict()[k], after[k]) for k in after)


def test_prime_recipe_filters_by_accuracy_and_trains_the_prm():
    """PRIME: over-sampled generation, accuracy-band group filter, online PRM update before scoring, combined advantages."""
    b = DataProto.from_dict({"acc": torch.tensor([1.0, 1.0, 0.0, 1.0, 0.0, 0.0])}, non_tensors={"uid": np.array(list("aabbcc"), dtype=object)})
    kept, k, g = prime.filter_by_accuracy(b, 0.2, 0.8)
    assert (k, g) == (1, 3) and list(kept.non_tensor_batch["uid"]) == ["b", "b"]

    def reward_fn(batch):                       # a learnable-looking signal: 1 when the first response token is even
        return (batch.batch["responses"][:, 0] % 2 == 0).float()

    tr = _trainer(prime.PRIMETrainer, prime.DEFAULTS, ["actor_rollout_ref.rollout.n=4", "data.train_batch_size=2", "data.oversample_factor=2",
                                                       "actor_rollout_ref.actor.ppo_mini_batch_size=2", "reward_model.model.optim.lr=1e-3"],
                  reward_fn)
    try:
        assert tr.prm_params > 0
        before = tr.wg.workers[0].prm.engine.named_full_tensors()
        before = {n: t.clone() for n, t in before}
        hist = tr.fit(max_steps=1)
        after = dict(tr.wg.workers[0].prm.engine.named_full_tensors())
    finally:
        tr.shutdown()
    m = hist[0]
    assert m["num_samples"] in (4, 8) and m["actor/grad_norm"] > 0 and 0.0 < m["prime/kept_group_ratio"] <= 1.0
    assert {"reward_model/prm_ce_loss", "reward_model/pairwise_acc", "reward_model/process_reward_mean", "prime/acc_mean"} <= set(m)
    assert 0.2 <= m["prime/acc_mean"] <= 0.8                                            # only mixed groups were trained on
    assert any(not torch.equal(after[n], before[n]) for n in before)                    # the PRM itself was updated


def test_collabllm_multiturn_aware_reward_from_simulated_futures():
    """CollabLLM agent loop: the trained response is ONE assistant turn; ``num_repeat_rollouts`` futures are sampled with a simulated
    user (one more user turn each here, then the termination signal) inside the same engine; the reward of the trained turn is the
    weighted mean of the conversation-level metrics of its futures and lands on the last response token as ``rm_scores``."""
    from polyrl_b200.agent.agent_loop import get_agent_loop_cls, outputs_to_dataproto
    from polyrl_b200.agent.interactions import load_interactions
    from polyrl_b200.data.rl_dataset import load_tokenizer
    from polyrl_b200.models import TransformerWeights
    from polyrl_b200.recipes import collabllm as C
    from polyrl_b200.rollout.engine import RolloutEngine

    cfg = get_model_config("tiny-llama")
    w = TransformerWeights(cfg, "cpu", torch.float32, seed=3)
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=16, max_model_len=512, kv_cache_tokens=16 * 512, enable_prefix_cache=True)
    eng.set_weights(w.embed, w.layers, w.head)
    tok = load_tokenizer("byte")
    calls = []

    def user(messages, **kw):                       # the simulator sees the dialogue from the user's side + its hidden goal
        assert messages[0]["role"] == "system" and "find x" in messages[0]["content"] and messages[-1]["role"] == "user"
        calls.append(len(messages))
        mine = sum(m["role"] == "assistant" for m in messages)             # its own earlier lines
        return "can you explain?" if mine < 2 else f"thanks {C.TERMINATION_SIGNAL}"

    its = load_interactions({"interaction": [{"name": "collabllm", "class_name": "collabllm", "config": {"user_model": user}}]})
    loop = get_agent_loop_cls("collabllm_agent")(eng, tok, interactions=its, num_repeat_rollouts=2, response_length=16,
                                                 metric_weights={"token_amount": -0.01, "interactivity": 1.0, "bleu_score": 0.5},
                                                 max_assistant_turns=4, max_user_turns=3)
    ik = {"name": "collabllm", "single_turn_prompt": "find x", "ground_truth": "x = 4"}
    samples = [{"messages": [{"role": "user", "content": f"help {i}"}], "interaction_kwargs": ik} for i in range(3)]
    outs = loop.generate(samples, {"temperature": 1.0})
    assert len(outs) == 3 and all(o.response_mask == [1] * len(o.response_ids) and 0 < len(o.response_ids) <= 16 for o in outs)
    m = C.ConversationMetrics(tok)
    for o in outs:
        assert o.metrics["collabllm/num_futures"] == 2 and o.reward_score is not None
        want = -0.01 * o.metrics["collabllm/token_amount"] + o.metrics["collabllm/interactivity"] + 0.5 * o.metrics["collabllm/bleu_score"]
        assert abs(o.reward_score - want) < 1e-9
        assert o.metrics["collabllm/token_amount"] > 0                       # futures hold at least the trained turn + one more
    assert len(calls) >= 6                                                   # >= one simulated reply per future
    dp = outputs_to_dataproto(outs, 64, 16, cfg.pad_token_id)
    rm, mask = dp.batch["rm_scores"], dp.batch["attention_mask"][:, 64:]
    last = mask.sum(-1) - 1
    assert torch.allclose(rm[torch.arange(3), last], torch.tensor([o.reward_score for o in outs], dtype=torch.float32))
    assert (rm != 0).sum() <= 3 and "collabllm/interactivity" in dp.non_tensor_batch
    # metrics
    assert C.bleu_score("the answer is x = 4", "the answer is x = 4") == pytest.approx(1.0)
    assert 0 < C.bleu_score("the answer is x = 5 maybe", "the answer is x = 4") < 1 and C.bleu_score("", "a") == 0
    conv = [{"role": "user", "content": "q"}, {"role": "assistant", "content": "which one?"}, {"role": "user", "content": "a"},
            {"role": "assistant", "content": "x = 4"}]
    assert m.interactivity(conv, {}) == 1.0 and m.bleu_score(conv, {"ground_truth": "x = 4"}) == pytest.approx(1.0)
    judged = C.ConversationMetrics(tok, judge=lambda msgs, **kw: '{"accuracy": 1}' if "x = 4" in msgs[0]["content"] else '{"accuracy": 0}')
    assert judged.accuracy(conv, {"ground_truth": "x = 4", "single_turn_prompt": "find x"}) == 1.0
    r, per = C.multiturn_aware_reward([conv, conv[:2]], {"ground_truth": "x = 4"}, {"interactivity": 2.0}, m)
    assert per["interactivity"] == 0.5 and r == 1.0                           # second future: only one assistant turn -> 0
    rows = C.build_rl_rows([{"single_turn_prompt": "find x", "single_turn_completion": "x = 4", "chat": conv}])
    assert len(rows) == 2 and rows[1]["prompt"] == conv[:3] and rows[0]["extra_info"]["interaction_kwargs"]["name"] == "collabllm"
