"""Advantage estimators, losses, KL (spirit of 3rdparty/verl/tests/trainer/ppo/test_core_algos_on_cpu.py)."""
import numpy as np
import pytest
import torch

from polyrl_b200.algos import core_algos as A


def _mask(b=6, t=5):
    m = torch.ones(b, t)
    m[1, 3:] = 0
    m[4, 1:] = 0
    return m


def test_registry():
    assert A.get_adv_estimator_fn("grpo") is A.get_adv_estimator_fn(A.AdvantageEstimator.GRPO)
    with pytest.raises(ValueError):
        A.get_adv_estimator_fn("nope")
    for name in ["vanilla", "gspo", "gpg", "clip_cov", "kl_cov", "geo_mean", "rollout_correction"]:
        assert callable(A.get_policy_loss_fn(name))


def test_grpo_matches_loop_reference():
    torch.manual_seed(0)
    m = _mask()
    r = torch.zeros(6, 5)
    r[:, 0] = torch.tensor([1.0, 0.0, 1.0, 3.0, 2.0, 5.0])
    idx = np.array(["a", "a", "a", "b", "b", "c"], dtype=object)
    adv, ret = A.compute_grpo_outcome_advantage(r, m, idx)
    s = r.sum(-1)
    want = torch.zeros(6)
    for g in ["a", "b"]:
        sel = torch.tensor([i for i in range(6) if idx[i] == g])
        want[sel] = (s[sel] - s[sel].mean()) / (s[sel].std() + 1e-6)
    want[5] = s[5] / (1 + 1e-6)        # singleton group: mean 0, std 1
    assert torch.allclose(adv, want[:, None] * m, atol=1e-5) and torch.equal(adv, ret)
    adv2, _ = A.compute_grpo_outcome_advantage(r, m, idx, norm_adv_by_std_in_grpo=False)
    assert torch.allclose(adv2[0, 0], s[0] - s[:3].mean())


def test_rloo_opo_gpg_passk():
    m = torch.ones(4, 3)
    r = torch.zeros(4, 3)
    r[:, -1] = torch.tensor([1.0, 2.0, 6.0, 4.0])
    idx = np.array([0, 0, 0, 1])
    adv, _ = A.compute_rloo_outcome_advantage(r, m, idx)
    assert torch.allclose(adv[0, 0], torch.tensor(1.0 - (2.0 + 6.0) / 2))
    assert torch.allclose(adv[3, 0], torch.tensor(4.0))
    advv, _ = A.compute_rloo_vectorized_outcome_advantage(r, m, idx)
    assert torch.allclose(advv[:3], adv[:3]) and advv[3].abs().sum() == 0
    m2 = torch.tensor([[1.0, 1, 1], [1, 0, 0], [1, 1, 0], [1, 1, 1]])
    opo, _ = A.compute_opo_outcome_advantage(r, m2, idx)
    base = (3 * 1 + 1 * 2 + 2 * 6) / 6
    assert torch.allclose(opo[2, 0], torch.tensor(6.0 - base))
    pk, _ = A.compute_grpo_passk_outcome_advantage(r, m, np.array([0, 0, 0, 0]), config={"norm_adv_by_std_in_grpo": False})
    assert pk[2, 0] == 2.0 and pk[[0, 1, 3]].abs().sum() == 0
    gpg, _ = A.compute_gpg_outcome_advantage(r, m, idx)
    assert gpg.shape == (4, 3)


def test_gae_matches_manual_scan():
    torch.manual_seed(1)
    b, t = 3, 6
    r, v = torch.randn(b, t), torch.randn(b, t)
    m = torch.ones(b, t)
    m[1, 4:] = 0
    gamma, lam = 0.99, 0.95
    adv, ret = A.compute_gae_advantage_return(r, v, m, gamma, lam)
    raw = torch.zeros(b, t)
    for i in range(b):
        nv, last = 0.0, 0.0
        for k in reversed(range(t)):
            if m[i, k] == 0:
                raw[i, k] = last
                continue
            delta = r[i, k] + gamma * nv - v[i, k]
            last = delta + gamma * lam * last
            nv = v[i, k]
            raw[i, k] = last
    assert torch.allclose(ret, raw + v, atol=1e-5)
    assert torch.allclose(adv, A.masked_whiten(raw, m), atol=1e-5)


def test_reinforce_pp_and_remax():
    r = torch.tensor([[0.0, 0.0, 1.0, 0.0]])
    m = torch.tensor([[1.0, 1.0, 1.0, 0.0]])
    adv, ret = A.compute_reinforce_plus_plus_outcome_advantage(r, m, gamma=0.5)
    assert torch.allclose(ret[0, :3], torch.tensor([0.25, 0.5, 1.0]))
    adv2, ret2 = A.compute_remax_outcome_advantage(r, torch.tensor([0.3]), m)
    assert torch.allclose(ret2[0], torch.tensor([1.0, 1.0, 1.0, 0.0])) and torch.allclose(adv2[0, 0], torch.tensor(0.7))


def test_agg_loss_modes():
    x = torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    m = torch.tensor([[1.0, 1.0, 0.0], [1.0, 0.0, 0.0]])
    assert torch.isclose(A.agg_loss(x, m, "token-mean"), torch.tensor(7.0 / 3))
    assert torch.isclose(A.agg_loss(x, m, "seq-mean-token-sum"), torch.tensor((3.0 + 4.0) / 2))
    assert torch.isclose(A.agg_loss(x, m, "seq-mean-token-mean"), torch.tensor((1.5 + 4.0) / 2))
    assert torch.isclose(A.agg_loss(x, m, "seq-mean-token-sum-norm"), torch.tensor(7.0 / 3))
    with pytest.raises(ValueError):
        A.agg_loss(x, m, "bad")
    # fully masked rows (padded duplicates / filtered groups / vetoed sequences) do not dilute sequence means
    x2 = torch.cat([x, torch.full((2, x.shape[1]), 100.0)])
    m2 = torch.cat([m, torch.zeros(2, m.shape[1])])
    for mode in ("seq-mean-token-sum", "seq-mean-token-mean", "token-mean"):
        assert torch.isclose(A.agg_loss(x2, m2, mode), A.agg_loss(x, m, mode), atol=1e-6), mode


def test_vanilla_ppo_dual_clip():
    old = torch.zeros(1, 4)
    lp = torch.log(torch.tensor([[1.0, 1.5, 0.5, 5.0]]))
    adv = torch.tensor([[1.0, 1.0, -1.0, -1.0]])
    m = torch.ones(1, 4)
    loss, met = A.compute_policy_loss_vanilla(old, lp, adv, m, "token-mean", config={"clip_ratio": 0.2, "clip_ratio_c": 3.0})
    # token losses: -1*1 ; -1*1.2 (clipped) ; max(0.5, 0.8)=0.8 ; min(3, max(5,1.2))=3
    assert torch.isclose(loss, torch.tensor((-1.0 - 1.2 + 0.8 + 3.0) / 4))
    assert met["actor/pg_clipfrac"] == pytest.approx(0.5) and met["actor/pg_clipfrac_lower"] == pytest.approx(0.25)
    lp.requires_grad_(True)
    for name in ["gspo", "gpg", "clip_cov", "kl_cov", "geo_mean", "rollout_correction"]:
        l, _ = A.get_policy_loss_fn(name)(old, lp, adv, m, "token-mean", config={"clip_ratio": 0.2}, rollout_is_weights=None)
        l.backward()
        assert torch.isfinite(l) and torch.isfinite(lp.grad).all()
        lp.grad = None


def test_value_loss_and_kl():
    v, ret, old = torch.tensor([[1.0, 3.0]]), torch.tensor([[0.0, 0.0]]), torch.tensor([[0.5, 0.5]])
    loss, frac = A.compute_value_loss(v, ret, old, torch.ones(1, 2), 0.5)
    # clipped preds: 1.0, 1.0 -> max((1)^2,(1)^2)=1 ; max(9,1)=9 -> 0.5*mean = 2.5
    assert torch.isclose(loss, torch.tensor(2.5)) and frac == 0
    a, b = torch.tensor([0.1, -0.3]), torch.tensor([0.0, 0.2])
    assert torch.allclose(A.kl_penalty(a, b, "kl"), a - b)
    assert torch.allclose(A.kl_penalty(a, b, "mse"), 0.5 * (a - b) ** 2)
    d = b - a
    assert torch.allclose(A.kl_penalty(a, b, "low_var_kl"), torch.exp(d) - d - 1)
    a.requires_grad_(True)
    k = A.kl_penalty(a, b, "low_var_kl+")
    k.sum().backward()
    assert torch.allclose(a.grad, (a - b).detach())          # straight-through k2 gradient
    c = A.AdaptiveKLController(0.1, 0.05, 100)
    c.update(0.1, 10)
    assert c.value == pytest.approx(0.1 * (1 + 0.2 * 10 / 100))
    assert A.get_kl_controller({"type": "fixed", "kl_coef": 0.3}).value == 0.3


def test_dapo_filter_groups_masks_uninformative_prompts():
    """Dynamic sampling: groups whose samples all got the same score lose their loss mask; mixed groups are untouched."""
    import numpy as np

    from polyrl_b200.protocol import DataProto
    from polyrl_b200.trainer.core import filter_uninformative_groups

    mask = torch.ones(6, 4, dtype=torch.long)
    mask[1, 2:] = 0
    scores = torch.zeros(6, 4)
    scores[0, 3], scores[1, 1] = 1.0, 1.0          # group a: both correct -> filtered
    scores[2, 3] = 1.0                             # group b: 1 and 0 -> kept
    #                                                group c: both 0 -> filtered
    dp = DataProto.from_dict({"response_mask": mask, "token_level_scores": scores},
                             non_tensors={"uid": np.array(["a", "a", "b", "b", "c", "c"], dtype=object),
                                          "acc": np.array([1, 1, 1, 1, 0, 1], dtype=object)})
    out, m = filter_uninformative_groups(dp, "seq_reward")
    assert out.batch["response_mask"].sum(-1).tolist() == [0, 0, 4, 4, 0, 0]
    assert m == {"dapo/num_groups": 3.0, "dapo/filtered_groups": 2.0, "dapo/kept_group_ratio": 1.0 - 2 / 3}
    dp.batch["response_mask"] = mask
    out, m = filter_uninformative_groups(dp, "acc")          # per-sample metric from the non-tensor batch
    assert out.batch["response_mask"].sum(-1).tolist() == [0, 0, 0, 0, 4, 4] and m["dapo/filtered_groups"] == 2.0
    with pytest.raises(KeyError):
        filter_uninformative_groups(dp, "nope")


def test_cispo_sppo_online_dpo_losses():
    g = torch.Generator().manual_seed(0)
    b, t = 6, 5
    old = -torch.rand(b, t, generator=g)
    lp = (old + 0.3 * torch.randn(b, t, generator=g)).requires_grad_(True)
    adv = torch.randn(b, t, generator=g)
    mask = torch.ones(b, t)
    mask[2, 3:] = 0
    cfg = {"clip_ratio": 0.2, "policy_loss": {"sppo_eta": 2.0, "dpo_beta": 0.5}}
    # CISPO: gradient = -clip(ratio).detach() * A on EVERY unmasked token (also the clipped ones), token-mean
    loss, m = A.get_policy_loss_fn("cispo")(old, lp, adv, mask, "token-mean", cfg)
    (grad,) = torch.autograd.grad(loss, lp)
    ratio = torch.exp(lp.detach() - old)
    want = -(ratio.clamp(0.8, 1.2) * adv) * mask / mask.sum()
    assert torch.allclose(grad, want, atol=1e-6) and 0 < m["actor/pg_clipfrac"] < 1
    assert (grad[mask.bool()] != 0).all()
    # SPPO: squared distance between the sequence log-ratio and eta * reward
    seq_r = torch.tensor([1.0, -1.0, 0.5, 0.0, 1.0, -0.5])
    loss, _ = A.get_policy_loss_fn("sppo")(old, lp, seq_r[:, None].expand(b, t), mask, "token-mean", cfg)
    lr_ = ((lp.detach() - old) * mask).sum(-1)
    assert torch.allclose(loss, ((lr_ - 2.0 * seq_r) ** 2).mean(), atol=1e-6)
    # online DPO: pairs (0,1), (2,3), (4,5); the higher score is chosen; a tied pair is ignored
    score = torch.tensor([1.0, 0.0, 0.0, 1.0, 0.5, 0.5])
    assert A.pair_preferences(score).tolist() == [True, False, False, True, True, False]
    loss, m = A.get_policy_loss_fn("online_dpo")(old, lp, score[:, None].expand(b, t), mask, "token-mean", cfg)
    s_lp, s_old = (lp.detach() * mask).sum(-1), (old * mask).sum(-1)
    logit = lambda c, r: (s_lp[c] - s_lp[r]) - (s_old[c] - s_old[r])                       # noqa: E731
    want = (-torch.nn.functional.logsigmoid(0.5 * logit(0, 1)) - torch.nn.functional.logsigmoid(0.5 * logit(3, 2))) / 2
    assert torch.allclose(loss, want, atol=1e-6) and 0.0 <= m["actor/dpo_acc"] <= 1.0
    ipo = A.dpo_loss(torch.tensor([1.0]), torch.tensor([0.0]), torch.tensor([0.2]), torch.tensor([0.1]), beta=0.5, loss_type="ipo")
    assert torch.allclose(ipo, torch.tensor([(0.9 - 1.0) ** 2]))
    with pytest.raises(ValueError):
        A.pair_preferences(torch.zeros(3))


def test_prime_implicit_reward_blocks():
    from polyrl_b200.algos import prime as P

    g = torch.Generator().manual_seed(0)
    b, t, n = 4, 5, 2
    mask = torch.ones(b, t)
    mask[1, 3:] = 0
    prm, refp = -torch.rand(b, t, generator=g), -torch.rand(b, t, generator=g)
    q = P.implicit_process_rewards(prm, refp, mask, beta=0.5)
    assert torch.allclose(q, 0.5 * (prm - refp) * mask)
    whole = P.implicit_process_rewards(prm, refp, mask, beta=0.5, granularity="whole")
    assert torch.allclose(whole[1, 2], q[1].sum()) and whole[1].abs().sum() == whole[1, 2].abs() and torch.allclose(whole[0, 4], q[0].sum())
    acc = torch.tensor([1.0, 0.0, 0.0, 0.0])
    scores = (prm - refp)
    loss = P.prm_ce_loss(scores, acc, mask, beta=0.5)
    want = torch.nn.functional.binary_cross_entropy(torch.sigmoid(0.5 * (scores * mask).sum(-1)), acc)
    assert torch.allclose(loss, want, atol=1e-6)
    # leave-one-out: with outcome rewards only, pair (1, 0) -> +1*2-1 = 1 and 0*2-1 = -1 on the last tokens; a tied pair -> 0
    adv, ret = P.prime_advantages(mask, n, acc=acc, reward_gt_coef=1.0, reward_dpo_coef=0.0)
    assert ret[0].tolist() == [1.0] * 5 and ret[1].tolist() == [-1.0, -1.0, -1.0, 0.0, 0.0] and ret[2].abs().sum() == 0 and ret[3].abs().sum() == 0
    assert abs(float((adv * mask).sum() / mask.sum())) < 1e-5
    adv2, ret2 = P.prime_advantages(mask, n, acc=acc, process_rewards=q, reward_gt_coef=1.0, reward_dpo_coef=1.0)
    assert not torch.allclose(ret2, ret) and torch.isfinite(adv2).all()
    # ordering accuracy of the implicit reward
    s = torch.zeros(b, t)
    s[0, 0], s[1, 0] = 2.0, 1.0            # pair 0 ordered like acc (1 > 0); pair 1 tied labels -> ignored
    assert P.pairwise_accuracy(s, acc, mask, n) == 1.0 and P.pairwise_accuracy(-s, acc, mask, n) == 0.0
    assert P.pairwise_accuracy(s, torch.zeros(b), mask, n) == 0.5
    with pytest.raises(ValueError):
        P.prime_advantages(mask, 1, acc=acc)


def test_implicit_prm_learns_the_outcome_labels():
    """ImplicitPRM over ZeRO engines: starts at zero reward (PRM == reference), its CE loss drops when trained on outcome labels and
    the sequence reward then separates correct from wrong responses."""
    import numpy as np

    from polyrl_b200.algos.prime import ImplicitPRM, pairwise_accuracy
    from polyrl_b200.models import get_model_config
    from polyrl_b200.parallel.zero3 import Zero3Engine
    from polyrl_b200.protocol import DataProto

    cfg = get_model_config("tiny-llama")
    prm_e = Zero3Engine(cfg, "cpu", lr=3e-3, weight_decay=0.0, seed=2)
    ref_e = Zero3Engine(cfg, "cpu", trainable=False, seed=2)
    prm = ImplicitPRM(prm_e, ref_e, beta=0.5, micro_batch_size=4)
    g = torch.Generator().manual_seed(0)
    b, p, r = 8, 5, 6
    ids = torch.randint(3, 200, (b, p + r), generator=g)
    mask = torch.ones_like(ids)
    data = DataProto.from_dict({"input_ids": ids, "attention_mask": mask, "position_ids": torch.arange(p + r).expand(b, -1).contiguous(),
                                "responses": ids[:, p:], "response_mask": mask[:, p:]})
    data.meta_info["temperature"] = 1.0
    acc = torch.tensor([1.0, 0.0] * 4)
    q0 = prm.score(data)
    assert q0.shape == (b, r) and q0.abs().max() < 1e-6                      # identical weights -> zero implicit reward
    losses = [prm.update(data, acc)["reward_model/prm_ce_loss"] for _ in range(6)]
    assert abs(losses[0] - 0.6931) < 1e-3 and losses[-1] < 0.5 * losses[0]
    q = prm.score(data)
    seq = q.sum(-1)
    assert (seq[acc > 0].min() > seq[acc == 0].max()) and pairwise_accuracy(q, acc, data.batch["response_mask"].float(), 2) == 1.0


def test_pf_ppo_resampling_follows_the_scores():
    """algorithm.use_pf_ppo (reference: compute_pf_ppo_reweight_data): same batch size, rows drawn with replacement by score."""
    import numpy as np

    from polyrl_b200.protocol import DataProto
    from polyrl_b200.trainer.core import pf_ppo_resample

    scores = torch.tensor([0.0, 0.1, 1.0, 0.5, 0.9, 0.0, 0.2, 1.0])
    b = DataProto.from_dict({"token_level_scores": scores.unsqueeze(-1), "row": torch.arange(8)},
                            non_tensors={"uid": np.array([f"u{i}" for i in range(8)], dtype=object)})
    g = torch.Generator().manual_seed(0)
    out = pf_ppo_resample(b, "max_min", generator=g)
    assert len(out) == 8 and set(out.batch["row"].tolist()) <= {0, 2, 5, 7}              # only best / worst scored rows survive
    assert all(u == f"u{r}" for u, r in zip(out.non_tensor_batch["uid"], out.batch["row"].tolist()))
    rows = torch.cat([pf_ppo_resample(b, "pow", 2.0, generator=g).batch["row"] for _ in range(200)])
    freq = torch.bincount(rows, minlength=8).float() / rows.numel()
    assert freq[2] > 5 * freq[6] and freq[0] < 0.01                                       # ~ score^2
    rows = torch.cat([pf_ppo_resample(b, "max_random", generator=g).batch["row"] for _ in range(200)])
    freq = torch.bincount(rows, minlength=8).float() / rows.numel()
    assert 2.5 < (freq[2] / freq[1]).item() < 6                                           # best samples four times as likely
    with pytest.raises(ValueError):
        pf_ppo_resample(b, "nope")
