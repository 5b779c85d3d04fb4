"""Driver-side PPO step pieces shared by the streaming trainer, the synchronous trainer and bench.py:
reward -> (KL-in-reward) -> advantages, plus the per-(micro)batch pipeline that turns a rollout batch
into trained gradients.

Reference: ``apply_kl_penalty`` / ``compute_advantage`` (3rdparty/verl/verl/trainer/ppo/ray_trainer.py:
121-259) and the streaming hot loop (rlboost/verl_stream/trainer/ppo/stream_ray_trainer.py:391-564).
"""
from __future__ import annotations

import numpy as np
import torch

from ..algos import core_algos
from ..algos.core_algos import AdvantageEstimator
from ..protocol import DataProto


def apply_kl_penalty(data: DataProto, kl_ctrl, kl_penalty: str = "kl", group=None, real_rows: torch.Tensor | None = None,
                     distributed: bool = False) -> tuple[DataProto, dict]:
    """Token-level KL-in-reward.  ``data`` may be ONE data-parallel shard of the streamed batch: with ``distributed``
    (over ``group``, None = default group) the mean KL that drives the adaptive controller is all-reduced (sum of per-sequence KL, count of real sequences), so every
    rank's ``kl_ctrl`` stays bit-identical — the reference updates one controller on the driver from the whole batch
    (3rdparty/verl/verl/trainer/ppo/ray_trainer.py:121-159).  ``real_rows`` [n] bool excludes padded duplicates."""
    mask = data.batch["response_mask"].float()
    scores = data.batch["token_level_scores"]
    kld = core_algos.kl_penalty(data.batch["old_log_probs"], data.batch["ref_log_prob"], kl_penalty) * mask
    beta = kl_ctrl.value
    data.batch["token_level_rewards"] = scores - beta * kld
    per_seq = core_algos.masked_mean(kld, mask, dim=-1)
    real = torch.ones_like(per_seq) if real_rows is None else real_rows.to(per_seq)
    stat = torch.stack([(per_seq * real).sum(), real.sum()]).double()
    if distributed:
        torch.distributed.all_reduce(stat, group=group)
    n_real = max(int(round(stat[1].item())), 1)
    current_kl = float(stat[0].item() / n_real)
    kl_ctrl.update(current_kl=current_kl, n_steps=n_real)
    return data, {"actor/reward_kl_penalty": current_kl, "actor/reward_kl_penalty_coeff": beta}


def compute_group_advantage_sharded(local: DataProto, shard_ctx: dict, adv_estimator, gamma: float, lam: float,
                                    num_repeat: int, norm_adv_by_std_in_grpo: bool, config, group) -> DataProto:
    """Group-level advantages (GRPO / RLOO / OPO / GPG / ...) when the rewards only exist per DP shard (KL-in-reward needs
    the shard's old/ref log-probs).  The shards' ``token_level_rewards`` and masks are all-gathered in the global (balanced)
    row order recorded by the trainer, the estimator runs on the WHOLE streamed batch exactly like the reference's driver
    does (stream_ray_trainer.py:516-541), padded duplicates get private group ids so they cannot skew any group's statistics,
    and the rank keeps its own rows."""
    import torch.distributed as dist

    dp, dp_rank, sp = shard_ctx["dp"], shard_ctx["dp_rank"], shard_ctx.get("sp", 1)
    tlr, mask = local.batch["token_level_rewards"].contiguous(), local.batch["response_mask"].contiguous()
    world = dist.get_world_size(group)
    def gather(x):
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x, group=group)
        return torch.cat([parts[r * sp] for r in range(dp)], dim=0)      # SP peers hold the same shard: take one per DP rank
    g_tlr, g_mask = gather(tlr), gather(mask)
    uid = np.array(shard_ctx["uid"], dtype=object).copy()
    for i in np.nonzero(shard_ctx["is_pad"])[0]:
        uid[i] = f"__pad_{i}"
    whole = DataProto.from_dict({"token_level_rewards": g_tlr, "response_mask": g_mask}, non_tensors={"uid": uid})
    whole = compute_advantage(whole, adv_estimator, gamma, lam, num_repeat, norm_adv_by_std_in_grpo, config)
    n = len(local)
    sl = slice(dp_rank * n, (dp_rank + 1) * n)
    local.batch["advantages"] = whole.batch["advantages"][sl].clone()
    local.batch["returns"] = whole.batch["returns"][sl].clone()
    return local


def compute_advantage(data: DataProto, adv_estimator, gamma: float = 1.0, lam: float = 1.0, num_repeat: int = 1,
                      norm_adv_by_std_in_grpo: bool = True, config=None) -> DataProto:
    b = data.batch
    if "response_mask" not in b:
        r = b["responses"].shape[-1]
        b["response_mask"] = b["attention_mask"][:, -r:]
    mask = b["response_mask"].float()
    name = adv_estimator.value if isinstance(adv_estimator, AdvantageEstimator) else str(adv_estimator)
    fn = core_algos.get_adv_estimator_fn(name)
    kwargs = dict(token_level_rewards=b["token_level_rewards"], response_mask=mask, config=config)
    if name == "gae":
        kwargs.update(values=b["values"], gamma=gamma, lam=lam)
    elif name == "remax":
        kwargs.update(reward_baselines=b["reward_baselines"])
    elif name == "reinforce_plus_plus":
        kwargs.update(gamma=gamma)
    else:
        kwargs.update(index=data.non_tensor_batch["uid"], norm_adv_by_std_in_grpo=norm_adv_by_std_in_grpo)
    adv, ret = fn(**kwargs)
    b["advantages"], b["returns"] = adv, ret
    return data


def place_outcome_rewards(scores, response_mask: torch.Tensor) -> torch.Tensor:
    """Put each sample's scalar score on its last valid response token (reward-manager convention,
    reference: verl/workers/reward_manager/naive.py:46,104)."""
    s = torch.as_tensor(scores, dtype=torch.float32, device=response_mask.device)
    lens = response_mask.sum(-1).long()
    out = torch.zeros_like(response_mask, dtype=torch.float32)
    has = lens > 0
    idx = (lens - 1).clamp_min(0)
    out[torch.arange(out.shape[0], device=out.device)[has], idx[has]] = s[has]
    return out


def uid_array(n_prompts: int, n: int, prefix: str = "") -> np.ndarray:
    ids = np.empty(n_prompts * n, dtype=object)
    for i in range(n_prompts):
        for j in range(n):
            ids[i * n + j] = f"{prefix}{i}"
    return ids


def filter_uninformative_groups(batch: DataProto, metric: str = "seq_reward") -> tuple[DataProto, dict]:
    """DAPO dynamic sampling for streamed batches (reference: recipe/dapo/dapo_ray_trainer.py:190-260 keeps only prompts whose
    samples disagree on ``metric``).  Streamed batches hold complete groups and the optimizer-step boundaries are counted in
    samples, so uninformative groups are not removed but masked out: their ``response_mask`` becomes zero (no policy-gradient, KL,
    entropy or token-mean weight).  Returns the batch and ``dapo/*`` metrics."""
    uid = batch.non_tensor_batch["uid"]
    mask = batch.batch["response_mask"]
    if metric in ("seq_reward", "seq_final_reward"):
        key = "token_level_rewards" if metric == "seq_final_reward" and "token_level_rewards" in batch.batch else "token_level_scores"
        vals = (batch.batch[key] * mask).sum(-1).float().cpu()
    elif metric in batch.non_tensor_batch:
        vals = torch.tensor([float(v) for v in batch.non_tensor_batch[metric]])
    elif metric in batch.batch:
        vals = batch.batch[metric].float().reshape(len(batch), -1).sum(-1).cpu()
    else:
        raise KeyError(f"filter_groups.metric={metric!r} is not in the batch")
    groups: dict = {}
    for i, u in enumerate(uid):
        groups.setdefault(u, []).append(i)
    drop = [idx for idx in groups.values() if len(idx) > 1 and float(vals[idx].max() - vals[idx].min()) == 0.0]
    if drop:
        rows = torch.tensor([i for idx in drop for i in idx], dtype=torch.long, device=mask.device)
        new_mask = mask.clone()
        new_mask[rows] = 0
        batch.batch["response_mask"] = new_mask
    n = max(1, len(groups))
    return batch, {"dapo/num_groups": float(len(groups)), "dapo/filtered_groups": float(len(drop)),
                   "dapo/kept_group_ratio": 1.0 - len(drop) / n}


def pf_ppo_resample(batch: DataProto, reweight_method: str = "pow", weight_pow: float = 2.0, generator: torch.Generator | None = None) -> DataProto:
    """PF-PPO policy filtration (arXiv:2409.06957; reference: verl/trainer/ppo/core_algos.py compute_pf_ppo_reweight_data, switched by
    ``algorithm.use_pf_ppo``): resample the batch WITH replacement, the probability of a sample following its sequence score —
    ``pow``: |score|^weight_pow; ``max_min``: only the best- and worst-scored samples; ``max_random``: the best samples four
    times as likely as the rest.  Unreliable middle-of-the-road reward signals are filtered out, the batch size is unchanged."""
    scores = batch.batch["token_level_scores"].sum(-1).float()
    if reweight_method == "pow":
        w = scores.abs().pow(weight_pow)
    elif reweight_method == "max_min":
        w = ((scores == scores.max()) | (scores == scores.min())).float()
    elif reweight_method == "max_random":
        w = torch.where(scores == scores.max(), torch.full_like(scores, 0.4), torch.full_like(scores, 0.1))
    else:
        raise ValueError(f"unsupported pf_ppo reweight_method {reweight_method!r} (pow | max_min | max_random)")
    w = (w + 1e-8).clamp(min=1e-8)
    idx = torch.multinomial(w.cpu(), len(batch), replacement=True, generator=generator)
    return batch.select_idxs(idx)
