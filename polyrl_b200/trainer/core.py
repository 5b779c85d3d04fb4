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


def apply_kl_penalty(data: DataProto, kl_ctrl, kl_penalty: str = "kl") -> tuple[DataProto, dict]:
    mask = data.batch["response_mask"].float()
    scores = data.batch["token_level_scores"]
    kld = core_algos.kl_penalty(data.batch["old_log_probs"], data.batch["ref_log_prob"], kl_penalty) * mask
    beta = kl_ctrl.value
    data.batch["token_level_rewards"] = scores - beta * kld
    current_kl = core_algos.masked_mean(kld, mask, dim=-1).mean().item()
    kl_ctrl.update(current_kl=current_kl, n_steps=scores.shape[0])
    return data, {"actor/reward_kl_penalty": current_kl, "actor/reward_kl_penalty_coeff": beta}


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
