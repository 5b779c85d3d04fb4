"""DAPO (arXiv:2503.14476) recipe trainer.  Reference behaviour: 3rdparty/verl/recipe/dapo/dapo_ray_trainer.py:76-330 —
clip-higher + token-mean loss + over-long reward shaping + **dynamic sampling**: generation batches of ``data.gen_batch_size``
prompts are drawn until ``data.train_batch_size`` prompts whose n samples disagree on ``filter_groups.metric`` have been collected
(at most ``filter_groups.max_num_gen_batches`` batches per step), the surplus is cut and ONE update runs on exactly
``train_batch_size * n`` trajectories.  (The streaming trainer cannot re-generate inside a step; it masks the uninformative
groups instead — ``trainer/core.py::filter_uninformative_groups``.)

    python -m polyrl_b200.recipes.dapo data.train_batch_size=512 data.gen_batch_size=1536 actor_rollout_ref.rollout.n=16 ...
"""
from __future__ import annotations

import sys
import time

import numpy as np
import torch

from ..protocol import DataProto, collate_fn
from ..trainer.controller_ppo import ControllerPPOTrainer

DEFAULTS = [
    "algorithm.adv_estimator=grpo", "algorithm.use_kl_in_reward=False", "algorithm.filter_groups.enable=True",
    "algorithm.filter_groups.metric=seq_reward", "algorithm.filter_groups.max_num_gen_batches=10",
    "actor_rollout_ref.actor.clip_ratio_low=0.2", "actor_rollout_ref.actor.clip_ratio_high=0.28",
    "actor_rollout_ref.actor.loss_agg_mode=token-mean", "actor_rollout_ref.actor.use_kl_loss=False",
    "reward_model.reward_manager=dapo",
]


def group_metric(batch: DataProto, metric: str) -> np.ndarray:
    mask = batch.batch["response_mask"]
    if metric in ("seq_reward", "seq_final_reward"):
        key = "token_level_rewards" if metric == "seq_final_reward" and "token_level_rewards" in batch.batch else "token_level_scores"
        return (batch.batch[key] * mask).sum(-1).float().numpy()
    if metric in batch.non_tensor_batch:
        return np.asarray([float(v) for v in batch.non_tensor_batch[metric]])
    if metric in batch.batch:
        return batch.batch[metric].float().reshape(len(batch), -1).sum(-1).numpy()
    raise KeyError(f"filter_groups.metric={metric!r} is not in the batch")


def keep_informative_groups(batch: DataProto, metric: str) -> tuple[DataProto, int, int]:
    """Rows of the groups whose samples do not all share one metric value (singleton groups are kept).  -> (batch, kept, total)"""
    vals = group_metric(batch, metric)
    groups: dict = {}
    for i, u in enumerate(batch.non_tensor_batch["uid"]):
        groups.setdefault(u, []).append(i)
    keep_rows: list[int] = []
    kept = 0
    for rows in groups.values():
        if len(rows) == 1 or float(np.std(vals[rows])) > 0.0:
            keep_rows.extend(rows)
            kept += 1
    return batch.select_idxs(keep_rows) if len(keep_rows) != len(batch) else batch, kept, len(groups)


class DAPOTrainer(ControllerPPOTrainer):
    def _gen_batches(self):
        cfg = self.config.data
        bs = cfg.gen_batch_size or cfg.train_batch_size
        while True:                                             # epochs are driven by fit(); a step may straddle an epoch end
            order = list(self.sampler) if self.sampler is not None else list(range(len(self.train_dataset)))
            for i in range(0, len(order) - bs + 1, bs):
                yield DataProto.from_single_dict(collate_fn([self.train_dataset[j] for j in order[i:i + bs]]))

    def fit(self, max_steps: int | None = None) -> list[dict]:
        cfg = self.config
        fg = cfg.algorithm.filter_groups
        n = cfg.actor_rollout_ref.rollout.n
        want = cfg.data.train_batch_size
        total = max_steps or cfg.trainer.total_training_steps or max(1, cfg.trainer.total_epochs) * max(
            1, len(self.train_dataset) // (cfg.data.gen_batch_size or want))
        history: list[dict] = []
        gen = self._gen_batches()
        self.gen_steps = 0
        while self.global_step < total:
            timing: dict[str, float] = {}
            t0 = time.time()
            acc: DataProto | None = None
            kept_prompts = num_gen = seen_groups = 0
            while True:
                new = self.rollout_and_score(next(gen), timing)
                self.gen_steps += 1
                num_gen += 1
                if fg.enable:
                    new, kept, groups = keep_informative_groups(new, fg.metric)
                else:
                    kept = groups = len(set(new.non_tensor_batch["uid"]))
                kept_prompts += kept
                seen_groups += groups
                if len(new):
                    acc = new if acc is None else DataProto.concat([acc, new])
                if not fg.enable or kept_prompts >= want:
                    break
                if 0 < fg.max_num_gen_batches <= num_gen:
                    raise RuntimeError(f"dynamic sampling: {kept_prompts} informative prompts after {num_gen} generation batches "
                                       f"(need {want}); the data may be too easy / too hard — set "
                                       "algorithm.filter_groups.max_num_gen_batches=0 to keep trying")
            batch = acc.slice(0, want * n) if fg.enable else acc          # align: exactly train_batch_size groups
            m = self.train_on(batch, timing)
            m.update({"dapo/num_gen_batches": float(num_gen), "dapo/kept_group_ratio": kept_prompts / max(1, seen_groups),
                      "dapo/gen_steps": float(self.gen_steps)})
            timing["step"] = time.time() - t0
            m.update({f"timing_s/{k}": v for k, v in timing.items()})
            if cfg.trainer.test_freq > 0 and self.global_step % cfg.trainer.test_freq == 0:
                m.update(self.validate())
            if self.tracking is not None:
                self.tracking.log(m, step=self.global_step)
            history.append(m)
        return history


def main(argv: list[str] | None = None) -> int:
    from ._main import run_recipe

    return run_recipe(DAPOTrainer, DEFAULTS, argv)


if __name__ == "__main__":
    sys.exit(main())
