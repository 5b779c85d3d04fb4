"""PRIME recipe (arXiv:2502.01456; reference: 3rdparty/verl/recipe/prime/{prime_ray_trainer,prime_dp_rm,prime_core_algos}.py): the
policy is trained on the sum of the sparse outcome reward and DENSE token rewards from an implicit process reward model
``beta * log(pi_prm / pi_ref)`` that is itself updated online from the outcome labels.

Per step (single-controller driver over the worker group):
  1. over-sample ``data.oversample_factor`` x ``train_batch_size`` prompts, generate n responses each, score outcomes (``acc``);
  2. ``data.filter_accuracy``: keep prompt groups whose mean accuracy lies in [accuracy_lower_bound, accuracy_upper_bound]
     (all-right / all-wrong groups carry no preference signal), refill from further generation batches, cut to the batch size;
  3. update the PRM on (responses, acc) before or after scoring (``reward_model.model.update``), score the batch with it;
  4. advantages = leave-one-out outcome reward x ``reward_gt_coef`` + leave-one-out process rewards x ``reward_dpo_coef``
     (``algos/prime.py::prime_advantages``), then the usual PPO-clip update.

    python -m polyrl_b200.recipes.prime data.filter_accuracy=True data.oversample_factor=4 reward_model.model.update=before ...
"""
from __future__ import annotations

import sys
import time

import numpy as np
import torch

from ..algos import prime as P
from ..protocol import DataProto, collate_fn
from ..trainer.controller_ppo import ControllerPPOTrainer
from ..utils import metrics as M

DEFAULTS = ["algorithm.adv_estimator=rloo", "algorithm.use_kl_in_reward=False", "actor_rollout_ref.actor.use_kl_loss=False",
            "reward_model.reward_manager=prime", "reward_model.model.update=before", "reward_model.model.beta_train=0.05",
            "data.filter_accuracy=True", "data.accuracy_lower_bound=0.2", "data.accuracy_upper_bound=0.8", "data.oversample_factor=4"]


def filter_by_accuracy(batch: DataProto, lo: float, hi: float) -> tuple[DataProto, int, int]:
    """Rows of the prompt groups whose mean ``acc`` lies inside [lo, hi].  -> (batch, kept groups, all groups)"""
    acc = batch.batch["acc"].float().numpy()
    groups: dict = {}
    for i, u in enumerate(batch.non_tensor_batch["uid"]):
        groups.setdefault(u, []).append(i)
    rows, kept = [], 0
    for idx in groups.values():
        m = float(np.mean(acc[idx]))
        if lo <= m <= hi:
            rows.extend(idx)
            kept += 1
    return (batch.select_idxs(rows) if len(rows) != len(batch) else batch), kept, len(groups)


class PRIMETrainer(ControllerPPOTrainer):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        if self.config.actor_rollout_ref.rollout.n < 2:
            raise ValueError("PRIME's leave-one-out baselines need rollout.n >= 2")
        self.prm_params = self.wg.init_prm()[0]

    def _prompt_batches(self, size: int):
        while True:
            order = list(self.sampler) if self.sampler is not None else list(range(len(self.train_dataset)))
            for i in range(0, len(order) - size + 1, size):
                yield DataProto.from_single_dict(collate_fn([self.train_dataset[j] for j in order[i:i + size]]))

    def _acc(self, batch: DataProto) -> torch.Tensor:
        """Outcome label per sample in [0, 1]: the sequence score of the rule-based / custom scorer."""
        return (batch.batch["token_level_scores"] * batch.batch["response_mask"]).sum(-1).clamp(0.0, 1.0)

    def compute_advantages(self, batch: DataProto, metrics: dict) -> DataProto:
        algo, rm = self.config.algorithm, self.config.reward_model
        n = self.config.actor_rollout_ref.rollout.n
        mask = batch.batch["response_mask"].float()
        acc = batch.batch["acc"]
        when = rm.model.update
        upd: dict[str, list] = {}
        if when == "before":
            M.append_to_dict(upd, self.wg.update_prm(batch).meta_info["metrics"])
        q = self.wg.compute_prm_score(batch).batch["prm_scores"]
        if when == "after":
            M.append_to_dict(upd, self.wg.update_prm(batch).meta_info["metrics"])
        adv, ret = P.prime_advantages(mask, n, acc=acc, process_rewards=q, reward_gt_coef=algo.reward_gt_coef, reward_dpo_coef=algo.reward_dpo_coef)
        batch.batch["token_level_rewards"] = batch.batch["token_level_scores"]
        batch.batch["advantages"], batch.batch["returns"] = adv, ret
        metrics.update({k: float(np.mean(v)) for k, v in upd.items()})
        metrics["reward_model/pairwise_acc"] = P.pairwise_accuracy(q, acc, mask, n)
        metrics["reward_model/process_reward_mean"] = float((q * mask).sum() / mask.sum().clamp(min=1))
        return batch

    def fit(self, max_steps: int | None = None) -> list[dict]:
        cfg, d = self.config, self.config.data
        n, want = cfg.actor_rollout_ref.rollout.n, cfg.data.train_batch_size
        gen_size = max(want, int(round(want * (d.oversample_factor if d.filter_accuracy else 1.0))))
        gen_size -= gen_size % self.wg.world_size or 0
        total = max_steps or cfg.trainer.total_training_steps or max(1, len(self.train_dataset) // max(1, gen_size))
        gen = self._prompt_batches(max(gen_size, self.wg.world_size))
        history: list[dict] = []
        while self.global_step < total:
            timing: dict[str, float] = {}
            t0 = time.time()
            acc_batch, kept_total, seen_total, rounds = None, 0, 0, 0
            while True:
                new = self.rollout_and_score(next(gen), timing)
                new.batch["acc"] = self._acc(new)
                rounds += 1
                if d.filter_accuracy:
                    new, kept, seen = filter_by_accuracy(new, d.accuracy_lower_bound, d.accuracy_upper_bound)
                else:
                    kept = seen = len(set(new.non_tensor_batch["uid"]))
                kept_total, seen_total = kept_total + kept, seen_total + seen
                if len(new):
                    acc_batch = new if acc_batch is None else DataProto.concat([acc_batch, new])
                if kept_total >= want or not d.filter_accuracy:
                    break
                if rounds >= 8:                               # the data offers too few informative prompts: train on what there is
                    break
            if acc_batch is None or len(acc_batch) < n * self.wg.world_size:
                raise RuntimeError("accuracy filter removed (almost) every prompt group; widen data.accuracy_lower/upper_bound")
            groups = min(want, len(acc_batch) // n)
            groups -= groups % max(1, self.wg.world_size // np.gcd(self.wg.world_size, n))      # rows must split over the workers
            batch = acc_batch.slice(0, max(1, groups) * n)
            m = self.train_on(batch, timing)
            m.update({"prime/kept_group_ratio": kept_total / max(1, seen_total), "prime/gen_rounds": float(rounds),
                      "prime/acc_mean": float(batch.batch["acc"].mean())})
            timing["step"] = time.time() - t0
            m.update({f"timing_s/{k}": v for k, v in timing.items()})
            if self.tracking is not None:
                self.tracking.log(m, step=self.global_step)
            history.append(m)
        return history


def main(argv: list[str] | None = None) -> int:
    from ._main import run_recipe

    return run_recipe(PRIMETrainer, DEFAULTS, argv)


if __name__ == "__main__":
    sys.exit(main())
