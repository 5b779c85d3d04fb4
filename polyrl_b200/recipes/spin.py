"""SPIN / online DPO (arXiv:2401.01335).  Reference behaviour: 3rdparty/verl/recipe/spin/spin_trainer.py:300-1400 +
core_algos.py:59-158: two responses per prompt, the better-scored one is 'chosen', the DPO logit is taken against the reference
policy and the update is a pairwise sigmoid (or IPO) loss; the reference policy can be refreshed from the actor every
``ref_update_freq`` steps (self-play iterations).  Here: ``rollout.n = 2`` keeps the two samples of a prompt adjacent, the sequence
score rides in ``advantages`` and the registered ``online_dpo`` policy loss forms the pairs.

    python -m polyrl_b200.recipes.spin actor_rollout_ref.actor.policy_loss.dpo_beta=0.1 ...
"""
from __future__ import annotations

import sys

from ..protocol import DataProto
from ..trainer.controller_ppo import ControllerPPOTrainer

DEFAULTS = ["algorithm.adv_estimator=grpo", "algorithm.use_kl_in_reward=False", "actor_rollout_ref.rollout.n=2",
            "actor_rollout_ref.actor.policy_loss.loss_mode=online_dpo", "actor_rollout_ref.actor.use_kl_loss=False",
            "actor_rollout_ref.actor.entropy_coeff=0.0"]


class SPINTrainer(ControllerPPOTrainer):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        if self.config.actor_rollout_ref.rollout.n != 2:
            raise ValueError("SPIN / online DPO pairs the two samples of a prompt: set actor_rollout_ref.rollout.n=2")

    def compute_advantages(self, batch: DataProto, metrics: dict) -> DataProto:
        mask = batch.batch["response_mask"].float()
        scores = (batch.batch["token_level_scores"] * mask).sum(-1)
        batch.batch["token_level_rewards"] = batch.batch["token_level_scores"]
        batch.batch["advantages"] = scores.unsqueeze(-1) * mask          # the loss reads the per-sequence score back out
        batch.batch["returns"] = batch.batch["advantages"]
        s = scores.view(-1, 2)
        metrics["spin/pair_margin"] = (s.max(-1).values - s.min(-1).values).mean().item()
        metrics["spin/tied_pairs"] = float((s[:, 0] == s[:, 1]).float().mean().item())
        return batch


def main(argv: list[str] | None = None) -> int:
    from ._main import run_recipe

    return run_recipe(SPINTrainer, DEFAULTS, argv)


if __name__ == "__main__":
    sys.exit(main())
