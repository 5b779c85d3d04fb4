"""GSPO — group sequence policy optimisation (arXiv:2507.18071; reference: 3rdparty/verl/recipe/gspo/*.sh drive the stock
trainer with ``loss_mode=gspo``, sequence-level importance ratios, ``seq-mean-token-mean`` aggregation and very small clip
ranges).  The recipe is its defaults.

    python -m polyrl_b200.recipes.gspo ...
"""
from __future__ import annotations

import sys

from ..trainer.controller_ppo import ControllerPPOTrainer

DEFAULTS = ["algorithm.adv_estimator=grpo", "algorithm.use_kl_in_reward=False", "actor_rollout_ref.actor.policy_loss.loss_mode=gspo",
            "actor_rollout_ref.actor.loss_agg_mode=seq-mean-token-mean", "actor_rollout_ref.actor.clip_ratio_low=0.0003",
            "actor_rollout_ref.actor.clip_ratio_high=0.0004", "actor_rollout_ref.actor.use_kl_loss=False"]


class GSPOTrainer(ControllerPPOTrainer):
    pass


def main(argv: list[str] | None = None) -> int:
    from ._main import run_recipe

    return run_recipe(GSPOTrainer, DEFAULTS, argv)


if __name__ == "__main__":
    sys.exit(main())
