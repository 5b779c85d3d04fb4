"""FlowRL recipe (arXiv:2509.15207; reference: 3rdparty/verl/recipe/flowrl/{flowrl_actor,flowrl_fsdp_worker,flowrl_ray_trainer}.py):
GRPO-style sampling and group-normalised rewards, but the actor update is the trajectory-balance objective with a learned
log-partition head (``algos/flowrl.py``, ``workers/actor.py::FlowRLActor``).  The reference policy is always on.

    python -m polyrl_b200.recipes.flowrl actor_rollout_ref.actor.policy_loss.flowrl_beta=15 ...
"""
from __future__ import annotations

import sys

from ..trainer.controller_ppo import ControllerPPOTrainer

DEFAULTS = ["algorithm.adv_estimator=grpo", "algorithm.use_kl_in_reward=False", "actor_rollout_ref.actor.policy_loss.loss_mode=flowrl",
            "actor_rollout_ref.actor.use_kl_loss=False", "actor_rollout_ref.actor.entropy_coeff=0.0",
            "actor_rollout_ref.actor.clip_ratio_low=0.2", "actor_rollout_ref.actor.clip_ratio_high=0.28"]


class FlowRLTrainer(ControllerPPOTrainer):
    """Control flow of the stock trainer; the objective lives in the actor the workers build for ``loss_mode=flowrl``."""


def main(argv: list[str] | None = None) -> int:
    from ._main import run_recipe

    return run_recipe(FlowRLTrainer, DEFAULTS, argv)


if __name__ == "__main__":
    sys.exit(main())
