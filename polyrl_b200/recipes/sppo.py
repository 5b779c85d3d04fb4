"""SPPO — self-play preference optimisation (arXiv:2405.00675).  Reference behaviour: 3rdparty/verl/recipe/sppo/
sppo_ray_trainer.py:75-358 + dp_actor.py:34-57: no critic, no advantage estimator; every sampled response carries its scalar
score as an estimate of the win probability P(y > pi | x) and the policy regresses its sequence log-ratio onto
``eta * (P - 1/2)``.  Here: rewards are squashed to [0, 1] per group (min-max over the n samples of a prompt, 1/2 when they tie),
``advantages`` carries ``P - 1/2`` over the response tokens and the registered ``sppo`` policy loss does the regression.

    python -m polyrl_b200.recipes.sppo actor_rollout_ref.rollout.n=5 actor_rollout_ref.actor.policy_loss.sppo_eta=1.0 ...
"""
from __future__ import annotations

import sys

import torch

from ..protocol import DataProto
from ..trainer.controller_ppo import ControllerPPOTrainer

DEFAULTS = ["algorithm.adv_estimator=grpo", "algorithm.use_kl_in_reward=False", "actor_rollout_ref.actor.policy_loss.loss_mode=sppo",
            "actor_rollout_ref.actor.use_kl_loss=False", "actor_rollout_ref.actor.entropy_coeff=0.0"]


def win_probability(scores: torch.Tensor, uid) -> torch.Tensor:
    """Per-sample estimate of P(y beats the policy) from outcome scores: min-max inside the prompt group, ties -> 1/2."""
    p = torch.full_like(scores, 0.5)
    groups: dict = {}
    for i, u in enumerate(uid):
        groups.setdefault(u, []).append(i)
    for rows in groups.values():
        s = scores[rows]
        lo, hi = s.min(), s.max()
        if hi > lo:
            p[rows] = (s - lo) / (hi - lo)
    return p


class SPPOTrainer(ControllerPPOTrainer):
    def compute_advantages(self, batch: DataProto, metrics: dict) -> DataProto:
        mask = batch.batch["response_mask"].float()
        scores = (batch.batch["token_level_scores"] * mask).sum(-1)
        batch.batch["token_level_rewards"] = batch.batch["token_level_scores"]
        p = win_probability(scores, batch.non_tensor_batch["uid"])
        batch.batch["advantages"] = (p - 0.5).unsqueeze(-1) * mask
        batch.batch["returns"] = batch.batch["advantages"]
        metrics["sppo/win_prob_mean"] = p.mean().item()
        return batch


def main(argv: list[str] | None = None) -> int:
    from ._main import run_recipe

    return run_recipe(SPPOTrainer, DEFAULTS, argv)


if __name__ == "__main__":
    sys.exit(main())
