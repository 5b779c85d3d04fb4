"""Entropy-mechanism recipe (arXiv:2505.22617; reference: 3rdparty/verl/recipe/entropy/entropy_ray_trainer.py + the
``clip_cov`` / ``kl_cov`` policy losses): GRPO whose loss detaches (Clip-Cov) or KL-penalises (KL-Cov) the small set of tokens
with the largest log-prob / advantage covariance, with DAPO-style dynamic sampling.  ``entropy.mode=clip_cov|kl_cov``.

    python -m polyrl_b200.recipes.entropy entropy.mode=kl_cov ...
"""
from __future__ import annotations

import sys

from .dapo import DAPOTrainer

DEFAULTS = ["algorithm.adv_estimator=grpo", "algorithm.use_kl_in_reward=False", "algorithm.filter_groups.enable=True",
            "algorithm.filter_groups.metric=seq_reward", "algorithm.filter_groups.max_num_gen_batches=10",
            "actor_rollout_ref.actor.loss_agg_mode=token-mean", "actor_rollout_ref.actor.use_kl_loss=False",
            "actor_rollout_ref.actor.clip_ratio_low=0.2", "actor_rollout_ref.actor.clip_ratio_high=0.2"]


class EntropyMechanismTrainer(DAPOTrainer):
    """Same control flow as DAPO; the mechanism lives in the policy loss selected by the entry point."""


def main(argv: list[str] | None = None) -> int:
    from ._main import run_recipe

    argv = list(sys.argv[1:] if argv is None else argv)
    mode = "clip_cov"
    rest = []
    for a in argv:
        if a.startswith("entropy.mode="):
            mode = a.split("=", 1)[1]
        else:
            rest.append(a)
    if mode not in ("clip_cov", "kl_cov"):
        raise SystemExit(f"entropy.mode must be clip_cov or kl_cov, got {mode!r}")
    return run_recipe(EntropyMechanismTrainer, DEFAULTS + [f"actor_rollout_ref.actor.policy_loss.loss_mode={mode}"], rest)


if __name__ == "__main__":
    sys.exit(main())
