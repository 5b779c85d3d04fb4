"""Shared entry point of the recipe trainers:  python -m polyrl_b200.recipes.<name> [controller.backend=mp|local|ray] key=value ..."""
from __future__ import annotations

import sys


def run_recipe(trainer_cls, defaults: list[str], argv: list[str] | None = None) -> int:
    from ..config.schema import load_config
    from ..models import get_model_config
    from ..trainer.controller_ppo import make_worker_group
    from ..trainer.main_stream import build_datasets
    from ..utils.tracking import Tracking

    argv = list(sys.argv[1:] if argv is None else argv)
    backend = "mp"
    rest = []
    for a in argv:
        if a.startswith("controller.backend="):
            backend = a.split("=", 1)[1]
        else:
            rest.append(a)
    cfg = load_config(overrides=defaults + rest)              # user overrides win over the recipe defaults
    world = max(1, cfg.trainer.n_gpus_per_node * cfg.trainer.nnodes)
    mc = get_model_config(cfg.actor_rollout_ref.model.path, **cfg.actor_rollout_ref.model.override_config)
    train, val, sampler, tok = build_datasets(cfg, mc)
    reward_fn = None
    if tok is not None:
        from ..data.reward_manager import load_reward_manager

        reward_fn = load_reward_manager(cfg, tok)
    tracking = Tracking(cfg.trainer.project_name, cfg.trainer.experiment_name, cfg.trainer.logger)
    trainer = trainer_cls(cfg, make_worker_group(cfg, backend, world), train, val, reward_fn, tracking, sampler)
    try:
        trainer.fit()
    finally:
        trainer.shutdown()
    return 0
