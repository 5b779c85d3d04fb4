"""Synchronous (non-streaming) colocated baseline entry — the counterpart of the reference's
``rlboost.verl_stream.trainer.main_ppo`` (verbatim verl ``main_ppo``): same config tree, every GPU rolls out then
trains, one streamed batch = the whole rollout batch.  Implemented by forcing the colocated settings and delegating
to ``main_stream`` (SPMD under torchrun), or — ``trainer.controller_backend=local|mp|ray`` — to the single-controller driver
``controller_ppo`` that spawns its own worker group."""
from __future__ import annotations

import sys

from .main_stream import main as _main


def main(argv: list[str] | None = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    backend = next((a.split("=", 1)[1] for a in argv if a.startswith("trainer.controller_backend=")), "spmd")
    if backend != "spmd":
        # single-controller execution (one driver process + worker group: local | mp | ray), the reference's Ray-driven layout
        from .controller_ppo import main as controller_main

        return controller_main([a for a in argv if not a.startswith("trainer.controller_backend=")] + [f"controller.backend={backend}"])
    forced = ["trainer.n_rollout_gpus=0"]
    bs = [a for a in argv if a.startswith("data.train_batch_size=")]
    n = [a for a in argv if a.startswith("actor_rollout_ref.rollout.n=")]
    if bs:
        total = int(bs[-1].split("=")[1]) * (int(n[-1].split("=")[1]) if n else 1)
        forced.append(f"actor_rollout_ref.rollout.min_stream_batch_size={total}")     # consume the batch in one piece
    return _main(argv + forced)


if __name__ == "__main__":
    sys.exit(main())
