"""Replay dumped generations instead of running the rollout engine — for debugging the training half of a step with a fixed,
reproducible batch (reference: 3rdparty/verl/verl/utils/rollout_skip.py:19 ``RolloutSkip``).

    skip = RolloutSkip(config, dump_dir)          # config.actor_rollout_ref.rollout.skip_rollout-style switch
    skip.wrap_generate_sequences(worker)          # worker.generate_sequences now loads <dump_dir>/..._step{k}.pt when present,
                                                  # otherwise generates and dumps
"""
from __future__ import annotations

from pathlib import Path

from ..protocol import DataProto


class RolloutSkip:
    def __init__(self, dump_dir: str, exp_name: str = "rollout", n: int = 1, train_batch_size: int = 0):
        self.dir = Path(dump_dir).expanduser()
        self.dir.mkdir(parents=True, exist_ok=True)
        self.prefix = f"{exp_name}_n{n}_bs{train_batch_size}"
        self.step = 0
        self.hits = 0

    def path(self, step: int) -> Path:
        return self.dir / f"{self.prefix}_step{step}.pt"

    def wrap_generate_sequences(self, worker) -> None:
        inner = worker.generate_sequences

        def generate_sequences(prompts: DataProto) -> DataProto:
            p = self.path(self.step)
            self.step += 1
            if p.exists():
                self.hits += 1
                out = DataProto.load_from_disk(str(p))
                dev = prompts.batch["input_ids"].device if prompts.batch is not None else None
                return out.to(dev) if dev is not None else out
            out = inner(prompts)
            out.clone().to("cpu").save_to_disk(str(p))       # DataProto.to moves in place: dump a copy
            return out

        worker.generate_sequences = generate_sequences
