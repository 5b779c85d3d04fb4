"""Training data that changes between steps: a dataset that grows from a generator fed with the latest rollouts, and samplers that
re-weight prompts from training feedback (curriculum).  Counterpart of the reference's experimental dynamic dataset and sampler
interfaces (3rdparty/verl/verl/experimental/dynamic_dataset/dynamicgen_dataset.py, experimental/dataset/sampler.py:24-40).  The
stream trainer calls ``dataset.on_batch_end(batch)`` and ``sampler.update(batch)`` after every step when they exist."""
from __future__ import annotations

import importlib.util
from abc import ABC, abstractmethod

import numpy as np
from torch.utils.data import Sampler

from ..protocol import DataProto
from .rl_dataset import RLHFDataset


def load_extern_type(path: str | None, name: str):
    """Class ``name`` from a python file (or from this module when ``path`` is empty)."""
    from ..utils.import_utils import load_extern_type as _load

    if not path:
        return globals()[name]
    return _load(path, name)


class AbstractDataGenerator(ABC):
    def __init__(self, config: dict | None = None):
        self.config = config or {}

    @abstractmethod
    def generate(self, dataset: "DynamicGenDataset", batch: DataProto | None) -> list[dict]:
        """New dataset rows (same schema as the data files) proposed from the current dataset and the last training batch."""


class MockDataGenerator(AbstractDataGenerator):
    """Re-appends the first row — placeholder / test generator."""

    def generate(self, dataset, batch):
        return [dict(dataset.rows[0])]


class DynamicGenDataset(RLHFDataset):
    def __init__(self, *a, generator: AbstractDataGenerator | None = None, max_rows: int | None = None, **kw):
        super().__init__(*a, **kw)
        self.generator = generator or MockDataGenerator()
        self.max_rows = max_rows
        self.generated = 0

    def append_rows(self, rows: list[dict]) -> None:
        rows = [r for r in rows if self._prompt_len(r) <= self.max_len]
        if self.max_rows is not None:
            rows = rows[: max(0, self.max_rows - len(self.rows))]
        self.rows.extend(rows)
        self.generated += len(rows)

    def on_batch_end(self, batch: DataProto | None) -> None:
        self.append_rows(self.generator.generate(self, batch))


class AbstractSampler(Sampler):
    @abstractmethod
    def __init__(self, data_source, data_config=None):
        ...


class AbstractCurriculumSampler(AbstractSampler):
    @abstractmethod
    def update(self, batch: DataProto) -> None:
        ...


class SuccessRateCurriculumSampler(AbstractCurriculumSampler):
    """Draws prompts with probability proportional to ``p(1-p) + floor`` where ``p`` is the prompt's running success rate — prompts
    that are always solved or never solved are visited less often (they carry no GRPO signal).  Checkpointable like the default
    sampler (``state_dict`` / ``load_state_dict``)."""

    def __init__(self, data_source, data_config=None, seed: int = 1, floor: float = 0.05, momentum: float = 0.7):
        self.n, self.seed, self.floor, self.momentum = len(data_source), seed, floor, momentum
        self.rate = np.full(self.n, 0.5)
        self.epoch, self.cursor = 0, 0

    def _weights(self):
        w = self.rate * (1.0 - self.rate) + self.floor
        return w / w.sum()

    def __iter__(self):
        rng = np.random.default_rng(self.seed + self.epoch)
        while self.cursor < self.n:
            self.cursor += 1
            yield int(rng.choice(self.n, p=self._weights()))
        self.cursor, self.epoch = 0, self.epoch + 1

    def __len__(self) -> int:
        return self.n

    def update(self, batch: DataProto) -> None:
        idx = batch.non_tensor_batch.get("index")
        if idx is None or "token_level_scores" not in batch.batch:
            return
        score = batch.batch["token_level_scores"].sum(-1).float().cpu().numpy()
        per: dict[int, list] = {}
        for i, s in zip(idx, score):
            per.setdefault(int(i) % self.n, []).append(float(s > 0))
        for i, v in per.items():
            self.rate[i] = self.momentum * self.rate[i] + (1 - self.momentum) * float(np.mean(v))

    def state_dict(self) -> dict:
        return {"epoch": self.epoch, "cursor": self.cursor, "seed": self.seed, "rate": self.rate.tolist()}

    def load_state_dict(self, sd: dict) -> None:
        self.epoch, self.cursor, self.seed = sd["epoch"], sd["cursor"], sd["seed"]
        self.rate = np.asarray(sd["rate"], dtype=float)
