"""Experiment tracking fan-out: console / file(JSONL) / tensorboard always available; wandb / mlflow / swanlab
when importable (reference: 3rdparty/verl/verl/utils/tracking.py:27-172, ValidationGenerationsLogger :336-497)."""
from __future__ import annotations

import json
import os
import time
from pathlib import Path
from typing import Any


class _Console:
    def __init__(self, print_to_console=True):
        self.p = print_to_console

    def log(self, data: dict, step: int):
        if self.p:
            body = " - ".join(f"{k}:{v:.4g}" if isinstance(v, float) else f"{k}:{v}" for k, v in sorted(data.items()))
            print(f"step:{step} - {body}", flush=True)

    def finish(self):
        pass


class _File:
    def __init__(self, project, experiment):
        root = Path(os.environ.get("POLYRL_FILE_LOGGER_ROOT", "outputs/logs")) / project
        root.mkdir(parents=True, exist_ok=True)
        self.f = open(root / f"{experiment}.jsonl", "a")

    def log(self, data: dict, step: int):
        self.f.write(json.dumps({"step": step, "time": time.time(), **{k: _plain(v) for k, v in data.items()}}) + "\n")
        self.f.flush()

    def finish(self):
        self.f.close()


class _TensorBoard:
    def __init__(self, project, experiment):
        from torch.utils.tensorboard import SummaryWriter

        self.w = SummaryWriter(os.path.join(os.environ.get("TENSORBOARD_DIR", "outputs/tensorboard"), project, experiment))

    def log(self, data: dict, step: int):
        for k, v in data.items():
            if isinstance(v, (int, float)):
                self.w.add_scalar(k, v, step)

    def finish(self):
        self.w.close()


def _plain(v: Any):
    try:
        return float(v) if not isinstance(v, (str, bool, int, type(None), list, dict)) else v
    except Exception:
        return str(v)


class Tracking:
    supported = ["console", "file", "tensorboard", "wandb", "mlflow", "swanlab"]

    def __init__(self, project_name: str, experiment_name: str, default_backend="console", config: dict | None = None):
        if isinstance(default_backend, str):
            default_backend = [default_backend]
        self.loggers = {}
        for b in default_backend:
            if b not in self.supported:
                raise ValueError(f"unknown logger backend {b!r}; supported: {self.supported}")
            try:
                if b == "console":
                    self.loggers[b] = _Console()
                elif b == "file":
                    self.loggers[b] = _File(project_name, experiment_name)
                elif b == "tensorboard":
                    self.loggers[b] = _TensorBoard(project_name, experiment_name)
                elif b == "wandb":
                    import wandb

                    wandb.init(project=project_name, name=experiment_name, config=config)
                    self.loggers[b] = type("W", (), {"log": staticmethod(lambda d, s: wandb.log(d, step=s)), "finish": staticmethod(wandb.finish)})
                elif b == "mlflow":
                    import mlflow

                    mlflow.start_run(run_name=experiment_name)
                    self.loggers[b] = type("M", (), {"log": staticmethod(lambda d, s: mlflow.log_metrics({k.replace('@', '_at_'): v for k, v in d.items() if isinstance(v, (int, float))}, step=s)),
                                                     "finish": staticmethod(mlflow.end_run)})
                elif b == "swanlab":
                    import swanlab

                    swanlab.init(project=project_name, experiment_name=experiment_name, config=config)
                    self.loggers[b] = type("S", (), {"log": staticmethod(lambda d, s: swanlab.log(d, step=s)), "finish": staticmethod(swanlab.finish)})
            except ImportError as e:
                print(f"[tracking] backend {b!r} unavailable ({e}); skipping")

    def log(self, data: dict, step: int, backend=None) -> None:
        for name, lg in self.loggers.items():
            if backend is None or name in backend:
                lg.log(data, step)

    def finish(self) -> None:
        for lg in self.loggers.values():
            lg.finish()

    def __del__(self):
        try:
            self.finish()
        except Exception:
            pass


class ValidationGenerationsLogger:
    """Persist a sample of validation generations (inputs, outputs, scores) as JSONL per step."""

    def __init__(self, project_name: str = "", experiment_name: str = "", root: str = "outputs/val_generations"):
        self.dir = Path(root) / project_name / experiment_name

    def log(self, loggers, samples: list[tuple], step: int) -> None:
        self.dir.mkdir(parents=True, exist_ok=True)
        with open(self.dir / f"{step}.jsonl", "w") as f:
            for inp, out, score in samples:
                f.write(json.dumps({"input": inp, "output": out, "score": score, "step": step}) + "\n")
