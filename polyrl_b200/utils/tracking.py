"""Experiment tracking fan-out.  ``trainer.logger`` names any of: console, file (JSONL), tensorboard, wandb, trackio, mlflow, swanlab,
clearml, vemlp_wandb.  The first three need nothing beyond this image; the others import their client library when selected (a
missing library skips that backend with a message, it never stops a run).  Each backend is a small adapter with ``log(data, step)``
/ ``finish()``; ``ValidationGenerationsLogger`` writes the sampled validation generations to every selected backend in that
backend's native form (wandb table that accumulates one row per step, swanlab text cards, an mlflow JSON artifact, a clearml table,
tensorboard text) and always to JSONL.  Capability parity: 3rdparty/verl/verl/utils/tracking.py:27-172 (back-ends, mlflow key
sanitisation + flattened params), :175-225 (ClearML), :336-497 (ValidationGenerationsLogger)."""
from __future__ import annotations

import dataclasses
import json
import logging
import os
import re
import time
from enum import Enum
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


def _scalars(data: dict) -> dict:
    """The numeric entries of a metrics dict as plain floats / ints (tensors and numpy scalars included)."""
    out = {}
    for k, v in data.items():
        if isinstance(v, bool):
            out[k] = int(v)
        elif isinstance(v, (int, float)):
            out[k] = v
        elif hasattr(v, "item") and (getattr(v, "ndim", 1) == 0 or (hasattr(v, "numel") and v.numel() == 1)):
            out[k] = float(v.item())
    return out


def flatten_params(params, sep: str = "/") -> dict[str, Any]:
    """A nested config (dicts, dataclasses, lists, Paths, Enums) as one flat ``a/b/c -> value`` dict: what parameter stores such as
    mlflow accept.  Lists become ``{list_len: n, "0": ..., "1": ...}`` so every leaf keeps a stable key."""
    flat: dict[str, Any] = {}

    def walk(x, prefix: str) -> None:
        if dataclasses.is_dataclass(x) and not isinstance(x, type):
            x = dataclasses.asdict(x)
        if isinstance(x, dict):
            for k, v in x.items():
                walk(v, f"{prefix}{sep}{k}" if prefix else str(k))
        elif isinstance(x, (list, tuple)):
            flat[f"{prefix}{sep}list_len" if prefix else "list_len"] = len(x)
            for i, v in enumerate(x):
                walk(v, f"{prefix}{sep}{i}" if prefix else str(i))
        elif isinstance(x, Path):
            flat[prefix] = str(x)
        elif isinstance(x, Enum):
            flat[prefix] = x.value
        else:
            flat[prefix] = x

    if params is not None:
        walk(params, "")
    return flat


_MLFLOW_BAD = re.compile(r"[^/\w.\- :]")            # mlflow metric keys: alphanumerics, _ - . / : and space


def sanitize_mlflow_key(key: str) -> str:
    """``val-core/acc@8`` -> ``val-core/acc_at_8``; every other character mlflow rejects becomes ``_``."""
    return _MLFLOW_BAD.sub("_", key.replace("@", "_at_"))


class _Wandb:
    """wandb and API-compatible clients (trackio, volcengine's vemlp_wandb): module-level init / log / finish."""

    def __init__(self, module, project, experiment, config, **init_kw):
        self.m = module
        module.init(project=project, name=experiment, config=config, **init_kw)

    def log(self, data: dict, step: int):
        self.m.log(data, step=step)

    def finish(self):
        getattr(self.m, "finish", lambda: None)()


class _Mlflow:
    def __init__(self, project, experiment, config):
        import mlflow

        self.m = mlflow
        self._warned: set[str] = set()
        mlflow.set_tracking_uri(os.environ.get("MLFLOW_TRACKING_URI", "sqlite:////tmp/mlruns.db"))
        exp = mlflow.set_experiment(project)
        mlflow.start_run(experiment_id=getattr(exp, "experiment_id", None), run_name=experiment)
        params = flatten_params(config)
        if params:
            mlflow.log_params(params)

    def log(self, data: dict, step: int):
        out = {}
        for k, v in _scalars(data).items():
            sk = sanitize_mlflow_key(k)
            if sk != k and k not in self._warned:
                self._warned.add(k)
                logging.getLogger(__name__).warning("[mlflow] metric key %r logged as %r", k, sk)
            out[sk] = v
        self.m.log_metrics(metrics=out, step=step)

    def finish(self):
        self.m.end_run()


class _Swanlab:
    def __init__(self, project, experiment, config):
        import swanlab

        self.m = swanlab
        if os.environ.get("SWANLAB_API_KEY"):
            swanlab.login(os.environ["SWANLAB_API_KEY"])
        swanlab.init(project=project, experiment_name=experiment, config={"FRAMEWORK": "polyrl_b200", **(config or {})},
                     logdir=os.environ.get("SWANLAB_LOG_DIR", "swanlog"), mode=os.environ.get("SWANLAB_MODE", "cloud"))

    def log(self, data: dict, step: int):
        self.m.log(data=data, step=step)

    def finish(self):
        self.m.finish()


class _ClearML:
    """ClearML: one task per run; a metric ``a/b/c`` is reported as title ``a`` / series ``b/c``; non-numeric values go to the text log."""

    def __init__(self, project, experiment, config):
        import clearml

        self.task = clearml.Task.init(task_name=experiment, project_name=project, continue_last_task=True, output_uri=False)
        self.task.connect_configuration(config or {}, name="Hyperparameters")

    def log(self, data: dict, step: int):
        lg = self.task.get_logger()
        nums = _scalars(data)
        for k, v in data.items():
            title, _, series = k.partition("/")
            if k in nums:
                lg.report_scalar(title=title, series=series or title, value=nums[k], iteration=step)
            else:
                lg.report_text(f"step {step}: {k} = {v}")

    def finish(self):
        self.task.close()


def _make_backend(name: str, project: str, experiment: str, config):
    if name == "console":
        return _Console()
    if name == "file":
        return _File(project, experiment)
    if name == "tensorboard":
        return _TensorBoard(project, experiment)
    if name in ("wandb", "tracking"):               # "tracking" = the reference's deprecated alias of wandb
        import wandb

        kw = {"entity": os.environ.get("WANDB_ENTITY")}
        proxy = ((config or {}).get("trainer") or {}).get("wandb_proxy") if isinstance(config, dict) else None
        if proxy:
            kw["settings"] = wandb.Settings(https_proxy=proxy)
        return _Wandb(wandb, project, experiment, config, **kw)
    if name == "trackio":
        import trackio

        return _Wandb(trackio, project, experiment, config)
    if name == "vemlp_wandb":
        import volcengine_ml_platform
        from volcengine_ml_platform import wandb as vemlp_wandb

        volcengine_ml_platform.init(ak=os.environ["VOLC_ACCESS_KEY_ID"], sk=os.environ["VOLC_SECRET_ACCESS_KEY"],
                                    region=os.environ["MLP_TRACKING_REGION"])
        return _Wandb(vemlp_wandb, project, experiment, config, sync_tensorboard=True)
    if name == "mlflow":
        return _Mlflow(project, experiment, config)
    if name == "swanlab":
        return _Swanlab(project, experiment, config)
    if name == "clearml":
        return _ClearML(project, experiment, config)
    raise ValueError(name)


class Tracking:
    supported = ["console", "file", "tensorboard", "wandb", "tracking", "trackio", "mlflow", "swanlab", "clearml", "vemlp_wandb"]

    def __init__(self, project_name: str, experiment_name: str, default_backend="console", config: dict | None = None):
        if isinstance(default_backend, str):
            default_backend = [default_backend]
        self.loggers = {}
        for b in default_backend:
            if b not in self.supported:
                raise ValueError(f"unknown logger backend {b!r}; supported: {self.supported}")
            try:
                self.loggers["wandb" if b == "tracking" else b] = _make_backend(b, project_name, experiment_name, config)
            except ImportError as e:
                print(f"[tracking] backend {b!r} unavailable ({e}); skipping")

    def log(self, data: dict, step: int, backend=None) -> None:
        for name, lg in self.loggers.items():
            if backend is None or name in backend:
                lg.log(data, step)

    def finish(self) -> None:
        for lg in self.loggers.values():
            lg.finish()
        self.loggers = {}

    def __del__(self):
        try:
            self.finish()
        except Exception:
            pass


class ValidationGenerationsLogger:
    """Persist a sample of validation generations (input, output, score) per validation step: always as JSONL, and in the native form
    of every selected tracking backend."""

    _wandb_rows: dict = {}            # (project, experiment) -> accumulated table rows: the wandb table grows by one row per step

    def __init__(self, project_name: str = "", experiment_name: str = "", root: str = "outputs/val_generations"):
        self.project, self.experiment = project_name, experiment_name
        self.dir = Path(root) / project_name / experiment_name

    def log(self, loggers, samples: list[tuple], step: int) -> None:
        self.dir.mkdir(parents=True, exist_ok=True)
        with open(self.dir / f"{step}.jsonl", "w") as f:
            for inp, out, score in samples:
                f.write(json.dumps({"input": inp, "output": out, "score": score, "step": step}) + "\n")
        names = [loggers] if isinstance(loggers, str) else list(loggers or ())
        for name in names:
            fn = getattr(self, f"_to_{'wandb' if name == 'tracking' else name}", None)
            if fn is None:
                continue
            try:
                fn(samples, step)
            except ImportError as e:
                print(f"[tracking] validation generations: backend {name!r} unavailable ({e}); skipping")

    # ---- one method per backend
    def _wandb_like(self, wandb, samples, step):
        cols = ["step"] + [f"{k}_{i + 1}" for i in range(len(samples)) for k in ("input", "output", "score")]
        rows = self._wandb_rows.setdefault((self.project, self.experiment, id(wandb)), [])
        rows.append([step] + [x for s in samples for x in s])
        width = max(len(r) for r in rows)                          # a later step may bring more samples: pad the older rows
        cols = cols if len(cols) == width else ["step"] + [f"{k}_{i + 1}" for i in range((width - 1) // 3) for k in ("input", "output", "score")]
        table = wandb.Table(columns=cols, data=[r + [None] * (width - len(r)) for r in rows])
        wandb.log({"val/generations": table}, step=step)

    def _to_wandb(self, samples, step):
        import wandb

        self._wandb_like(wandb, samples, step)

    def _to_vemlp_wandb(self, samples, step):
        from volcengine_ml_platform import wandb as vemlp_wandb

        self._wandb_like(vemlp_wandb, samples, step)

    def _to_swanlab(self, samples, step):
        import swanlab

        cards = [swanlab.Text(f"input: {i}\n\n---\n\noutput: {o}\n\n---\n\nscore: {s}", caption=f"sample {k + 1}")
                 for k, (i, o, s) in enumerate(samples)]
        swanlab.log({"val/generations": cards}, step=step)

    def _to_mlflow(self, samples, step):
        import mlflow

        rows = [{"input": i, "output": o, "score": s} for i, o, s in samples]
        if hasattr(mlflow, "log_text"):
            mlflow.log_text(json.dumps(rows, ensure_ascii=False, indent=1), f"val_step{step}.json")
        else:                                                      # older clients: write, then upload the file
            import tempfile

            with tempfile.TemporaryDirectory() as d:
                path = Path(d) / f"val_step{step}.json"
                path.write_text(json.dumps(rows, ensure_ascii=False, indent=1))
                mlflow.log_artifact(str(path))

    def _to_clearml(self, samples, step):
        import clearml
        import pandas as pd

        task = clearml.Task.current_task()
        if task is None:
            return
        df = pd.DataFrame([{"step": step, "input": i, "output": o, "score": s} for i, o, s in samples])
        task.get_logger().report_table(series="Validation generations", title="Validation", table_plot=df, iteration=step)

    def _to_tensorboard(self, samples, step):
        from torch.utils.tensorboard import SummaryWriter

        w = SummaryWriter(os.path.join(os.environ.get("TENSORBOARD_DIR", "outputs/tensorboard"), self.project, self.experiment))
        body = "\n\n".join(f"### sample {k + 1}\n- input: {i}\n- output: {o}\n- score: {s}" for k, (i, o, s) in enumerate(samples))
        w.add_text("val/generations", body, step)
        w.close()
