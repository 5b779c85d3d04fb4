"""Per-rank sharded checkpoints with rotation, latest-iteration tracker and RNG state.

Layout (reference: 3rdparty/verl/verl/utils/checkpoint/fsdp_checkpoint_manager.py:122-300,
verl/trainer/ppo/ray_trainer.py:777-895):
  {root}/global_step_{N}/{role}/model_world_size_{W}_rank_{r}.pt   bf16 shard + fp32 master
  {root}/global_step_{N}/{role}/optim_world_size_{W}_rank_{r}.pt   Adam moments + step
  {root}/global_step_{N}/{role}/extra_state_world_size_{W}_rank_{r}.pt   lr step, RNG states
  {root}/global_step_{N}/{role}/config.json                        model config (rank 0)
  {root}/latest_checkpointed_iteration.txt
"""
from __future__ import annotations

import os
import random
import re
import shutil
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist


def find_latest_ckpt_path(root: str | os.PathLike) -> str | None:
    root = Path(root)
    tracker = root / "latest_checkpointed_iteration.txt"
    if tracker.exists():
        step = int(tracker.read_text().strip())
        p = root / f"global_step_{step}"
        if p.exists():
            return str(p)
    cands = []
    if root.exists():
        for d in root.iterdir():
            m = re.fullmatch(r"global_step_(\d+)", d.name)
            if m and d.is_dir():
                cands.append((int(m.group(1)), d))
    return str(max(cands)[1]) if cands else None


def write_latest(root: str | os.PathLike, step: int) -> None:
    Path(root).mkdir(parents=True, exist_ok=True)
    tmp = Path(root) / ".latest.tmp"
    tmp.write_text(str(step))
    os.replace(tmp, Path(root) / "latest_checkpointed_iteration.txt")


def _rng_state(device) -> dict:
    st = {"cpu": torch.get_rng_state(), "numpy": np.random.get_state(), "random": random.getstate()}
    if torch.device(device).type == "cuda":
        st["cuda"] = torch.cuda.get_rng_state(device)
    return st


def _set_rng_state(st: dict, device) -> None:
    torch.set_rng_state(st["cpu"])
    np.random.set_state(st["numpy"])
    random.setstate(st["random"])
    if "cuda" in st and torch.device(device).type == "cuda":
        torch.cuda.set_rng_state(st["cuda"], device)


def rotate_checkpoints(latest_path: str | Path, keep: int) -> list[str]:
    """Keep the ``keep`` newest checkpoints, decided from what is ON DISK (a fresh manager per save, or a restarted job,
    rotates the same way).  ``latest_path`` is either ``<root>/global_step_N`` or a role directory below it
    (``<root>/global_step_N/actor``): in both cases whole ``global_step_*`` directories are removed, so no orphaned
    ``data.pt`` / sibling role directory stays behind.  Returns the removed directories."""
    p = Path(latest_path)
    step_dir = p if p.name.startswith("global_step_") else p.parent
    if not step_dir.name.startswith("global_step_") or keep is None or keep <= 0:
        return []
    def step_of(d: Path) -> int:
        try:
            return int(d.name.rsplit("_", 1)[-1])
        except ValueError:
            return -1
    steps = sorted((d for d in step_dir.parent.glob("global_step_*") if d.is_dir() and step_of(d) >= 0), key=step_of)
    doomed = [d for d in steps[:-keep] if d != step_dir]
    for d in doomed:
        shutil.rmtree(d, ignore_errors=True)
    return [str(d) for d in doomed]


class ShardedCheckpointManager:
    def __init__(self, engine, model_cfg, save_contents=("model", "optimizer", "extra"), load_contents=None, async_save: bool = False):
        self.engine, self.model_cfg = engine, model_cfg
        self.save_contents = set(save_contents)
        self.load_contents = set(load_contents or save_contents)
        self.saved: list[str] = []
        # checkpoint.async_save: the state is snapshotted to host memory synchronously (training may mutate the device tensors right
        # after), the files are written by a background thread; the next save / load / wait() joins it first
        self.async_save = bool(async_save)
        self._writer = None
        self._writer_error: BaseException | None = None

    def wait(self) -> None:
        w, self._writer = self._writer, None
        if w is not None:
            w.join()
        if self._writer_error is not None:
            e, self._writer_error = self._writer_error, None
            raise RuntimeError(f"asynchronous checkpoint write failed: {e!r}") from e

    def _names(self, path: str):
        w, r = self.engine.world, self.engine.rank
        p = Path(path)
        return (p / f"model_world_size_{w}_rank_{r}.pt", p / f"optim_world_size_{w}_rank_{r}.pt",
                p / f"extra_state_world_size_{w}_rank_{r}.pt")

    def save(self, path: str, global_step: int = 0, max_ckpt_to_keep: int | None = None, extra: dict | None = None) -> None:
        eng = self.engine
        self.wait()
        Path(path).mkdir(parents=True, exist_ok=True)
        mp, op, ep = self._names(path)
        files: list[tuple] = []                # (host-memory object, file): written inline or by the background thread

        def host(t):                           # async: a private host copy (``.cpu()`` of a CPU tensor would alias live memory)
            c = t.detach().cpu()
            return c.clone() if (self.async_save and c.data_ptr() == t.data_ptr()) else c

        if "model" in self.save_contents:
            files.append(({u.name: {"shard": host(u.shard), "master": None if u.master is None else host(u.master)} for u in eng.units}, mp))
        if "optimizer" in self.save_contents and eng.trainable:
            files.append(({"opt_step": eng.opt_step,
                           "units": {u.name: {"exp_avg": host(u.exp_avg), "exp_avg_sq": host(u.exp_avg_sq)} for u in eng.units
                                     if u.exp_avg is not None}}, op))
        if "hf_model" in self.save_contents:
            from ..models.hf_io import save_hf

            save_hf(eng, self.model_cfg, Path(path) / "huggingface")      # consolidated HF directory (every rank gathers, rank 0 writes)
        if "extra" in self.save_contents:
            files.append(({"extra": extra or {}, "rng": _rng_state(eng.device), "global_step": global_step}, ep))

        def write():
            for obj, f in files:
                tmp = Path(str(f) + ".tmp")
                torch.save(obj, tmp)
                os.replace(tmp, f)
            if eng.rank == 0:
                self.model_cfg.save(Path(path) / "config.json")

        if self.async_save:
            import threading

            def run():
                try:
                    write()
                except BaseException as e:          # noqa: BLE001  (surfaced by the next wait())
                    self._writer_error = e

            self._writer = threading.Thread(target=run, name="ckpt-writer", daemon=False)
            self._writer.start()
        else:
            write()
            if dist.is_initialized() and eng.world > 1:
                dist.barrier(group=eng.group)
        self.saved.append(path)
        if max_ckpt_to_keep and max_ckpt_to_keep > 0 and eng.rank == 0:
            rotate_checkpoints(path, max_ckpt_to_keep)

    def load(self, path: str) -> dict | None:
        eng = self.engine
        self.wait()
        mp, op, ep = self._names(path)
        if not mp.exists():
            raise FileNotFoundError(f"{mp} not found (checkpoints are per-rank; world size must match)")
        if "model" in self.load_contents:
            sd = torch.load(mp, map_location="cpu", weights_only=False)
            for u in eng.units:
                u.shard.copy_(sd[u.name]["shard"])
                if u.master is not None and sd[u.name]["master"] is not None:
                    u.master.copy_(sd[u.name]["master"])
                u.drop_gathered()
            eng._resident_stale = eng._reload_all = True
        if "optimizer" in self.load_contents and eng.trainable and op.exists():
            sd = torch.load(op, map_location="cpu", weights_only=False)
            eng.opt_step = sd["opt_step"]
            for u in eng.units:
                if u.exp_avg is None:
                    continue
                u.exp_avg.copy_(sd["units"][u.name]["exp_avg"])
                u.exp_avg_sq.copy_(sd["units"][u.name]["exp_avg_sq"])
        extra = None
        if "extra" in self.load_contents and ep.exists():
            sd = torch.load(ep, map_location="cpu", weights_only=False)
            _set_rng_state(sd["rng"], eng.device)
            extra = sd.get("extra")
        if dist.is_initialized() and eng.world > 1:
            dist.barrier(group=eng.group)
        return extra
