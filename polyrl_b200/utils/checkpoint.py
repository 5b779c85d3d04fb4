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
    def __init__(self, engine, model_cfg, save_contents=("model", "optimizer", "extra"), load_contents=None):
        self.engine, self.model_cfg = engine, model_cfg
        self.save_contents = set(save_contents)
        self.load_contents = set(load_contents or save_contents)
        self.saved: list[str] = []

    def _names(self, path: str):
        w, r = self.engine.world, self.engine.rank
        p = Path(path)
        return (p / f"model_world_size_{w}_rank_{r}.pt", p / f"optim_world_size_{w}_rank_{r}.pt",
                p / f"extra_state_world_size_{w}_rank_{r}.pt")

    def save(self, path: str, global_step: int = 0, max_ckpt_to_keep: int | None = None, extra: dict | None = None) -> None:
        eng = self.engine
        Path(path).mkdir(parents=True, exist_ok=True)
        mp, op, ep = self._names(path)
        if "model" in self.save_contents:
            torch.save({u.name: {"shard": u.shard.cpu(), "master": None if u.master is None else u.master.cpu()}
                        for u in eng.units}, mp)
        if "optimizer" in self.save_contents and eng.trainable:
            torch.save({"opt_step": eng.opt_step,
                        "units": {u.name: {"exp_avg": u.exp_avg.cpu(), "exp_avg_sq": u.exp_avg_sq.cpu()} for u in eng.units
                                  if u.exp_avg is not None}}, op)
        if "hf_model" in self.save_contents:
            from ..models.hf_io import save_hf

            save_hf(eng, self.model_cfg, Path(path) / "huggingface")      # consolidated HF directory (every rank gathers, rank 0 writes)
        if "extra" in self.save_contents:
            torch.save({"extra": extra or {}, "rng": _rng_state(eng.device), "global_step": global_step}, ep)
        if eng.rank == 0:
            self.model_cfg.save(Path(path) / "config.json")
        if dist.is_initialized() and eng.world > 1:
            dist.barrier(group=eng.group)
        self.saved.append(path)
        if max_ckpt_to_keep and max_ckpt_to_keep > 0 and eng.rank == 0:
            rotate_checkpoints(path, max_ckpt_to_keep)

    def load(self, path: str) -> dict | None:
        eng = self.engine
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
