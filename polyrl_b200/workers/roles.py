"""Role workers over the training-engine abstraction: the PPO-facing API (``compute_log_prob`` / ``update_actor`` /
``compute_values`` / ``update_critic`` on ``DataProto``) written once against ``BaseEngine``, so the same role code drives the ZeRO
engine, the pipeline-parallel engine or any other registered backend (reference: the new-style workers
3rdparty/verl/verl/workers/roles/actor.py / critic.py over verl/workers/engine/base.py, losses roles/utils/losses.py:27-108).

    eng = get_engine_cls("language_model", "pp")(model, device, pp_group, dp_group, lr=...)
    actor = ActorWorker(eng, ppo_mini_batch_size=64, ppo_epochs=1, clip_ratio=0.2)
    actor.init_model()
    data = data.union(actor.compute_log_prob(data))
    metrics = actor.update_actor(data)
"""
from __future__ import annotations

import torch

from ..protocol import DataProto
from .engine_api import BaseEngine, ppo_loss, value_loss

_TENSORS = ("input_ids", "attention_mask", "position_ids", "responses", "response_mask", "old_log_probs", "advantages", "returns",
            "values", "ref_log_prob", "rollout_is_weights")


def _as_batch(data: DataProto) -> dict:
    b = {k: data.batch[k] for k in _TENSORS if k in data.batch}
    if "response_mask" not in b:
        r = b["responses"].shape[1]
        b["response_mask"] = b["attention_mask"][:, -r:]
    return b


def _slice(b: dict, lo: int, hi: int) -> dict:
    return {k: v[lo:hi] for k, v in b.items()}


class ActorWorker:
    def __init__(self, engine: BaseEngine, ppo_mini_batch_size: int = 64, ppo_epochs: int = 1, clip_ratio: float = 0.2,
                 loss_agg_mode: str = "token-mean", loss_config=None, shuffle: bool = False, seed: int = 0):
        self.engine = engine
        self.mini, self.epochs, self.shuffle = ppo_mini_batch_size, ppo_epochs, shuffle
        self.loss_fn = ppo_loss(loss_config, clip_ratio, loss_agg_mode)
        self.gen = torch.Generator().manual_seed(seed)

    def init_model(self) -> None:
        self.engine.initialize()

    @torch.no_grad()
    def compute_log_prob(self, data: DataProto) -> DataProto:
        with self.engine.eval_mode():
            out = self.engine.infer_batch(_as_batch(data), temperature=data.meta_info.get("temperature", 1.0))
        return DataProto.from_dict({"old_log_probs": out["log_probs"].float(), "entropys": out["entropy"].float()})

    def update_actor(self, data: DataProto) -> dict:
        """``ppo_epochs`` passes over the batch in mini-batches; one optimizer step per mini-batch (the engine splits a mini-batch
        into micro-batches and accumulates)."""
        b = _as_batch(data)
        n = b["input_ids"].shape[0]
        temperature = data.meta_info.get("temperature", 1.0)
        metrics: dict[str, list] = {}
        with self.engine.train_mode():
            for _ in range(self.epochs):
                order = torch.randperm(n, generator=self.gen) if self.shuffle else torch.arange(n)
                shuffled = {k: v[order.to(v.device)] for k, v in b.items()}
                for lo in range(0, n, self.mini):
                    self.engine.optimizer_zero_grad()
                    m = self.engine.train_batch(_slice(shuffled, lo, min(n, lo + self.mini)), self.loss_fn, temperature=temperature)
                    st = self.engine.optimizer_step()
                    self.engine.lr_scheduler_step()
                    for k, v in {**m, "actor/grad_norm": st["grad_norm"], "actor/lr": st["lr"]}.items():
                        metrics.setdefault(k, []).append(float(v))
        return {k: sum(v) / len(v) for k, v in metrics.items()}

    def save_checkpoint(self, path: str, global_step: int = 0, max_ckpt_to_keep=None) -> None:
        self.engine.save_checkpoint(path, global_step, max_ckpt_to_keep)

    def load_checkpoint(self, path: str):
        return self.engine.load_checkpoint(path)


class CriticWorker:
    def __init__(self, engine: BaseEngine, ppo_mini_batch_size: int = 64, ppo_epochs: int = 1, cliprange_value: float = 0.5):
        self.engine = engine
        self.mini, self.epochs = ppo_mini_batch_size, ppo_epochs
        self.loss_fn = value_loss(cliprange_value)

    def init_model(self) -> None:
        self.engine.initialize()

    @torch.no_grad()
    def compute_values(self, data: DataProto) -> DataProto:
        with self.engine.eval_mode():
            out = self.engine.infer_batch(_as_batch(data))
        v = out["values"].float()
        mask = data.batch.get("response_mask")
        return DataProto.from_dict({"values": v * mask if mask is not None else v})

    def update_critic(self, data: DataProto) -> dict:
        b = _as_batch(data)
        n = b["input_ids"].shape[0]
        metrics: dict[str, list] = {}
        with self.engine.train_mode():
            for _ in range(self.epochs):
                for lo in range(0, n, self.mini):
                    self.engine.optimizer_zero_grad()
                    m = self.engine.train_batch(_slice(b, lo, min(n, lo + self.mini)), self.loss_fn)
                    st = self.engine.optimizer_step()
                    self.engine.lr_scheduler_step()
                    for k, v in {**m, "critic/grad_norm": st["grad_norm"]}.items():
                        metrics.setdefault(k, []).append(float(v))
        return {k: sum(v) / len(v) for k, v in metrics.items()}
