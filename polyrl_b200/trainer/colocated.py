"""Colocated SPMD PPO/GRPO step: every rank owns one GPU, rolls out its shard of the prompts with its own
engine replica, then trains the ZeRO-3 sharded actor (and critic) together with its peers.

This is the "hybrid engine" mode of the reference (actor + rollout + ref fused per GPU,
3rdparty/verl/verl/single_controller/ray/base.py:749; time-sliced by rollout_mode()/trainer_mode(),
verl/workers/fsdp_workers.py:651-750) without the Ray round-trips: batches never leave the device
between phases.  ``ColocatedPPO.step`` is the public entry used by bench.py, the smoke test and the
synchronous trainer; the streaming/disaggregated orchestration lives in ``trainer/stream_trainer.py``.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from ..algos import core_algos
from ..config.schema import PPOConfig, need_critic, need_reference_policy
from ..protocol import DataProto
from ..workers.hybrid_worker import ActorRolloutRefWorker, CriticWorker
from . import core


class ColocatedPPO:
    def __init__(self, config: PPOConfig, device=None, group=None, reward_fn=None):
        self.config = config
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        role = "actor_rollout_ref" if need_reference_policy(config) else "actor_rollout"
        self.worker = ActorRolloutRefWorker(config, role=role, device=device, group=group)
        self.worker.init_model()
        self.device = self.worker.device
        self.critic = None
        if need_critic(config):
            self.critic = CriticWorker(config, device=self.device, group=group)
            self.critic.init_model(heap=self.worker.actor_engine.heap)
        self.kl_ctrl = core_algos.get_kl_controller(config.algorithm.kl_ctrl) if config.algorithm.use_kl_in_reward else None
        self.reward_fn = reward_fn
        self.global_step = 0

    # ------------------------------------------------------------------
    def _reward(self, batch: DataProto) -> torch.Tensor:
        """Scalar outcome reward per sample -> token-level scores on the last valid token."""
        mask = batch.batch["response_mask"]
        if self.reward_fn is not None:
            scores = self.reward_fn(batch)
            if torch.is_tensor(scores) and scores.dim() == 2:
                return scores.to(mask.device)
        else:
            from ..data.synthetic import synthetic_token_reward

            resp = batch.batch["responses"].cpu().numpy()
            lens = mask.sum(-1).cpu().numpy()
            rm = batch.non_tensor_batch["reward_model"]
            v = self.worker.model_cfg.vocab_size
            scores = [synthetic_token_reward(resp[i, : lens[i]], rm[i]["ground_truth"], v) for i in range(len(batch))]
        return core.place_outcome_rewards(scores, mask)

    def step(self, prompts: DataProto) -> dict:
        """One full PPO step on this rank's prompt shard.  ``prompts`` may live in (pinned) host memory."""
        cfg = self.config
        algo = cfg.algorithm
        timing: dict[str, float] = {}
        t_step = time.time()
        prompts = prompts.to(self.device, non_blocking=True)
        n = cfg.actor_rollout_ref.rollout.n
        if "uid" not in prompts.non_tensor_batch:
            prompts.non_tensor_batch["uid"] = core.uid_array(len(prompts), 1, prefix=f"s{self.global_step}r{self.rank}p")
        # ---- rollout
        t0 = time.time()
        batch = self.worker.generate_sequences(prompts)
        timing["gen"] = time.time() - t0
        # ---- reward
        t0 = time.time()
        batch.batch["token_level_scores"] = self._reward(batch)
        timing["reward"] = time.time() - t0
        # ---- old log-prob (recomputed by the trainer, as the reference does)
        t0 = time.time()
        batch.meta_info["temperature"] = cfg.actor_rollout_ref.rollout.temperature
        old = self.worker.compute_log_prob(batch)
        batch.batch["old_log_probs"] = old.batch["old_log_probs"]
        entropy = core_algos.masked_mean(old.batch["entropys"], batch.batch["response_mask"].float()).item()
        timing["old_log_prob"] = time.time() - t0
        if self.worker.has_ref:
            t0 = time.time()
            batch.batch["ref_log_prob"] = self.worker.compute_ref_log_prob(batch).batch["ref_log_prob"]
            timing["ref"] = time.time() - t0
        if self.critic is not None:
            t0 = time.time()
            batch.batch["values"] = self.critic.compute_values(batch).batch["values"]
            timing["values"] = time.time() - t0
        # ---- advantages
        t0 = time.time()
        metrics: dict = {}
        if algo.use_kl_in_reward:
            batch, klm = core.apply_kl_penalty(batch, self.kl_ctrl, algo.kl_penalty)
            metrics.update(klm)
        else:
            batch.batch["token_level_rewards"] = batch.batch["token_level_scores"]
        batch = core.compute_advantage(batch, algo.adv_estimator, algo.gamma, algo.lam, n, algo.norm_adv_by_std_in_grpo, algo)
        if algo.use_pf_ppo:
            batch = core.pf_ppo_resample(batch, algo.pf_ppo.get("reweight_method", "pow"), float(algo.pf_ppo.get("weight_pow", 2.0)))
        timing["adv"] = time.time() - t0
        # ---- update: split at mini-batch boundaries; optimizer step per mini-batch
        t0 = time.time()
        mini = self.worker.actor.config.ppo_mini_batch_size
        bounds = list(range(mini, len(batch) + 1, mini)) or [len(batch)]
        if bounds[-1] != len(batch):
            bounds.append(len(batch))
        start = 0
        upd_metrics: dict[str, list] = {}
        for bi, end in enumerate(bounds):
            part = batch.slice(start, end)
            last = bi == len(bounds) - 1
            part.meta_info = {**batch.meta_info, "is_opt_step": True, "is_lr_step": last}
            if self.critic is not None:
                cm = self.critic.update_critic_stream(part).meta_info["metrics"]
                for k, v in cm.items():
                    upd_metrics.setdefault(k, []).extend(v)
            if self.critic is None or self.global_step >= cfg.trainer.critic_warmup:
                am = self.worker.update_actor_stream(part).meta_info["metrics"]
                for k, v in am.items():
                    upd_metrics.setdefault(k, []).extend(v)
            start = end
        timing["update_actor"] = time.time() - t0
        # ---- weight sync to the rollout engine
        t0 = time.time()
        self.worker.update_weight_remote()
        timing["update_weight"] = time.time() - t0
        self.global_step += 1
        mask = batch.batch["response_mask"].float()
        seq_scores = batch.batch["token_level_scores"].sum(-1)
        metrics.update({k: float(np.mean(v)) for k, v in upd_metrics.items()})
        metrics.update({
            "critic/score/mean": seq_scores.mean().item(), "critic/advantages/mean": core_algos.masked_mean(batch.batch["advantages"], mask).item(),
            "response_length/mean": mask.sum(-1).mean().item(), "actor/entropy_old": entropy,
            "perf/total_num_tokens": int(batch.batch["attention_mask"].sum().item()), "num_samples": len(batch),
        })
        timing["step"] = time.time() - t_step
        metrics.update({f"timing_s/{k}": v for k, v in timing.items()})
        return metrics
