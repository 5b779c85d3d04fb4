"""Streaming PPO/GRPO trainer: rollout results arrive in micro-batches and the actor/critic forward/backward
run while generation is still in flight; optimizer steps happen at mini-batch boundaries.

Behavioural parity with ``StreamRayPPOTrainer.fit`` (rlboost/verl_stream/trainer/ppo/stream_ray_trainer.py:
282-689; mini-batch boundaries :239-278,500-564; metrics feedback :691-704) and verl's ``RayPPOTrainer``
checkpoint / validation / balance logic (3rdparty/verl/verl/trainer/ppo/ray_trainer.py:529-668,777-954).
Execution model is SPMD: every trainer rank runs this loop; rank 0 owns the manager connection and
broadcasts each streamed batch, every rank then takes its data-parallel shard (no driver<->worker object-store
round trips).  Without a manager (colocated sync mode) the batch is generated locally and then consumed in
``min_stream_batch_size`` slices, so both modes share one code path.
"""
from __future__ import annotations

import os
import time
import uuid
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

from ..algos import core_algos
from ..config.schema import PPOConfig
from ..protocol import DataProto, collate_fn, pad_dataproto_to_divisor
from ..utils import metrics as M
from ..utils.checkpoint import find_latest_ckpt_path, write_latest
from ..utils.profiler import marked_timer
from ..utils.seqlen_balancing import get_seqlen_balanced_partitions, log_seqlen_unbalance
from . import core


def _bcast_obj(obj, src: int = 0, group=None):
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


class RemoteRolloutPlaceholder:
    """What trainer ranks other than 0 hold instead of the manager connection: rank 0 talks to the rollout manager and broadcasts
    every streamed batch; the other ranks only need to know that generation is remote (``client is not None``)."""

    timing: dict = {}


class StreamPPOTrainer:
    def __init__(self, config: PPOConfig, worker, critic=None, reward_fn=None, val_reward_fn=None, train_dataset=None,
                 val_dataset=None, rollout_client=None, group=None, tracking=None, train_sampler=None):
        self.config = config
        self.worker, self.critic = worker, critic
        self.reward_fn, self.val_reward_fn = reward_fn, val_reward_fn or reward_fn
        self.train_dataset, self.val_dataset = train_dataset, val_dataset
        self.client = rollout_client
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.tracking = tracking
        self.sampler = train_sampler
        self.device = worker.device
        self.global_steps = 0
        self.use_critic = critic is not None
        self.use_reference = worker.has_ref
        self.kl_ctrl = core_algos.get_kl_controller(config.algorithm.kl_ctrl) if config.algorithm.use_kl_in_reward else None
        bs = config.data.train_batch_size
        n_batches = max(1, len(train_dataset) // bs) if train_dataset is not None else 0
        self.total_training_steps = config.trainer.total_training_steps or n_batches * config.trainer.total_epochs
        config.actor_rollout_ref.actor.optim.total_training_steps = self.total_training_steps
        self._esi_deadline = float(os.environ.get("POLYRL_ESI_EXPIRATION_TS", "0") or 0)
        self.rm = None              # model-based reward (reward_model.enable): RewardModelWorker set by the entry point

    def _model_reward(self, batch: DataProto) -> None:
        """reward_model.enable: the value-head reward model scores the batch — every rank takes a contiguous slice, the slices are
        all-gathered — and the result rides along as ``rm_scores``, which the reward managers use instead of (or combine with) the
        rule-based scorers (reference: ray_trainer.py compute_rm_score before compute_reward)."""
        if self.rm is None or "rm_scores" in batch.batch:
            return
        n = len(batch)
        per = (n + self.world - 1) // self.world
        lo, hi = min(n, self.rank * per), min(n, (self.rank + 1) * per)
        r = batch.batch["responses"].shape[1]
        mine = torch.zeros(per, r, device=self.device)
        if hi > lo:
            mine[: hi - lo] = self.rm.compute_rm_score(batch.slice(lo, hi)).batch["rm_scores"].to(self.device)
        if self.world > 1:
            parts = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            mine = torch.cat(parts)
        batch.batch["rm_scores"] = mine[:n]

    # ------------------------------------------------------------------ data
    def _iter_batches(self):
        bs = self.config.data.train_batch_size
        ds = self.train_dataset
        for _ in range(self.config.trainer.total_epochs):
            if self.sampler is not None and hasattr(self.sampler, "update"):
                # curriculum samplers are re-weighted after every step: draw lazily, one batch at a time
                import itertools

                it = iter(self.sampler)
                while True:
                    idx = list(itertools.islice(it, bs))
                    if len(idx) < bs:
                        for _ in it:                 # run the iterator out so it closes its epoch
                            pass
                        break
                    yield DataProto.from_single_dict(collate_fn([ds[j % len(ds)] for j in idx]))
                continue
            order = list(iter(self.sampler)) if self.sampler is not None else list(range(len(ds)))
            for i in range(0, len(order) - bs + 1, bs):
                yield DataProto.from_single_dict(collate_fn([ds[j] for j in order[i:i + bs]]))

    def _data_feedback(self, full: DataProto) -> None:
        """Dynamic datasets / curriculum samplers learn from the finished step (reference: experimental/dataset/sampler.py
        ``update(batch)``, dynamicgen_dataset.py ``on_batch_end``).  Every rank must keep an identical dataset and sampler, so the
        per-rank shards are merged first."""
        want_s = self.sampler is not None and hasattr(self.sampler, "update")
        want_d = hasattr(self.train_dataset, "on_batch_end")
        if not (want_s or want_d):
            return
        fb = DataProto.from_dict({"token_level_scores": full.batch["token_level_scores"].detach().cpu(),
                                  "responses": full.batch["responses"].cpu(), "response_mask": full.batch["response_mask"].cpu()},
                                 non_tensors={k: full.non_tensor_batch[k] for k in ("index", "uid", "data_source") if k in full.non_tensor_batch})
        if self.world > 1:
            parts = [None] * self.world
            dist.all_gather_object(parts, fb, group=self.group)
            fb = DataProto.concat(parts)
        if want_s:
            self.sampler.update(fb)
        if want_d:
            self.train_dataset.on_batch_end(fb)

    def _mini_batch_boundaries(self, total: int) -> list[int]:
        """Cumulative sample counts at which the optimizer steps (global samples = prompts * n)."""
        a = self.config.actor_rollout_ref
        mini = a.actor.ppo_mini_batch_size * a.rollout.n
        b = list(range(mini, total + 1, mini))
        if not b or b[-1] != total:
            b.append(total)           # clamp: the reference crashes when train_batch_size*n < mini
        return b

    # ------------------------------------------------------------------ rollout stream
    def _stream(self, prompts: DataProto):
        """Yield streamed rollout batches (identical on every rank)."""
        r = self.config.actor_rollout_ref.rollout
        if self.client is None:
            batch = self.worker.generate_sequences(prompts.to(self.device))
            self._timing["gen"] = batch.meta_info.get("timing", {}).get("generate_sequences", 0.0)
            step = r.min_stream_batch_size
            for i in range(0, len(batch), step):
                yield batch.slice(i, min(i + step, len(batch)))
            return
        if getattr(self, "_prefetched", None) is not None:
            # one-step-off: this batch was generated during the previous step (rank 0 holds the list)
            items, self._prefetched = self._prefetched, None
            k = 0
            while True:
                item = items[k] if (self.rank == 0 and k < len(items)) else None
                k += 1
                item = _bcast_obj(item, 0, self.group)
                if item is None:
                    break
                yield item.to(self.device)
            return
        it = None
        hybrid = getattr(self.worker, "local_server", None) is not None
        if hybrid:
            self.worker.rollout_mode()                  # every trainer GPU generates during the manager's local window
            if dist.is_initialized() and self.world > 1:
                dist.barrier(group=self.group)
        if self.rank == 0:
            sp = {"temperature": r.temperature, "top_p": r.top_p, "top_k": r.top_k}
            it = iter(self.client.generate_sequences_remote(prompts, sp))
        first = True
        while True:
            item = None
            if self.rank == 0:
                try:
                    item = next(it)                     # the first item arrives after the notifier = end of the local window
                except StopIteration:
                    item = None
            item = _bcast_obj(item, 0, self.group)
            if first and hybrid:
                self.worker.trainer_mode()              # local engines were aborted by the manager: HBM goes back to training
            first = False
            if item is None:
                break
            yield item.to(self.device)
        if self.rank == 0:
            self._timing["gen"] = self.client.timing.get("gen", 0.0)
            self._timing["local_gen"] = self.client.timing.get("local_gen", 0.0)

    def _balance_and_shard(self, batch: DataProto, metrics: dict) -> DataProto:
        """Karmarkar-Karp reorder so every DP rank gets a similar token count, then take this rank's shard."""
        dp, dp_rank = getattr(self.worker, "dp_size", self.world), getattr(self.worker, "dp_rank", self.rank)
        self._shard_ctx = None
        if dp == 1:
            return batch
        batch, pad = pad_dataproto_to_divisor(batch, dp)
        is_pad = np.zeros(len(batch), dtype=bool)
        if pad:
            batch.batch["response_mask"][-pad:] = 0        # padded duplicates carry no loss
            is_pad[-pad:] = True
        if self.config.trainer.balance_batch:
            lens = batch.batch["attention_mask"].sum(-1).tolist()
            parts = get_seqlen_balanced_partitions([int(x) for x in lens], dp, equal_size=True)
            metrics.update(log_seqlen_unbalance([int(x) for x in lens], parts, "global_seqlen"))
            order = [i for p in parts for i in p]
            batch.reorder(torch.tensor(order, dtype=torch.long))
            is_pad = is_pad[np.asarray(order)]
        # global (post-balance) row order: lets a shard-local phase re-assemble whole GRPO/RLOO groups across DP ranks
        self._shard_ctx = {"dp": dp, "dp_rank": dp_rank, "sp": getattr(self.worker, "sp_size", 1), "is_pad": is_pad,
                           "uid": batch.non_tensor_batch["uid"].copy() if "uid" in batch.non_tensor_batch else None}
        return batch.chunk(dp)[dp_rank]       # ranks of one Ulysses SP group share the same shard

    def _reward(self, batch: DataProto) -> torch.Tensor:
        if self.reward_fn is None and "rm_scores" in batch.batch:
            return batch.batch["rm_scores"].to(self.device)            # reward model only (no rule-based scorer configured)
        if self.reward_fn is None:
            from ..data.synthetic import synthetic_token_reward

            mask = batch.batch["response_mask"]
            resp, lens = batch.batch["responses"].cpu().numpy(), mask.sum(-1).cpu().numpy()
            rm = batch.non_tensor_batch["reward_model"]
            scores = [synthetic_token_reward(resp[i, : lens[i]], rm[i]["ground_truth"], 0) for i in range(len(batch))]
            return core.place_outcome_rewards(scores, mask)
        out = self.reward_fn(batch)
        return out.to(self.device) if torch.is_tensor(out) else out["reward_tensor"].to(self.device)

    def _dump_generations(self, batch: DataProto, scores: torch.Tensor, dump_dir: str | None, tag: str) -> None:
        """Append one JSONL line per sample (prompt / response text or ids, score, step) — reference:
        ray_trainer.py:428-486 ``_dump_generations`` / ``trainer.rollout_data_dir`` / ``trainer.validation_data_dir``."""
        if not dump_dir or self.rank != 0:
            return
        import json

        path = Path(dump_dir)
        path.mkdir(parents=True, exist_ok=True)
        tok = getattr(self, "tokenizer", None)
        prompts, resp = batch.batch["prompts"].cpu(), batch.batch["responses"].cpu()
        pm = batch.batch["attention_mask"][:, : prompts.shape[1]].bool().cpu()
        rm = batch.batch["attention_mask"][:, prompts.shape[1]:].bool().cpu()
        srcs = batch.non_tensor_batch.get("data_source")
        with open(path / f"{tag}_{self.global_steps}.jsonl", "a") as f:
            for i in range(len(batch)):
                p_ids, r_ids = prompts[i][pm[i]].tolist(), resp[i][rm[i]].tolist()
                row = {"step": self.global_steps, "score": float(scores[i]), "data_source": None if srcs is None else str(srcs[i])}
                if tok is not None:
                    row["input"], row["output"] = tok.decode(p_ids), tok.decode(r_ids)
                else:
                    row["input_ids"], row["output_ids"] = p_ids, r_ids
                f.write(json.dumps(row) + "\n")

    # ------------------------------------------------------------------ one streamed batch
    def _process_ibatch(self, ibatch: DataProto, seen: int, boundaries: list[int], metrics: dict) -> None:
        cfg, algo = self.config, self.config.algorithm
        t = self._timing
        with marked_timer("reward", t):
            self._model_reward(ibatch)
            ibatch.batch["token_level_scores"] = self._reward(ibatch)
            self._dump_generations(ibatch, ibatch.batch["token_level_scores"].sum(-1).tolist(), self.config.trainer.rollout_data_dir, "train")
        full_len = len(ibatch)
        uid_full = ibatch.non_tensor_batch["uid"]
        if algo.filter_groups.enable:
            ibatch, fm = core.filter_uninformative_groups(ibatch, algo.filter_groups.metric)
            for k, v in fm.items():
                metrics.setdefault(k, []).append(v)
        # advantages that need whole groups are computed on the full streamed batch (groups arrive together)
        group_level = algo.adv_estimator not in ("gae", "reinforce_plus_plus", "remax")
        if group_level and not algo.use_kl_in_reward:
            ibatch.batch["token_level_rewards"] = ibatch.batch["token_level_scores"]
            with marked_timer("adv", t):
                ibatch = core.compute_advantage(ibatch, algo.adv_estimator, algo.gamma, algo.lam, cfg.actor_rollout_ref.rollout.n,
                                                algo.norm_adv_by_std_in_grpo, algo)
        local = self._balance_and_shard(ibatch, metrics)
        local.meta_info["temperature"] = cfg.actor_rollout_ref.rollout.temperature
        with marked_timer("old_log_prob", t):
            bypass = cfg.algorithm.rollout_correction.bypass_mode and "rollout_log_probs" in local.batch
            if bypass:
                # rollout_correction.bypass_mode (reference: rollout_corr_helper / fully-async use_rollout_log_probs): the engine's
                # own log-probs serve as pi_old — no recompute pass; the ratio then also absorbs the engine / trainer mismatch
                local.batch["old_log_probs"] = local.batch["rollout_log_probs"].to(torch.float32)
            else:
                old = self.worker.compute_log_prob(local)
                local.batch["old_log_probs"] = old.batch["old_log_probs"]
                metrics.setdefault("actor/entropy_rollout", []).append(
                    core_algos.masked_mean(old.batch["entropys"], local.batch["response_mask"].float()).item())
            if "rollout_log_probs" in local.batch and not bypass:
                m = local.batch["response_mask"].bool()
                diff = (local.batch["rollout_log_probs"] - local.batch["old_log_probs"]).abs()[m]
                if diff.numel():
                    metrics.setdefault("training/rollout_probs_diff_mean", []).append(diff.mean().item())
                    metrics.setdefault("training/rollout_probs_diff_max", []).append(diff.max().item())
                    metrics.setdefault("training/rollout_probs_diff_std", []).append(diff.std().item() if diff.numel() > 1 else 0.0)
                    # Pearson correlation of the engine's and the trainer's token probabilities (reference:
                    # verl/utils/debug/metrics.py calculate_debug_metrics): drops when the two forward passes disagree systematically
                    a = local.batch["rollout_log_probs"][m].float().exp()
                    b = local.batch["old_log_probs"][m].float().exp()
                    if a.numel() > 1 and float(a.std()) > 0 and float(b.std()) > 0:
                        metrics.setdefault("training/rollout_actor_probs_pearson_corr", []).append(
                            float(torch.corrcoef(torch.stack([a, b]))[0, 1]))
        rc = cfg.algorithm.rollout_correction
        if (rc.rollout_is or rc.rollout_rs) and "rollout_log_probs" in local.batch:
            from ..algos.rollout_corr import compute_rollout_correction

            w, new_mask, rcm = compute_rollout_correction(local.batch["old_log_probs"], local.batch["rollout_log_probs"],
                                                          local.batch["response_mask"], rc.rollout_is, rc.rollout_is_threshold,
                                                          rc.rollout_rs, rc.rollout_rs_threshold, rc.rollout_rs_threshold_lower,
                                                          rc.rollout_token_veto_threshold)
            if w is not None:
                local.batch["rollout_is_weights"] = w
            local.batch["response_mask"] = new_mask
            for k, v in rcm.items():
                metrics.setdefault(k, []).append(v)
        if self.use_reference:
            with marked_timer("ref", t):
                local.batch["ref_log_prob"] = self.worker.compute_ref_log_prob(local).batch["ref_log_prob"]
        if self.use_critic:
            with marked_timer("values", t):
                local.batch["values"] = self.critic.compute_values(local).batch["values"]
        if not (group_level and not algo.use_kl_in_reward):
            with marked_timer("adv", t):
                ctx = getattr(self, "_shard_ctx", None)
                dp_group = getattr(self.worker, "group", None)
                if algo.use_kl_in_reward:
                    real = None
                    if ctx is not None:
                        n = len(local)
                        real = torch.from_numpy(~ctx["is_pad"][ctx["dp_rank"] * n:(ctx["dp_rank"] + 1) * n]).to(self.device)
                        if ctx["sp"] > 1:      # SP peers hold the same shard: count it once
                            real = real & bool(self.rank % ctx["sp"] == 0)
                    local, klm = core.apply_kl_penalty(local, self.kl_ctrl, algo.kl_penalty,
                                                       group=dp_group, real_rows=real,
                                                       distributed=ctx is not None and dist.is_initialized())
                    for k, v in klm.items():
                        metrics.setdefault(k, []).append(v)
                else:
                    local.batch["token_level_rewards"] = local.batch["token_level_scores"]
                if group_level and ctx is not None and dist.is_initialized():
                    # groups were split across DP ranks by the balanced sharding: gather rewards, run the estimator on the
                    # whole streamed batch, keep this rank's rows
                    local = core.compute_group_advantage_sharded(local, ctx, algo.adv_estimator, algo.gamma, algo.lam,
                                                                 cfg.actor_rollout_ref.rollout.n, algo.norm_adv_by_std_in_grpo,
                                                                 algo, dp_group)
                else:
                    local = core.compute_advantage(local, algo.adv_estimator, algo.gamma, algo.lam,
                                                   cfg.actor_rollout_ref.rollout.n, algo.norm_adv_by_std_in_grpo, algo)
        if algo.use_pf_ppo:
            # PF-PPO: resample this rank's shard by score (every rank keeps its row count, so mini-batch boundaries are unchanged)
            local = core.pf_ppo_resample(local, algo.pf_ppo.get("reweight_method", "pow"), float(algo.pf_ppo.get("weight_pow", 2.0)))
        self._step_batches.append(local)
        # ---- split at cumulative mini-batch boundaries (in global sample counts)
        end_global = seen + full_len
        cuts = [b for b in boundaries if seen < b <= end_global]
        per_rank = len(local)
        start = 0
        for b in cuts + ([end_global] if (not cuts or cuts[-1] != end_global) else []):
            frac = (b - seen) / full_len
            stop = per_rank if b == end_global else int(round(frac * per_rank))
            if stop <= start and b not in cuts:
                continue
            part = local.slice(start, stop) if stop > start else None
            is_opt = b in cuts
            is_lr = is_opt and b == boundaries[-1]
            if part is not None and len(part):
                part.meta_info = {**local.meta_info, "is_opt_step": is_opt, "is_lr_step": is_lr}
                if self.use_critic:
                    with marked_timer("update_critic", t):
                        M.append_to_dict(metrics, self.critic.update_critic_stream(part).meta_info["metrics"])
                if self.global_steps >= cfg.trainer.critic_warmup or not self.use_critic:
                    with marked_timer("update_actor", t):
                        M.append_to_dict(metrics, self.worker.update_actor_stream(part).meta_info["metrics"])
            elif is_opt:
                # boundary hit exactly at a batch edge with nothing left in this slice: step actor AND critic on the
                # accumulated grads, at the scheduled learning rate
                if self.use_critic:
                    M.append_to_dict(metrics, self.critic.optimizer_step_only(is_lr))
                if self.global_steps >= cfg.trainer.critic_warmup or not self.use_critic:
                    M.append_to_dict(metrics, self.worker.optimizer_step_only(is_lr))
            start = stop

    # ------------------------------------------------------------------ validation / checkpoint
    def _validate(self) -> dict:
        if self.val_dataset is None:
            return {}
        cfg = self.config
        bs = cfg.data.val_batch_size or len(self.val_dataset)
        scores, sources, uids = [], [], []
        val_samples: list[tuple] = []        # trainer.log_val_generations: a few (input, output, score) triples per validation
        # trainers without a colocated engine (disaggregated deployments) validate through the rollout manager like training
        # does (reference: stream_ray_trainer.py _validate -> generate_sequences on the remote rollout); rank 0 owns the client
        remote = getattr(self.worker, "rollout", None) is None
        vk = cfg.actor_rollout_ref.rollout.val_kwargs
        for i in range(0, len(self.val_dataset), bs):
            rows = [self.val_dataset[j] for j in range(i, min(i + bs, len(self.val_dataset)))]
            if remote:
                shard = rows if self.client is not None and self.rank == 0 else []
            else:
                shard = rows[self.rank::self.world] if self.world > 1 else rows
            if not shard:
                continue
            batch = DataProto.from_single_dict(collate_fn(shard)).to(self.device)
            batch.meta_info = {"validate": True, "do_sample": vk.get("do_sample", False)}
            if remote:
                temp = vk.get("temperature", 0.0) if vk.get("do_sample", False) else 0.0
                out = self.client.generate_once(batch, 1, {"temperature": temp, "top_p": vk.get("top_p", 1.0), "top_k": vk.get("top_k", -1)})
            else:
                out = self.worker.generate_sequences(batch)
            fn = self.val_reward_fn
            if fn is None:
                old, self.reward_fn = self.reward_fn, None
                r = self._reward(out)
                self.reward_fn = old
            else:
                r = fn(out)
                r = r if torch.is_tensor(r) else r["reward_tensor"]
            self._dump_generations(out, r.sum(-1).tolist(), cfg.trainer.validation_data_dir, "val")
            if cfg.trainer.log_val_generations > 0 and self.rank == 0 and len(val_samples) < cfg.trainer.log_val_generations:
                tok = getattr(self, "tokenizer", None)
                pr, rs = out.batch["prompts"].cpu(), out.batch["responses"].cpu()
                pm = out.batch["attention_mask"][:, : pr.shape[1]].bool().cpu()
                rmk = out.batch["attention_mask"][:, pr.shape[1]:].bool().cpu()
                for j in range(min(len(out), cfg.trainer.log_val_generations - len(val_samples))):
                    p_ids, r_ids = pr[j][pm[j]].tolist(), rs[j][rmk[j]].tolist()
                    val_samples.append((tok.decode(p_ids) if tok is not None else p_ids, tok.decode(r_ids) if tok is not None else r_ids,
                                        float(r[j].sum())))
            scores += r.sum(-1).tolist()
            sources += list(out.non_tensor_batch.get("data_source", ["unknown"] * len(out)))
            uids += [str(uuid.uuid4()) for _ in range(len(out))]
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (scores, sources, uids), group=self.group)
            scores = [s for g in gathered for s in g[0]]
            sources = [s for g in gathered for s in g[1]]
            uids = [s for g in gathered for s in g[2]]
        if val_samples:
            from ..utils.tracking import ValidationGenerationsLogger

            ValidationGenerationsLogger(cfg.trainer.project_name, cfg.trainer.experiment_name).log(
                cfg.trainer.logger, val_samples, self.global_steps)
        res = M.process_validation_metrics(sources, uids, {"reward": scores})
        out = {}
        for ds, vars_ in res.items():
            for var, ms in vars_.items():
                for mname, v in ms.items():
                    out[f"val-core/{ds}/{var}/{mname}"] = v
        return out

    def _ckpt_root(self) -> Path:
        return Path(self.config.trainer.default_local_dir)

    def _save_checkpoint(self) -> None:
        root = self._ckpt_root() / f"global_step_{self.global_steps}"
        t = self.config.trainer
        # rotation happens once per step, below, on whole global_step_* directories (a per-role rotation would strand
        # data.pt / the sibling role of a deleted step)
        self.worker.save_checkpoint(str(root / "actor"), None, self.global_steps, None)
        if self.use_critic:
            self.critic.save_checkpoint(str(root / "critic"), None, self.global_steps, None)
        if self.rank == 0:
            torch.save({"sampler": self.sampler.state_dict() if self.sampler is not None else None,
                        "global_steps": self.global_steps,
                        "kl_ctrl": getattr(self.kl_ctrl, "value", None)}, root / "data.pt")
            write_latest(self._ckpt_root(), self.global_steps)
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)
        if t.default_hdfs_dir and self.rank == 0:
            # second copy of the step (HDFS through the `hdfs` CLI, or any mounted directory): reference trainer.default_hdfs_dir
            from ..utils.fs import mirror_checkpoint

            try:
                mirror_checkpoint(root, t.default_hdfs_dir)
            except Exception as e:                  # noqa: BLE001  (a failed mirror must not stop training)
                print(f"[trainer] checkpoint mirror to {t.default_hdfs_dir!r} failed: {e}", flush=True)
        keeps = [k for k in (t.max_actor_ckpt_to_keep, t.max_critic_ckpt_to_keep if self.use_critic else None) if k]
        if keeps and self.rank == 0:
            from ..utils.checkpoint import rotate_checkpoints

            rotate_checkpoints(root, max(keeps))

    def _load_checkpoint(self) -> int:
        t = self.config.trainer
        if t.resume_mode == "disable":
            return 0
        path = t.resume_from_path if t.resume_mode == "resume_path" else find_latest_ckpt_path(self._ckpt_root())
        if not path:
            return 0
        path = Path(path)
        self.worker.load_checkpoint(str(path / "actor"))
        if self.use_critic and (path / "critic").exists():
            self.critic.load_checkpoint(str(path / "critic"))
        data = path / "data.pt"
        if data.exists():
            sd = torch.load(data, weights_only=False)
            self.global_steps = sd.get("global_steps", int(path.name.split("_")[-1]))
            if self.sampler is not None and sd.get("sampler"):
                self.sampler.load_state_dict(sd["sampler"])
            if self.kl_ctrl is not None and sd.get("kl_ctrl") is not None:
                self.kl_ctrl.value = sd["kl_ctrl"]
        else:
            self.global_steps = int(path.name.split("_")[-1])
        if self.rank == 0:
            print(f"[trainer] resumed from {path} at global step {self.global_steps}", flush=True)
        return self.global_steps

    def _should_save(self, is_last: bool, step_time: float) -> bool:
        t = self.config.trainer
        if t.save_freq > 0 and (is_last or self.global_steps % t.save_freq == 0):
            return True
        # ESI (spot-instance expiry) aware: force a checkpoint when the lease is about to end
        if self._esi_deadline > 0 and time.time() + step_time + t.esi_redundant_time >= self._esi_deadline:
            return True
        return False

    # ------------------------------------------------------------------ main loop
    def fit(self) -> dict:
        cfg = self.config
        resumed_steps = self._load_checkpoint()
        last_metrics: dict = {}
        if cfg.trainer.val_before_train and self.val_dataset is not None:
            vm = self._validate()
            if self.rank == 0 and self.tracking:
                self.tracking.log(vm, self.global_steps)
            if cfg.trainer.val_only:
                return vm
        if self.client is not None and self.rank == 0:
            self.client.update_weight_version()
        self.worker.update_weight_remote()
        prof = cfg.global_profiler
        one_off = bool(cfg.trainer.one_step_off and self.client is not None)
        batches = iter(self._iter_batches())
        if resumed_steps and not (self.sampler is not None and hasattr(self.sampler, "state_dict")):
            # no stateful sampler to restore the data position from: skip the batches the checkpointed steps consumed
            for _ in range(resumed_steps):
                if next(batches, None) is None:
                    break
        lookahead = next(batches, None)
        self._prefetched = None
        while lookahead is not None:
            prompts, lookahead = lookahead, next(batches, None)
            if self.global_steps >= self.total_training_steps:
                break
            self.global_steps += 1
            handle_next = None
            if one_off and lookahead is not None and self.global_steps < self.total_training_steps:
                # rollout GPUs start on the NEXT batch now (policy = weights of this step), while we train on the current one
                lookahead.non_tensor_batch["uid"] = core.uid_array(len(lookahead), 1, prefix=f"{self.global_steps + 1}-")
                if self.rank == 0:
                    r_ = cfg.actor_rollout_ref.rollout
                    handle_next = self.client.submit(lookahead, {"temperature": r_.temperature, "top_p": r_.top_p, "top_k": r_.top_k})
            do_profile = bool(prof.tool and prof.steps and self.global_steps in prof.steps)
            # global_profiler.profile_continuous_steps: consecutive profiled steps share ONE capture (start at the first, stop
            # after the last) instead of one capture per step
            cont = bool(prof.profile_continuous_steps)
            if do_profile and not (cont and getattr(self, "_profiling", False)):
                self.worker.start_profile()
                self._profiling = True
            self._timing = {}
            self._step_batches = []
            metrics: dict = {}
            t_step = time.time()
            n = cfg.actor_rollout_ref.rollout.n
            if "uid" not in prompts.non_tensor_batch:
                prompts.non_tensor_batch["uid"] = core.uid_array(len(prompts), 1, prefix=f"{self.global_steps}-")
            total = len(prompts) * n
            boundaries = self._mini_batch_boundaries(total)
            seen = 0
            for ibatch in self._stream(prompts):
                self._process_ibatch(ibatch, seen, boundaries, metrics)
                seen += len(ibatch)
            if one_off:
                # the next batch must be complete before the weights under the rollout engines are replaced
                with marked_timer("wait_next_rollout", self._timing):
                    if self.rank == 0 and handle_next is not None:
                        self._prefetched = list(self.client.collect(handle_next))
                    flag = _bcast_obj(handle_next is not None if self.rank == 0 else None, 0, self.group)
                    if flag and self.rank != 0:
                        self._prefetched = []           # non-empty marker: the items come through the broadcast in _stream
            with marked_timer("update_weight", self._timing):
                if self.client is not None and self.rank == 0:
                    self.client.update_weight_version()
                self.worker.update_weight_remote()
            ruf = cfg.trainer.ref_update_freq
            if self.use_reference and ruf and ruf > 0 and self.global_steps % ruf == 0:
                self.worker.update_reference_policy()             # self-play: the reference follows the actor every k steps
            step_time = time.time() - t_step
            self._timing["step"] = step_time
            is_last = self.global_steps >= self.total_training_steps
            if self.val_dataset is not None and cfg.trainer.test_freq > 0 and (is_last or self.global_steps % cfg.trainer.test_freq == 0):
                with marked_timer("testing", self._timing):
                    metrics.update(self._validate())
            if self._should_save(is_last, step_time):
                with marked_timer("save_checkpoint", self._timing):
                    self._save_checkpoint()
            if do_profile and not (cont and prof.steps and (self.global_steps + 1) in prof.steps and not is_last):
                self.worker.stop_profile()
                self._profiling = False
            # ---- metrics
            full = DataProto.concat(self._step_batches)
            self._data_feedback(full)
            out = M.reduce_metrics(metrics)
            out.update(M.compute_data_metrics(full, use_critic=self.use_critic))
            out.update(M.compute_timing_metrics(full, self._timing))
            ntok = torch.tensor([float(full.batch["attention_mask"].sum().item())], device=self.device)
            if self.world > 1:
                dist.all_reduce(ntok, group=self.group)
            full.meta_info["global_token_num"] = [int(ntok.item())]
            out.update(M.compute_throughout_metrics(full, self._timing, self.world))
            out["training/global_step"] = self.global_steps
            out["perf/samples_per_sec"] = total / step_time
            if self.client is not None and self.rank == 0:
                try:
                    fb = self.client.update_metrics(step_time, self._timing.get("gen", 0.0), out["perf/throughput_all_gpus"])
                    out["training/num_rollout_instances"] = fb.get("new_num_rollout_instances", 0)
                    out["training/new_max_gen_s"] = fb.get("new_max_gen_s", 0)
                except Exception as e:  # the feedback loop must never kill training
                    print(f"[trainer] update_metrics failed: {e}")
            if self.rank == 0 and self.tracking:
                self.tracking.log(out, self.global_steps)
            last_metrics = out
            cb = getattr(self, "step_callback", None)
            if cb is not None:                   # benchmark / monitoring hook: called on EVERY trainer rank after each step
                cb(self, self.global_steps, out)
        return last_metrics
