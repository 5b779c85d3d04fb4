"""Data-parallel PPO actor and critic compute: packed (remove-padding) micro-batched forward /
backward on the ZeRO-3 engine with the *streaming* update semantics of the reference.

Reference behaviour reproduced:
* ``_forward_micro_batch`` — unpad, packed forward, temperature-scaled log-prob + entropy of the
  response tokens (3rdparty/verl/verl/workers/actor/dp_actor.py:86-273); here the LM head, softmax,
  log-prob and entropy are ONE fused tcgen05 kernel and only response positions are projected.
* ``update_policy_stream`` — input is at most one mini-batch; it is split into micro-batches (fixed size
  or dynamic token budget), grads accumulate (they persist across calls), the optimizer steps only when
  ``is_opt_step``; ``ppo_epochs`` must be 1 (rlboost/verl_stream/workers/actor/stream_dp_actor.py:85-231).
* critic equivalent (rlboost/verl_stream/workers/critic/stream_dp_critic.py:68-141).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from ..algos import core_algos
from ..models import transformer as tf
from ..parallel.zero3 import Zero3Engine
from ..protocol import DataProto
from ..utils.seqlen_balancing import rearrange_micro_batches


@dataclass
class ActorConfig:
    ppo_mini_batch_size: int = 256            # per-GPU after normalisation
    ppo_micro_batch_size_per_gpu: int | None = 8
    use_dynamic_bsz: bool = False
    ppo_max_token_len_per_gpu: int = 16384
    clip_ratio: float = 0.2
    clip_ratio_low: float | None = None
    clip_ratio_high: float | None = None
    clip_ratio_c: float = 3.0
    loss_agg_mode: str = "token-mean"
    entropy_coeff: float = 0.0
    router_aux_loss_coef: float = 0.0         # MoE policies: weight of the router load-balancing loss
    use_kl_loss: bool = False
    kl_loss_coef: float = 0.001
    kl_loss_type: str = "low_var_kl"
    ppo_epochs: int = 1
    policy_loss_mode: str = "vanilla"
    policy_loss: dict = field(default_factory=dict)
    grad_clip: float = 1.0
    # critic
    cliprange_value: float = 0.5

    def get(self, k, default=None):
        return getattr(self, k, default)


def _select_response_rows(attention_mask: torch.Tensor, response_length: int):
    """Packed-row bookkeeping for a [b, S] batch whose last ``R`` columns are the response.

    Returns (inp, sel, labels, valid) where ``sel`` are packed-row indices of the positions whose NEXT
    token is a response token (columns S-R-1 .. S-2), ``labels`` the next-token ids for those rows and
    ``valid`` [b, R] marks response columns that are real tokens."""
    raise NotImplementedError


class PackedBatch:
    """Remove-padding view of a micro-batch plus the index maps back to [b, R]."""

    def __init__(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, position_ids: torch.Tensor | None,
                 response_length: int):
        b, s = input_ids.shape
        r = response_length
        self.b, self.s, self.r = b, s, r
        self.inp, self.idx = tf.pack_padded(input_ids, attention_mask, position_ids)
        mask = attention_mask.bool()
        # packed row index of every (row, col); -1 where padded
        flat_pos = torch.full((b * s,), -1, dtype=torch.long, device=input_ids.device)
        flat_pos[self.idx] = torch.arange(self.idx.numel(), device=input_ids.device)
        grid = flat_pos.view(b, s)
        src = grid[:, s - r - 1: s - 1]                  # hidden at col c-1 predicts token at col c
        tgt_valid = mask[:, s - r:] & (src >= 0)
        self.sel = src[tgt_valid]                        # packed rows to project through the LM head
        self.labels = input_ids[:, s - r:][tgt_valid]
        self.valid = tgt_valid                           # [b, R]
        self._grid, self._mask = grid, mask

    def prompt_mean(self, hidden: torch.Tensor) -> torch.Tensor:
        """Mean hidden state over every sample's prompt tokens (columns before the response, the last prompt token excluded —
        its hidden state already predicts the first response token) -> [b, H] fp32.  FlowRL's log-Z head reads this."""
        cols = self.s - self.r - 1
        pm = self._mask[:, :cols]
        rows = self._grid[:, :cols][pm]
        owner = torch.arange(self.b, device=hidden.device).unsqueeze(1).expand(self.b, cols)[pm]
        sums = torch.zeros(self.b, hidden.shape[-1], device=hidden.device, dtype=torch.float32)
        sums = sums.index_add(0, owner, hidden.index_select(0, rows).float())
        return sums / pm.sum(-1, keepdim=True).clamp(min=1).float()

    def scatter(self, values: torch.Tensor, fill: float = 0.0) -> torch.Tensor:
        out = values.new_full((self.b, self.r), fill)
        out[self.valid] = values
        return out


def attach_multi_modal(inp: "tf.PackedInput", mb: DataProto) -> None:
    """Vision-language micro-batches: hand the rows' image patches (row order = placeholder order in the packed stream) to the
    engine, which runs its tower unit (reference: dp_actor.py multi_modal_inputs concat, verl/workers/actor/dp_actor.py:100-125)."""
    mm = mb.non_tensor_batch.get("multi_modal_inputs") if mb.non_tensor_batch else None
    if mm is None:
        return
    from ..models.vision import collate_multi_modal

    pix, grids = collate_multi_modal(list(mm))
    if pix is not None:
        inp.pixel_values, inp.grid_thw = pix.to(inp.input_ids.device), grids


class DataParallelPPOActor:
    def __init__(self, engine: Zero3Engine, config: ActorConfig, dp_group=None):
        self.engine = engine
        self.config = config
        self.dp_group = dp_group
        if config.ppo_epochs != 1:
            raise ValueError("streaming updates require ppo_epochs == 1")
        if engine.trainable:
            engine.max_grad_norm = config.grad_clip
            engine.zero_grad()

    # ------------------------------------------------------------------ micro-batching
    def _micro_batches(self, data: DataProto, micro_bsz: int | None, max_tokens: int | None, dynamic: bool):
        n = len(data)
        if dynamic:
            lens = data.batch["attention_mask"].sum(-1).tolist()
            parts = rearrange_micro_batches([int(x) for x in lens], max_tokens, dp_group=self.dp_group)
            return [(data.select_idxs(torch.tensor(p, dtype=torch.long)), p) for p in parts]
        micro_bsz = micro_bsz or n
        out = []
        for i in range(0, n, micro_bsz):
            idx = list(range(i, min(i + micro_bsz, n)))
            out.append((data.slice(i, min(i + micro_bsz, n)), idx))
        return out

    def _forward_micro_batch(self, mb: DataProto, temperature: float, calculate_entropy: bool = True, return_prompt_hidden: bool = False):
        """-> (entropy [b,R], log_probs [b,R]) for the response tokens (+ the mean prompt hidden state [b,H] on request)."""
        ids = mb.batch["input_ids"]
        r = mb.batch["responses"].shape[-1]
        pb = PackedBatch(ids, mb.batch["attention_mask"], mb.batch.get("position_ids"), r)
        attach_multi_modal(pb.inp, mb)
        hidden = self.engine.forward_hidden(pb.inp)
        h_sel = hidden.index_select(0, pb.sel)
        lp, ent = self.engine.head_lp_runner(float(temperature), h_sel, pb.labels)
        if return_prompt_hidden:
            return pb.scatter(ent), pb.scatter(lp), pb.prompt_mean(hidden)
        return pb.scatter(ent), pb.scatter(lp)

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def compute_log_prob(self, data: DataProto, calculate_entropy: bool = True):
        """-> (log_probs [bs, R], entropys [bs, R]) in the original sample order."""
        meta = data.meta_info
        temperature = meta.get("temperature", 1.0)
        dynamic = meta.get("use_dynamic_bsz", self.config.use_dynamic_bsz)
        mbs = self._micro_batches(data, meta.get("micro_batch_size"), meta.get("max_token_len"), dynamic)
        n, r = len(data), data.batch["responses"].shape[-1]
        dev = data.batch["input_ids"].device
        lps = torch.zeros(n, r, device=dev)
        ents = torch.zeros(n, r, device=dev)
        for mb, idx in mbs:
            ent, lp = self._forward_micro_batch(mb, temperature, calculate_entropy)
            ix = torch.tensor(idx, dtype=torch.long, device=dev)
            lps[ix] = lp.float()
            ents[ix] = ent.float()
        return lps, ents

    # ------------------------------------------------------------------ training
    def update_policy_stream(self, data: DataProto, is_opt_step: bool = True, lr: float | None = None) -> dict:
        cfg = self.config
        meta = data.meta_info
        temperature = meta.get("temperature", 1.0)
        metrics: dict[str, list] = {}
        n = len(data)
        mbs = self._micro_batches(data, cfg.ppo_micro_batch_size_per_gpu, cfg.ppo_max_token_len_per_gpu,
                                  cfg.use_dynamic_bsz)
        grad_accum = max(1, cfg.ppo_mini_batch_size // max(1, (cfg.ppo_micro_batch_size_per_gpu or n)))
        loss_fn = core_algos.get_policy_loss_fn(cfg.policy_loss_mode)
        mbs = list(mbs)
        for mi, (mb, idx) in enumerate(mbs):
            b = mb.batch
            mask = b["response_mask"].float()
            scale = (len(mb) / cfg.ppo_mini_batch_size) if cfg.use_dynamic_bsz else 1.0 / grad_accum
            ent, lp = self._forward_micro_batch(mb, temperature, calculate_entropy=cfg.entropy_coeff != 0)
            pg_loss, pg_metrics = loss_fn(old_log_prob=b["old_log_probs"], log_prob=lp, advantages=b["advantages"],
                                          response_mask=mask, loss_agg_mode=cfg.loss_agg_mode, config=cfg,
                                          rollout_is_weights=b.get("rollout_is_weights"))
            loss = pg_loss
            m = {"actor/pg_loss": pg_loss.detach().item(), **pg_metrics}
            if cfg.entropy_coeff != 0:
                e_loss = core_algos.agg_loss(ent, mask, cfg.loss_agg_mode)
                loss = loss - cfg.entropy_coeff * e_loss
            m["actor/entropy"] = core_algos.masked_mean(ent.detach(), mask).item()
            if cfg.use_kl_loss:
                kld = core_algos.kl_penalty(lp, b["ref_log_prob"], cfg.kl_loss_type)
                kl_loss = core_algos.agg_loss(kld, mask, cfg.loss_agg_mode)
                loss = loss + cfg.kl_loss_coef * kl_loss
                m["actor/kl_loss"] = kl_loss.detach().item()
                m["actor/kl_coef"] = cfg.kl_loss_coef
            if getattr(self.engine, "moe_aux", False) and self.engine.last_aux_loss is not None:
                aux = self.engine.last_aux_loss        # router load-balancing term of the MoE blocks (this micro-batch's forward)
                loss = loss + cfg.router_aux_loss_coef * aux
                m["actor/router_aux_loss"] = aux.detach().item()
            if is_opt_step and mi == len(mbs) - 1:
                self.engine.arm_norm_fold()         # the last reduce-scatter of the mini-batch also produces the grad norm
            (loss * scale).backward()
            prof = getattr(self, "profiler", None)
            if prof is not None:
                prof.step()                         # torch-profiler step window counts micro-batches (tool_config.torch.step_*)
            for k, v in m.items():
                metrics.setdefault(k, []).append(v)
        if is_opt_step:
            st = self.engine.step(lr)
            metrics.setdefault("actor/grad_norm", []).append(st["grad_norm"])
            if st["skipped"]:
                metrics.setdefault("actor/skipped_steps", []).append(1.0)
        return metrics


class DataParallelPPOCritic:
    """Value model: same backbone, scalar head (reference: verl/workers/critic/dp_critic.py:53-261)."""

    def __init__(self, engine: Zero3Engine, config: ActorConfig, dp_group=None):
        self.engine, self.config, self.dp_group = engine, config, dp_group
        self._mb = DataParallelPPOActor._micro_batches.__get__(self)
        if engine.trainable:
            engine.max_grad_norm = config.grad_clip
            engine.zero_grad()

    def _forward_micro_batch(self, mb: DataProto) -> torch.Tensor:
        ids = mb.batch["input_ids"]
        r = mb.batch["responses"].shape[-1]
        pb = PackedBatch(ids, mb.batch["attention_mask"], mb.batch.get("position_ids"), r)
        attach_multi_modal(pb.inp, mb)
        hidden = self.engine.forward_hidden(pb.inp)
        v = self.engine.head_value_runner(None, hidden.index_select(0, pb.sel))
        return pb.scatter(v)                 # value of the state BEFORE each response token

    @torch.no_grad()
    def compute_values(self, data: DataProto) -> torch.Tensor:
        meta = data.meta_info
        dynamic = meta.get("use_dynamic_bsz", self.config.use_dynamic_bsz)
        mbs = self._mb(data, meta.get("micro_batch_size"), meta.get("max_token_len"), dynamic)
        n, r = len(data), data.batch["responses"].shape[-1]
        dev = data.batch["input_ids"].device
        out = torch.zeros(n, r, device=dev)
        for mb, idx in mbs:
            out[torch.tensor(idx, dtype=torch.long, device=dev)] = self._forward_micro_batch(mb).float()
        mask = data.batch.get("response_mask")
        return out * mask if mask is not None else out

    def update_critic_stream(self, data: DataProto, is_opt_step: bool = True, lr: float | None = None) -> dict:
        cfg = self.config
        metrics: dict[str, list] = {}
        n = len(data)
        mbs = self._mb(data, cfg.ppo_micro_batch_size_per_gpu, cfg.ppo_max_token_len_per_gpu, cfg.use_dynamic_bsz)
        grad_accum = max(1, cfg.ppo_mini_batch_size // max(1, (cfg.ppo_micro_batch_size_per_gpu or n)))
        mbs = list(mbs)
        for mi, (mb, idx) in enumerate(mbs):
            b = mb.batch
            mask = b["response_mask"].float()
            scale = (len(mb) / cfg.ppo_mini_batch_size) if cfg.use_dynamic_bsz else 1.0 / grad_accum
            vpred = self._forward_micro_batch(mb)
            vf_loss, clipfrac = core_algos.compute_value_loss(vpred, b["returns"], b["values"], mask,
                                                              cfg.cliprange_value, cfg.loss_agg_mode)
            if is_opt_step and mi == len(mbs) - 1:
                self.engine.arm_norm_fold()
            (vf_loss * scale).backward()
            metrics.setdefault("critic/vf_loss", []).append(vf_loss.detach().item())
            metrics.setdefault("critic/vf_clipfrac", []).append(clipfrac.detach().item())
            metrics.setdefault("critic/vpred_mean", []).append(core_algos.masked_mean(vpred.detach(), mask).item())
        if is_opt_step:
            st = self.engine.step(lr)
            metrics.setdefault("critic/grad_norm", []).append(st["grad_norm"])
        return metrics


class FlowRLActor(DataParallelPPOActor):
    """Actor of the FlowRL recipe (``policy_loss.loss_mode = flowrl``; reference: recipe/flowrl/flowrl_actor.py): the update
    minimises the trajectory-balance residual with a learned log-partition head on the mean prompt hidden state.  The head is a
    small replicated fp32 module with its own AdamW; its gradients are averaged over the data-parallel group at the optimizer
    step, the policy's gradients flow through the ZeRO engine as usual."""

    def __init__(self, engine: Zero3Engine, config: ActorConfig, dp_group=None, seed: int = 0):
        super().__init__(engine, config, dp_group)
        from ..algos.flowrl import LogZHead

        pl = config.policy_loss or {}
        gen = torch.Generator().manual_seed(seed + 4242)
        cpu_state = torch.random.get_rng_state()
        torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=gen)))          # identical init on every rank
        self.log_z = LogZHead(engine.cfg.hidden_size, int(pl.get("flowrl_logz_layers", 3)), float(pl.get("flowrl_logz_dropout", 0.1)))
        torch.random.set_rng_state(cpu_state)
        self.log_z.to(engine.device)
        self.beta = float(pl.get("flowrl_beta", 15.0))
        self.log_z_opt = torch.optim.AdamW(self.log_z.parameters(), lr=float(pl.get("flowrl_logz_lr", 1e-4)), weight_decay=0.0)

    def update_policy_stream(self, data: DataProto, is_opt_step: bool = True, lr: float | None = None) -> dict:
        from ..algos.flowrl import flowrl_objective

        cfg = self.config
        if "ref_log_prob" not in data.batch:
            raise ValueError("FlowRL needs the reference log-probs (the recipe enables the reference policy)")
        temperature = data.meta_info.get("temperature", 1.0)
        metrics: dict[str, list] = {}
        n = len(data)
        mbs = list(self._micro_batches(data, cfg.ppo_micro_batch_size_per_gpu, cfg.ppo_max_token_len_per_gpu, cfg.use_dynamic_bsz))
        grad_accum = max(1, cfg.ppo_mini_batch_size // max(1, (cfg.ppo_micro_batch_size_per_gpu or n)))
        self.log_z.train()
        for mi, (mb, idx) in enumerate(mbs):
            b = mb.batch
            scale = (len(mb) / cfg.ppo_mini_batch_size) if cfg.use_dynamic_bsz else 1.0 / grad_accum
            _, lp, pm = self._forward_micro_batch(mb, temperature, calculate_entropy=False, return_prompt_hidden=True)
            loss, m = flowrl_objective(lp, b["ref_log_prob"], b["old_log_probs"], self.log_z(pm), b["advantages"], b["response_mask"],
                                       cfg.clip_ratio_low, cfg.clip_ratio_high, self.beta)
            if is_opt_step and mi == len(mbs) - 1:
                self.engine.arm_norm_fold()
            (loss * scale).backward()
            prof = getattr(self, "profiler", None)
            if prof is not None:
                prof.step()
            for k, v in m.items():
                metrics.setdefault(k, []).append(v)
        if is_opt_step:
            st = self.engine.step(lr)
            metrics.setdefault("actor/grad_norm", []).append(st["grad_norm"])
            if st["skipped"]:
                metrics.setdefault("actor/skipped_steps", []).append(1.0)
            import torch.distributed as dist

            params = [p for p in self.log_z.parameters() if p.grad is not None]
            if dist.is_initialized() and dist.get_world_size(self.dp_group) > 1:
                for p in params:
                    dist.all_reduce(p.grad, group=self.dp_group)
                    p.grad /= dist.get_world_size(self.dp_group)
            if not st["skipped"]:
                torch.nn.utils.clip_grad_norm_(params, cfg.grad_clip if cfg.grad_clip > 0 else 1e9)
                self.log_z_opt.step()
            self.log_z_opt.zero_grad(set_to_none=True)
        return metrics

    # the head travels in the worker's checkpoint "extra" dict
    def extra_state(self) -> dict:
        return {"log_z": {k: v.detach().cpu() for k, v in self.log_z.state_dict().items()}, "log_z_opt": self.log_z_opt.state_dict()}

    def load_extra_state(self, st: dict | None) -> None:
        if st and "log_z" in st:
            self.log_z.load_state_dict(st["log_z"])
            if st.get("log_z_opt"):
                self.log_z_opt.load_state_dict(st["log_z_opt"])
