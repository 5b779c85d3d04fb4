"""Pipeline-parallel training engine (dp x pp) with the 1F1B schedule.

The decoder blocks are split into ``pp`` contiguous stages, one per rank of ``pp_group``; stage 0 also owns the embedding, the last
stage the final norm + LM head (fused linear-cross-entropy).  A step runs ``m`` micro-batches through the non-interleaved
one-forward-one-backward schedule: ``pp - stage - 1`` warm-up forwards, then alternating forward / backward, then the cool-down
backwards — at most ``pp - stage`` micro-batches hold activations on a stage.  Neighbour exchanges that happen at the same point of
the schedule (send activation + receive gradient, send gradient + receive next activation) are issued as ONE
``batch_isend_irecv`` so the two directions cannot dead-lock on a stream-ordered transport.

Capability parity: the reference's Megatron back-end pipeline parallelism (``pipeline_model_parallel_size``,
3rdparty/verl/verl/utils/megatron/pipeline_parallel.py:22-49, workers/megatron_workers.py forward_backward_func / 1F1B schedule of
megatron.core.pipeline_parallel.schedules).  ``InterleavedPipelineEngine`` (below) adds virtual stages; tensor parallelism inside a stage is the
separate ``TPTrainEngine``.  On one NVSwitch box ZeRO / TP are the faster layouts — this engine is for models whose optimizer state has to
be split by depth, and for multi-node layouts where only neighbour traffic should cross the slow links.
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import torch
import torch.distributed as dist

from .. import ops
from ..models import transformer as tf
from ..models.config import ModelConfig
from ..ops import reference as ref


def stage_layers(num_layers: int, pp: int) -> list[tuple[int, int]]:
    """Contiguous, balanced [lo, hi) block ranges; earlier stages take the remainder (the last stage also runs the head)."""
    base, rem = divmod(num_layers, pp)
    out, lo = [], 0
    for s in range(pp):
        n = base + (1 if s < rem else 0)
        out.append((lo, lo + n))
        lo += n
    return out


class PipelineEngine:
    def __init__(self, cfg: ModelConfig, device, pp_group, dp_group=None, lr: float = 1e-6, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.01, max_grad_norm: float = 1.0, seed: int = 0, dtype=None):
        if cfg.family != "llama":
            raise NotImplementedError("pipeline training is implemented for the llama / qwen block family")
        self.cfg, self.device = cfg, torch.device(device)
        self.pp_group, self.dp_group = pp_group, dp_group
        self.pp = dist.get_world_size(pp_group)
        self.stage = dist.get_rank(pp_group)
        if cfg.num_layers < self.pp:
            raise ValueError(f"{cfg.num_layers} layers cannot fill {self.pp} pipeline stages")
        self.dp = dist.get_world_size(dp_group) if dp_group is not None else 1
        self.first, self.last = self.stage == 0, self.stage == self.pp - 1
        self.lo, self.hi = stage_layers(cfg.num_layers, self.pp)[self.stage]
        ranks = dist.get_process_group_ranks(pp_group)
        self.prev_rank = ranks[self.stage - 1] if not self.first else None
        self.next_rank = ranks[self.stage + 1] if not self.last else None
        self.dtype = dtype or (torch.bfloat16 if self.device.type == "cuda" else torch.float32)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.opt_step = 0
        self.peak_in_flight = 0
        full = tf.TransformerWeights(cfg, device="cpu", dtype=torch.float32, init=True, seed=seed)     # same init as the other engines
        self.load_full(full.embed, full.layers, full.head)

    # ------------------------------------------------------------------ parameters
    def load_full(self, embed: dict, layers: list[dict], head: dict) -> None:
        """Keep this stage's slice of a full model (our naming): its blocks, the embedding on stage 0, norm + head on the last."""
        cfg = self.cfg
        named: dict[str, torch.Tensor] = {}
        if self.first or (self.last and cfg.tie_word_embeddings):
            named["embed"] = embed["embed"].clone()
        if self.last:
            named["norm_w"] = head["norm_w"].clone()
            if not cfg.tie_word_embeddings:
                named["lm_head"] = head["lm_head"].clone()
        for i in range(self.lo, self.hi):
            for k, v in layers[i].items():
                named[f"layers.{i}.{k}"] = v.clone()
        self.master = {k: v.to(self.device, torch.float32) for k, v in named.items()}
        self.params = {k: v.to(self.dtype).clone().requires_grad_(True) for k, v in self.master.items()}
        self.exp_avg = {k: torch.zeros_like(v) for k, v in self.master.items()}
        self.exp_avg_sq = {k: torch.zeros_like(v) for k, v in self.master.items()}

    def layer(self, i: int) -> dict:
        pre = f"layers.{i}."
        return {k[len(pre):]: v for k, v in self.params.items() if k.startswith(pre)}

    def num_local_params(self) -> int:
        return sum(v.numel() for v in self.master.values())

    # ------------------------------------------------------------------ neighbour exchange
    def _exchange(self, send_next=None, send_prev=None, recv_next: tuple | None = None, recv_prev: tuple | None = None):
        """One grouped exchange with the neighbours; returns (from_prev, from_next)."""
        ops_, got_prev, got_next = [], None, None
        if send_prev is not None:
            ops_.append(dist.P2POp(dist.isend, send_prev.contiguous(), self.prev_rank))
        if recv_prev is not None:
            got_prev = torch.empty(recv_prev, device=self.device, dtype=self.dtype)
            ops_.append(dist.P2POp(dist.irecv, got_prev, self.prev_rank))
        if send_next is not None:
            ops_.append(dist.P2POp(dist.isend, send_next.contiguous(), self.next_rank))
        if recv_next is not None:
            got_next = torch.empty(recv_next, device=self.device, dtype=self.dtype)
            ops_.append(dist.P2POp(dist.irecv, got_next, self.next_rank))
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()
        return got_prev, got_next

    # ------------------------------------------------------------------ stage compute
    def _stage_forward(self, inp: tf.PackedInput, x_in: torch.Tensor | None) -> torch.Tensor:
        if self.first:
            x = torch.nn.functional.embedding(inp.input_ids, self.params["embed"])
        else:
            x = x_in
        for i in range(self.lo, self.hi):
            x = tf.layer_forward(self.cfg, self.layer(i), x, inp)
        return x

    def _head(self, x: torch.Tensor, labels: torch.Tensor, temperature: float):
        h = ops.rmsnorm(x, self.params["norm_w"], self.cfg.rms_norm_eps)
        w = self.params["embed"] if self.cfg.tie_word_embeddings else self.params["lm_head"]
        return ops.linear_cross_entropy(h, w, labels, float(temperature))

    def _shape(self, inp: tf.PackedInput) -> tuple[int, int]:
        return (int(inp.input_ids.shape[0]), self.cfg.hidden_size)

    # ------------------------------------------------------------------ 1F1B
    def train_batch(self, micro_batches: Sequence[tuple[tf.PackedInput, torch.Tensor]],
                    loss_fn: Callable[[torch.Tensor, torch.Tensor, int], torch.Tensor], temperature: float = 1.0) -> list[float]:
        """Forward + backward of every micro-batch ``(inp, labels)`` (the same list on every stage of this pipeline).
        ``loss_fn(log_probs, entropy, k)`` returns micro-batch k's (already weighted) scalar loss; gradients accumulate into the
        stage's parameters.  Returns the losses on the last stage ([] elsewhere)."""
        n = len(micro_batches)
        warm = min(self.pp - self.stage - 1, n)
        steady = n - warm
        held: list[tuple[torch.Tensor | None, torch.Tensor]] = []       # FIFO of (stage input, stage output or loss)
        losses: list[float] = []
        self.peak_in_flight = 0

        def forward(k: int, x_in: torch.Tensor | None) -> torch.Tensor:
            inp, labels = micro_batches[k]
            if x_in is not None:
                x_in.requires_grad_(True)
            out = self._stage_forward(inp, x_in)
            if self.last:
                lp, ent = self._head(out, labels, temperature)
                out = loss_fn(lp, ent, k)
                losses.append(float(out.detach()))
            held.append((x_in, out))
            self.peak_in_flight = max(self.peak_in_flight, len(held))
            return out

        def backward(grad_out: torch.Tensor | None) -> torch.Tensor | None:
            x_in, out = held.pop(0)
            if self.last:
                out.backward()
            else:
                torch.autograd.backward(out, grad_out)
            return None if x_in is None else x_in.grad

        shp = [self._shape(mb[0]) for mb in micro_batches]
        # ---- warm-up forwards
        for k in range(warm):
            x_in, _ = self._exchange(recv_prev=None if self.first else shp[k])
            out = forward(k, x_in)
            self._exchange(send_next=out.detach())                    # never the last stage here (warm == 0 there)
        # ---- steady state
        x_in = None
        if steady > 0 and not self.first:
            x_in, _ = self._exchange(recv_prev=shp[warm])
        for i in range(steady):
            k = warm + i
            out = forward(k, x_in)
            grad_out = None
            if not self.last:
                _, grad_out = self._exchange(send_next=out.detach(), recv_next=shp[i])     # gradient of micro-batch i comes back
            grad_in = backward(grad_out)
            more = i + 1 < steady
            if self.first:
                x_in = None
            elif more:
                x_in, _ = self._exchange(send_prev=grad_in, recv_prev=shp[k + 1])
            else:
                self._exchange(send_prev=grad_in)
        # ---- cool-down backwards
        for i in range(steady, n):
            _, grad_out = self._exchange(recv_next=shp[i])
            grad_in = backward(grad_out)
            if not self.first:
                self._exchange(send_prev=grad_in)
        return losses

    @torch.no_grad()
    def forward_only(self, micro_batches: Sequence[tuple[tf.PackedInput, torch.Tensor]], temperature: float = 1.0,
                     broadcast: bool = True) -> list[tuple[torch.Tensor, torch.Tensor]]:
        """(log-prob, entropy) of the labels per micro-batch; computed on the last stage and (optionally) broadcast to the pipeline."""
        out: list = []
        for k, (inp, labels) in enumerate(micro_batches):
            x_in, _ = self._exchange(recv_prev=None if self.first else self._shape(inp))
            x = self._stage_forward(inp, x_in)
            if self.last:
                out.append(self._head(x, labels, temperature))
            else:
                self._exchange(send_next=x)
        if not broadcast:
            return out
        src = dist.get_process_group_ranks(self.pp_group)[-1]
        res = []
        for k, (inp, _) in enumerate(micro_batches):
            t = int(inp.input_ids.shape[0])
            buf = torch.stack([o.float() for o in out[k]]) if self.last else torch.empty(2, t, device=self.device, dtype=torch.float32)
            dist.broadcast(buf, src=src, group=self.pp_group)
            res.append((buf[0], buf[1]))
        return res

    # ------------------------------------------------------------------ optimizer
    def zero_grad(self) -> None:
        for p in self.params.values():
            p.grad = None

    def step(self, lr: float | None = None) -> dict:
        lr = self.lr if lr is None else lr
        grads = {}
        for k, p in self.params.items():
            g = p.grad.float() if p.grad is not None else torch.zeros_like(self.master[k])
            if self.dp > 1:
                dist.all_reduce(g, group=self.dp_group)
                g /= self.dp
            grads[k] = g
        if self.cfg.tie_word_embeddings and self.pp > 1:
            # input embedding (stage 0) and output projection (last stage) are one parameter: sum both gradient contributions
            g = grads["embed"] if "embed" in grads else torch.zeros(self.cfg.vocab_size, self.cfg.hidden_size, device=self.device)
            dist.all_reduce(g, group=self.pp_group)
            if "embed" in grads:
                grads["embed"] = g
        own = [g for k, g in grads.items() if not (k == "embed" and self.cfg.tie_word_embeddings and self.pp > 1 and not self.first)]
        sq = sum((g.pow(2).sum() for g in own), torch.zeros((), device=self.device))
        dist.all_reduce(sq, group=self.pp_group)
        norm = float(sq.sqrt().item())
        skipped = not math.isfinite(norm)
        if not skipped:
            self.opt_step += 1
            coef = min(1.0, self.max_grad_norm / (norm + 1e-6)) if self.max_grad_norm > 0 else 1.0
            b1, b2 = self.betas
            with torch.no_grad():
                for k, g in grads.items():
                    ref.adamw_step(self.master[k], self.exp_avg[k], self.exp_avg_sq[k], g, lr, b1, b2, self.eps, self.weight_decay,
                                   self.opt_step, coef)
                    self.params[k].copy_(self.master[k].to(self.dtype))
        self.zero_grad()
        return {"grad_norm": norm, "skipped": skipped}

    # ------------------------------------------------------------------ export
    def gather_full_state(self) -> dict[str, torch.Tensor]:
        """All stages' fp32 masters merged on every rank (naming of ``Zero3Engine.named_full_tensors``)."""
        mine = {k: v.detach().cpu() for k, v in self.master.items()}
        parts: list = [None] * self.pp
        dist.all_gather_object(parts, mine, group=self.pp_group)
        out: dict[str, torch.Tensor] = {}
        for p in parts:
            for k, v in p.items():
                out.setdefault(k, v)
        return out


# =====================================================================================================================
# Interleaved (virtual-stage) 1F1B
# =====================================================================================================================
def build_interleaved_schedule(pp: int, v: int, m: int, forward_only: bool = False) -> list[list[tuple | None]]:
    """Tick table of the interleaved schedule: ``table[t][rank]`` is ``('F' | 'B', vstage, micro_batch)`` or None.

    The model is cut into ``S = pp * v`` virtual stages, vstage ``s`` lives on rank ``s % pp`` (chunk ``s // pp`` of that rank), so a
    micro-batch visits every rank ``v`` times and the pipeline bubble shrinks by ``v`` (Megatron-LM interleaved 1F1B).  The table is
    produced by a deterministic list-scheduling simulation that every rank runs identically: one action per rank per tick; a forward
    needs the previous vstage's forward of that micro-batch from an EARLIER tick, a backward needs the next vstage's backward and
    its own forward; a ready backward wins except right after a backward, when a ready forward goes first (1F1B alternation; this reaches
    Megatron's tick count 2 v m + 2 (pp - 1)), and a rank stops issuing forwards while it already holds the Megatron
    warm-up depth ``2 (pp - rank - 1) + (v - 1) pp + 1`` of live activations.  Because the table is global knowledge, every
    neighbour exchange of a tick is known to both ends and is issued as one grouped send/recv — no dead-lock on stream-ordered
    transports."""
    S = pp * v
    INF = 1 << 60
    done_f = [[INF] * m for _ in range(S)]
    done_b = [[INF] * m for _ in range(S)]
    next_f = [0] * S                      # micro-batches run in order on every vstage
    next_b = [0] * S
    live = [0] * pp
    cap = [INF if forward_only else 2 * (pp - r - 1) + (v - 1) * pp + 1 for r in range(pp)]
    total = S * m * (1 if forward_only else 2)
    table: list[list[tuple | None]] = []
    last: list[str | None] = [None] * pp
    n_done, t, relaxed = 0, 0, False
    while n_done < total:
        row: list[tuple | None] = [None] * pp
        for r in range(pp):
            mine = range(r, S, pp)
            pick = None
            f_ready = any(next_f[s] < m and (s == 0 or done_f[s - 1][next_f[s]] < t) for s in mine)
            # steady state alternates: right after a backward the rank takes a forward if one is ready (one-forward-one-backward)
            if not forward_only and not (last[r] == "B" and f_ready and live[r] < cap[r]):
                cand = []
                for s in mine:
                    k = next_b[s]
                    if k < m and done_f[s][k] < t and (s == S - 1 or done_b[s + 1][k] < t):
                        cand.append((k // pp, -(s // pp), k % pp, s, k))
                if cand:
                    _, _, _, s, k = min(cand)
                    pick = ("B", s, k)
            if pick is None and (live[r] < cap[r] or relaxed):
                cand = []
                for s in mine:
                    k = next_f[s]
                    if k < m and (s == 0 or done_f[s - 1][k] < t):
                        cand.append((k // pp, s // pp, k % pp, s, k))
                if cand:
                    _, _, _, s, k = min(cand)
                    pick = ("F", s, k)
            row[r] = pick
            if pick is not None:
                last[r] = pick[0]
        progressed = False
        for r, a in enumerate(row):
            if a is None:
                continue
            progressed = True
            kind, s, k = a
            if kind == "F":
                done_f[s][k] = t
                next_f[s] += 1
                live[r] += 0 if forward_only else 1
            else:
                done_b[s][k] = t
                next_b[s] += 1
                live[r] -= 1
            n_done += 1
        if progressed:
            table.append(row)
            relaxed = False
            t += 1
        else:
            if relaxed:
                raise RuntimeError("interleaved pipeline schedule cannot make progress")
            relaxed = True                # the activation cap can only stall when m is not a multiple of pp: lift it for one tick
    return table


class InterleavedPipelineEngine(PipelineEngine):
    """``PipelineEngine`` with ``virtual_stages`` model chunks per rank (Megatron's virtual pipeline parallelism,
    ``virtual_pipeline_model_parallel_size``; reference: 3rdparty/verl/verl/workers/megatron_workers.py:204-211 passes it through to
    mcore).  Same parameters, optimizer step, tied-embedding handling and export as the base engine; only the layer placement and
    the schedule differ."""

    def __init__(self, cfg: ModelConfig, device, pp_group, dp_group=None, virtual_stages: int = 2, **kw):
        self.v = int(virtual_stages)
        pp = dist.get_world_size(pp_group)
        if self.v < 1 or cfg.num_layers < pp * self.v:
            raise ValueError(f"{cfg.num_layers} layers cannot fill {pp} x {self.v} virtual pipeline stages")
        self._ranges = stage_layers(cfg.num_layers, pp * self.v)
        super().__init__(cfg, device, pp_group, dp_group, **kw)

    def load_full(self, embed: dict, layers: list[dict], head: dict) -> None:
        self.chunks = [self._ranges[c * self.pp + self.stage] for c in range(self.v)]
        owned = [i for lo, hi in self.chunks for i in range(lo, hi)]
        self.lo, self.hi = owned[0], owned[0]                        # the base loader takes no layers; they are added below
        super().load_full(embed, layers, head)
        extra = {f"layers.{i}.{k}": t.clone() for i in owned for k, t in layers[i].items()}
        for k, t in extra.items():
            self.master[k] = t.to(self.device, torch.float32)
            self.params[k] = self.master[k].to(self.dtype).clone().requires_grad_(True)
            self.exp_avg[k] = torch.zeros_like(self.master[k])
            self.exp_avg_sq[k] = torch.zeros_like(self.master[k])

    # ---- one virtual stage
    def _chunk_forward(self, s: int, inp: tf.PackedInput, x_in: torch.Tensor | None) -> torch.Tensor:
        lo, hi = self._ranges[s]
        x = torch.nn.functional.embedding(inp.input_ids, self.params["embed"]) if s == 0 else x_in
        for i in range(lo, hi):
            x = tf.layer_forward(self.cfg, self.layer(i), x, inp)
        return x

    def _run_table(self, table, micro_batches, on_forward, on_backward) -> None:
        """Execute this rank's column of the tick table; after every tick exchange what the table says crosses a rank boundary."""
        pp, S = self.pp, self.pp * self.v
        ranks = dist.get_process_group_ranks(self.pp_group)
        inbox_act: dict = {}
        inbox_grad: dict = {}
        for row in table:
            mine = row[self.stage]
            out_act = out_grad = None
            if mine is not None:
                kind, s, k = mine
                if kind == "F":
                    out_act = on_forward(s, k, inbox_act.pop((s, k), None))
                else:
                    out_grad = on_backward(s, k, inbox_grad.pop((s, k), None))
            # messages of this tick, in one global order both ends agree on
            msgs = []
            for r, a in enumerate(row):
                if a is None:
                    continue
                kind, s, k = a
                if kind == "F" and s < S - 1:
                    msgs.append((r, (s + 1) % pp, 0, s + 1, k))
                elif kind == "B" and s > 0:
                    msgs.append((r, (s - 1) % pp, 1, s - 1, k))
            ops_ = []
            for src, dst, kind, s, k in sorted(msgs):
                if src == self.stage:
                    t = (out_act if kind == 0 else out_grad).contiguous()
                    ops_.append(dist.P2POp(dist.isend, t, ranks[dst]))
                if dst == self.stage:
                    buf = torch.empty(self._shape(micro_batches[k][0]), device=self.device, dtype=self.dtype)
                    (inbox_act if kind == 0 else inbox_grad)[(s, k)] = buf
                    ops_.append(dist.P2POp(dist.irecv, buf, ranks[src]))
            if ops_:
                for w in dist.batch_isend_irecv(ops_):
                    w.wait()

    # ---- training step
    def train_batch(self, micro_batches, loss_fn, temperature: float = 1.0) -> list[float]:
        S, m = self.pp * self.v, len(micro_batches)
        table = build_interleaved_schedule(self.pp, self.v, m)
        held: dict = {}
        losses: dict[int, float] = {}
        self.peak_in_flight = 0

        def fwd(s, k, x_in):
            inp, labels = micro_batches[k]
            if x_in is not None:
                x_in.requires_grad_(True)
            out = self._chunk_forward(s, inp, x_in)
            if s == S - 1:
                lp, ent = self._head(out, labels, temperature)
                out = loss_fn(lp, ent, k)
                losses[k] = float(out.detach())
            held[(s, k)] = (x_in, out)
            self.peak_in_flight = max(self.peak_in_flight, len(held))
            return None if s == S - 1 else out.detach()

        def bwd(s, k, grad_out):
            x_in, out = held.pop((s, k))
            if s == S - 1:
                out.backward()
            else:
                torch.autograd.backward(out, grad_out)
            return None if x_in is None else x_in.grad

        self._run_table(table, micro_batches, fwd, bwd)
        self.last_schedule_ticks = len(table)
        return [losses[k] for k in sorted(losses)]

    @torch.no_grad()
    def forward_only(self, micro_batches, temperature: float = 1.0, broadcast: bool = True):
        S = self.pp * self.v
        table = build_interleaved_schedule(self.pp, self.v, len(micro_batches), forward_only=True)
        out: dict[int, tuple] = {}

        def fwd(s, k, x_in):
            inp, labels = micro_batches[k]
            x = self._chunk_forward(s, inp, x_in)
            if s == S - 1:
                out[k] = self._head(x, labels, temperature)
                return None
            return x

        self._run_table(table, micro_batches, fwd, None)
        res_local = [out[k] for k in sorted(out)]
        if not broadcast:
            return res_local
        src = dist.get_process_group_ranks(self.pp_group)[-1]
        res = []
        for k, (inp, _) in enumerate(micro_batches):
            t = int(inp.input_ids.shape[0])
            buf = torch.stack([o.float() for o in out[k]]) if self.last else torch.empty(2, t, device=self.device, dtype=torch.float32)
            dist.broadcast(buf, src=src, group=self.pp_group)
            res.append((buf[0], buf[1]))
        return res
