"""HBM planner: per-GPU memory of a colocated rollout + training job on 180 GB B200s and the layout decisions that follow
(resident bf16 replica or per-layer gather, how many blocks keep their activations, whether optimizer state must be offloaded,
KV-cache budget).  The same arithmetic drives ``Zero3Engine``'s ``resident`` / ``act_ckpt=auto`` choices at run time; this module
makes it inspectable up front (BASELINE config 5: Llama-3-70B PPO, TP-8 rollout / ZeRO-3 trainer, "8xB200 HBM sizing").

    python -m polyrl_b200.utils.hbm_sizing llama3-70b --gpus 8 --rollout-tp 8 --critic
"""
from __future__ import annotations

import argparse
import json
from dataclasses import asdict, dataclass, field

from ..models.config import ModelConfig, get_model_config

GB = float(1 << 30)


def count_params(c: ModelConfig) -> int:
    if c.num_experts or c.attn_type != "mha" or c.vision is not None:
        return c.num_params()
    h, f, v, L = c.hidden_size, c.intermediate_size, c.vocab_size, c.num_layers
    qkv = (c.num_heads + 2 * c.num_kv_heads) * c.head_dim
    per_layer = h * qkv + c.num_heads * c.head_dim * h + 3 * h * f + 2 * h
    embed = v * h
    head = 0 if c.tie_word_embeddings else v * h
    return L * per_layer + embed + head + h


@dataclass
class HbmPlan:
    model: str
    n_gpus: int
    params_b: float
    hbm_gb: float
    # per-GPU GB
    zero_shard_bf16: float = 0.0
    optimizer_fp32: float = 0.0          # master + exp_avg + exp_avg_sq + grad accumulator (16 B / param / N)
    resident_replica: float = 0.0        # full bf16 replica kept on every GPU (0 when gathering per layer)
    gather_buffers: float = 0.0          # 2 x largest unit when gathering per layer (double-buffered prefetch)
    rollout_weights: float = 0.0         # extra rollout copy (0 when the engine aliases the resident replica / shard)
    kv_cache: float = 0.0
    critic_and_ref: float = 0.0
    activations_per_saved_block: float = 0.0
    saved_blocks: int = 0
    activations: float = 0.0
    reserve: float = 16.0
    total: float = 0.0
    decisions: dict = field(default_factory=dict)


def plan(model: str | ModelConfig, n_gpus: int = 8, rollout_tp: int = 1, tokens_per_microbatch: int = 16384, max_num_seqs: int = 256,
         max_model_len: int = 2048, hbm_gb: float = 180.0, critic: bool | str = False, ref: bool = False,
         free_kv_during_training: bool = True, optimizer_offload: bool | None = None, reserve_gb: float = 16.0,
         offload_idle_optimizer: bool = True) -> HbmPlan:
    """``critic``: False, True (same architecture as the actor) or a model name.  ``offload_idle_optimizer``: with actor and
    critic updated one after the other, the fp32 state of the model that is NOT being updated sits in pinned host memory
    (``Zero3Engine.offload_optimizer``), so only the larger of the two counts against HBM."""
    c = get_model_config(model) if isinstance(model, str) else model
    P = count_params(c)
    p = HbmPlan(model=c.name if hasattr(c, "name") else str(model), n_gpus=n_gpus, params_b=P / 1e9, hbm_gb=hbm_gb, reserve=reserve_gb)
    p.zero_shard_bf16 = 2 * P / n_gpus / GB
    opt = 16 * P / n_gpus / GB
    replica = 2 * P / GB
    resident = n_gpus > 1 and replica <= hbm_gb / 4
    p.resident_replica = replica if resident else 0.0
    qkv = (c.num_heads + 2 * c.num_kv_heads) * c.head_dim
    unit = 2 * (c.hidden_size * qkv + c.num_heads * c.head_dim * c.hidden_size + 3 * c.hidden_size * c.intermediate_size) / GB
    head_unit = 2 * c.vocab_size * c.hidden_size / GB
    p.gather_buffers = 0.0 if (resident or n_gpus == 1) else 2 * max(unit, head_unit)
    # rollout weights: TP=1 aliases the trainer's replica (or the shard when N == 1); TP>1 needs its own TP-sliced copy
    if rollout_tp > 1:
        p.rollout_weights = 2 * P / rollout_tp / GB
    elif not resident and n_gpus > 1:
        p.rollout_weights = replica
    kv_tok = 2 * c.num_layers * (c.num_kv_heads // max(1, rollout_tp) or 1) * c.head_dim * 2
    if c.attn_type == "mla":                    # one shared latent entry per token and layer (models/mla.py)
        kv_tok = c.num_layers * c.kv_cache_layout()[1] * 2
    p.kv_cache = max_num_seqs * max_model_len * kv_tok / GB
    Pc = 0 if not critic else (P if critic is True else count_params(get_model_config(critic)))
    critic_opt = 16 * Pc / n_gpus / GB
    extra_models = 2 * Pc / n_gpus / GB + (1 if ref else 0) * 2 * P / n_gpus / GB
    if Pc and not offload_idle_optimizer:
        extra_models += critic_opt
    p.critic_and_ref = extra_models
    fixed = p.zero_shard_bf16 + p.resident_replica + p.gather_buffers + p.rollout_weights + p.critic_and_ref + reserve_gb
    if optimizer_offload is None:       # offload the fp32 state to pinned host memory during rollout when it cannot co-exist with the KV pool
        optimizer_offload = fixed + opt + p.kv_cache > hbm_gb
    if Pc and offload_idle_optimizer:
        opt = max(opt, critic_opt)          # one optimizer state resident at a time
    p.optimizer_fp32 = opt
    train_fixed = fixed + opt + (0.0 if free_kv_during_training else p.kv_cache)
    inter = c.intermediate_size
    per_block = tokens_per_microbatch * (2 * (5 * c.hidden_size + qkv + 3 * inter) + 64) / GB
    p.activations_per_saved_block = per_block
    boundary = tokens_per_microbatch * 2 * c.hidden_size * c.num_layers / GB        # block inputs kept for recompute
    room = hbm_gb - train_fixed - boundary - 2 * per_block          # 2 blocks of working set while a recomputed block runs
    stable = resident or n_gpus == 1
    p.saved_blocks = int(max(0, min(c.num_layers, room // per_block))) if stable else 0
    p.activations = boundary + p.saved_blocks * per_block + 2 * per_block * (p.saved_blocks < c.num_layers)
    p.total = train_fixed + p.activations
    release_rollout_w = False
    if p.total > hbm_gb and p.rollout_weights > 0:          # lever 1: the engine's TP copy is rebuilt at the next weight sync anyway
        release_rollout_w = True
        p.total -= p.rollout_weights
    tokens_fit = tokens_per_microbatch
    while p.total > hbm_gb and tokens_fit > 2048:           # lever 2: smaller packed micro-batches (more of them per step)
        tokens_fit //= 2
        p.total -= p.activations / 2
        p.activations /= 2
    rollout_total = fixed + (0.0 if optimizer_offload else opt) + p.kv_cache
    p.decisions = {
        "bf16_replica": "resident (1 gather / optimizer step, aliased by the rollout engine)" if resident else
                        ("shard == full model" if n_gpus == 1 else "gathered per layer (ZeRO-3), double-buffered prefetch"),
        "activation_checkpoint": f"{p.saved_blocks}/{c.num_layers} blocks keep activations, rest recomputed",
        "optimizer_offload_during_rollout": bool(optimizer_offload),
        "release_rollout_weights_during_training": release_rollout_w,
        "max_tokens_per_microbatch": tokens_fit,
        "fits_training": bool(p.total <= hbm_gb),
        "fits_rollout": bool(rollout_total <= hbm_gb),
        "rollout_phase_gb": round(rollout_total, 1),
        "training_phase_gb": round(p.total, 1),
    }
    return p


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--rollout-tp", type=int, default=1)
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--max-num-seqs", type=int, default=256)
    ap.add_argument("--max-model-len", type=int, default=2048)
    ap.add_argument("--critic", nargs="?", const=True, default=False, help="flag = same model as the actor, or a model name")
    ap.add_argument("--ref", action="store_true")
    a = ap.parse_args(argv)
    p = plan(a.model, a.gpus, a.rollout_tp, a.tokens, a.max_num_seqs, a.max_model_len, critic=a.critic, ref=a.ref)
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in asdict(p).items()}, indent=1))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
