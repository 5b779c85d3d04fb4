"""Supervised fine-tuning on the same ZeRO engine and fused kernels (packed sequences, fused linear-CE, fused AdamW) — the
warm-start stage that usually precedes RL.

    python -m polyrl_b200.trainer.sft data.train_files=sft.jsonl model.path=qwen2.5-0.5b trainer.total_steps=100

Reference capability: verl's FSDP SFT trainer and SFT dataset (3rdparty/verl/verl/trainer/fsdp_sft_trainer.py,
verl/utils/dataset/sft_dataset.py; loss parity test tests/special_e2e/sft/test_sp_loss_match.py): prompt + response are
concatenated, the loss is the token-mean negative log-likelihood of the RESPONSE tokens only, optional Ulysses SP.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time
from dataclasses import dataclass, field

import torch
import torch.distributed as dist
from torch.utils.data import Dataset

from ..config.schema import BaseConfig, OptimConfig, apply_overrides
from ..data.rl_dataset import _read_rows, load_tokenizer
from ..models import get_model_config, pack_padded
from ..parallel.zero3 import Zero3Engine
from ..utils.checkpoint import ShardedCheckpointManager


@dataclass
class SFTDataConfig(BaseConfig):
    train_files: str | None = None
    val_files: str | None = None
    prompt_key: str = "prompt"
    response_key: str = "response"
    multiturn_messages_key: str | None = None      # set (e.g. "messages") to train on the assistant turns of chat transcripts
    max_length: int = 1024
    truncation: str = "right"            # right | error
    micro_batch_size_per_gpu: int = 4
    train_batch_size: int = 16           # global sequences per optimizer step
    synthetic: bool = False
    synthetic_num_samples: int = 64
    seed: int = 1
    # spellings of the reference's SFT configs (verl/trainer/config/sft_trainer*.yaml): resolved by finalize_sft_config
    messages_key: str | None = None                 # = multiturn_messages_key
    multiturn: dict = field(default_factory=dict)   # {"enable": bool, "messages_key": ..., "tools_key": ...}
    micro_batch_size: int | None = None             # GLOBAL micro-batch: divided by the world size
    max_token_len_per_gpu: int | None = None        # token budget of a micro-batch when use_dynamic_bsz (sequences are packed)
    use_dynamic_bsz: bool = False
    pad_mode: str = "right"                         # right | no_padding (packed; what the engine does internally either way)
    prompt_dict_keys: list | None = None
    response_dict_keys: list | None = None
    use_shm: bool = False
    apply_chat_template_kwargs: dict = field(default_factory=dict)


@dataclass
class SFTModelConfig(BaseConfig):
    path: str = "tiny-llama"
    tokenizer_path: str | None = None
    override_config: dict = field(default_factory=dict)
    lora_rank: int = 0
    lora_alpha: int = 16
    target_modules: str = "all-linear"
    partial_pretrain: str | None = None             # older spelling of path
    strategy: str = "fsdp"
    use_remove_padding: bool = True                 # always: sequences are packed
    use_liger: bool = False                         # third-party fused kernels of other stacks: the own fused kernels are always on
    trust_remote_code: bool = False
    enable_gradient_checkpointing: bool = True
    fsdp_config: dict = field(default_factory=dict)


@dataclass
class SFTEngineConfig(BaseConfig):
    """Parallel layout block of the reference's newer SFT trainer (``engine.*``)."""
    strategy: str = "fsdp"
    fsdp_size: int = -1
    ulysses_sequence_parallel_size: int = 1
    tensor_model_parallel_size: int = 1
    pipeline_model_parallel_size: int = 1
    virtual_pipeline_model_parallel_size: int | None = None
    context_parallel_size: int = 1
    use_mbridge: bool = False
    param_offload: bool = False
    optimizer_offload: bool = False


@dataclass
class SFTTrainerConfig(BaseConfig):
    device: str = "cuda"
    total_steps: int = 100
    total_epochs: int = 1
    save_freq: int = -1
    eval_freq: int = -1
    default_local_dir: str = "checkpoints/sft"
    ulysses_sequence_parallel_size: int = 1
    comm_backend: str = "p2p"
    logger: list = field(default_factory=lambda: ["console"])
    total_training_steps: int | None = None         # = total_steps
    test_freq: int | None = None                    # = eval_freq
    project_name: str = "polyrl_b200_sft"
    experiment_name: str = "sft"
    resume_mode: str = "auto"
    resume_from_path: str | None = None
    max_ckpt_to_keep: int | None = None
    nnodes: int = 1
    n_gpus_per_node: int = 8
    seed: int = 1


@dataclass
class SFTConfig(BaseConfig):
    data: SFTDataConfig = field(default_factory=SFTDataConfig)
    model: SFTModelConfig = field(default_factory=SFTModelConfig)
    optim: OptimConfig = field(default_factory=lambda: OptimConfig(lr=1e-5))
    trainer: SFTTrainerConfig = field(default_factory=SFTTrainerConfig)
    engine: SFTEngineConfig = field(default_factory=SFTEngineConfig)
    checkpoint: dict = field(default_factory=dict)          # {"save_contents": [...], "load_contents": [...]}
    ulysses_sequence_parallel_size: int | None = None       # top-level spelling of the older fsdp_sft_trainer config
    use_remove_padding: bool = True


def finalize_sft_config(cfg: SFTConfig, world: int = 1) -> SFTConfig:
    """Map the reference's spellings onto the fields the trainer reads."""
    d, m, t, e = cfg.data, cfg.model, cfg.trainer, cfg.engine
    if m.partial_pretrain:
        m.path = m.partial_pretrain
    mt = d.multiturn or {}
    if mt.get("enable") or d.messages_key:
        d.multiturn_messages_key = d.multiturn_messages_key or mt.get("messages_key") or d.messages_key or "messages"
    if d.micro_batch_size and d.micro_batch_size_per_gpu == SFTDataConfig.micro_batch_size_per_gpu:
        d.micro_batch_size_per_gpu = max(1, int(d.micro_batch_size) // max(1, world))
    if d.use_dynamic_bsz and d.max_token_len_per_gpu:
        d.micro_batch_size_per_gpu = max(1, int(d.max_token_len_per_gpu) // max(1, int(d.max_length)))
    if t.total_training_steps:
        t.total_steps = int(t.total_training_steps)
    if t.test_freq is not None:
        t.eval_freq = int(t.test_freq)
    sp = cfg.ulysses_sequence_parallel_size or e.ulysses_sequence_parallel_size
    if sp and sp > 1:
        t.ulysses_sequence_parallel_size = int(sp)
    if e.tensor_model_parallel_size > 1 or e.pipeline_model_parallel_size > 1 or e.context_parallel_size > 1:
        raise NotImplementedError("the SFT trainer runs on the ZeRO engine (+ Ulysses SP): tensor / pipeline / context parallel "
                                  "layouts are not available here")
    return cfg


class SFTDataset(Dataset):
    """rows {prompt, response} (strings, or chat ``messages`` for the prompt) -> input_ids / attention_mask / loss_mask,
    right-padded to ``max_length``; the loss mask covers the response tokens and the closing EOS."""

    def __init__(self, files, tokenizer, cfg: SFTDataConfig, vocab_size: int | None = None):
        self.tok, self.cfg = tokenizer, cfg
        if cfg.synthetic or not files:
            g = torch.Generator().manual_seed(cfg.seed)
            v = vocab_size or tokenizer.vocab_size
            self.rows = None
            self.synth = [(torch.randint(3, v, (int(torch.randint(4, 12, (1,), generator=g)),), generator=g).tolist(),
                           torch.randint(3, v, (int(torch.randint(4, 16, (1,), generator=g)),), generator=g).tolist())
                          for _ in range(cfg.synthetic_num_samples)]
        else:
            self.rows = _read_rows(files)

    def __len__(self) -> int:
        return len(self.rows) if self.rows is not None else len(self.synth)

    def _ids(self, i: int) -> tuple[list[int], list[int]]:
        if self.rows is None:
            return self.synth[i]
        row = self.rows[i]
        p = row[self.cfg.prompt_key]
        if isinstance(p, (list, tuple)):
            p_ids = list(self.tok.apply_chat_template(list(p), add_generation_prompt=True, tokenize=True))
        else:
            p_ids = list(self.tok.encode(str(p), add_special_tokens=False))
        r_ids = list(self.tok.encode(str(row[self.cfg.response_key]), add_special_tokens=False)) + [self.tok.eos_token_id]
        return p_ids, r_ids

    def __getitem__(self, i: int) -> dict:
        p_ids, r_ids = self._ids(i)
        ids = p_ids + r_ids
        L = self.cfg.max_length
        if len(ids) > L:
            if self.cfg.truncation == "error":
                raise ValueError(f"sample {i} has {len(ids)} tokens > max_length {L}")
            ids = ids[:L]
        n, np_ = len(ids), min(len(p_ids), L)
        pad = getattr(self.tok, "pad_token_id", 0) or 0
        input_ids = torch.full((L,), pad, dtype=torch.long)
        input_ids[:n] = torch.tensor(ids, dtype=torch.long)
        attn = torch.zeros(L, dtype=torch.long)
        attn[:n] = 1
        loss_mask = torch.zeros(L, dtype=torch.long)
        loss_mask[np_:n] = 1
        return {"input_ids": input_ids, "attention_mask": attn, "loss_mask": loss_mask}


class MultiTurnSFTDataset(Dataset):
    """rows {messages: [{role, content}, ...]} -> input_ids / attention_mask / loss_mask where only the ASSISTANT turns (their content
    and end-of-turn tokens, not the generation prompt that introduces them) are trained on.  Turn boundaries come from rendering the
    chat template on growing prefixes of the conversation, so any prefix-consistent template works (reference:
    verl/utils/dataset/multiturn_sft_dataset.py)."""

    def __init__(self, files, tokenizer, cfg: SFTDataConfig, messages_key: str = "messages"):
        self.tok, self.cfg, self.key = tokenizer, cfg, messages_key
        self.rows = _read_rows(files)

    def __len__(self) -> int:
        return len(self.rows)

    def tokenize(self, messages: list[dict]) -> tuple[list[int], list[int]]:
        ids: list[int] = []
        mask: list[int] = []
        for i, m in enumerate(messages):
            cur = list(self.tok.apply_chat_template(messages[:i + 1], add_generation_prompt=False, tokenize=True))
            if cur[:len(ids)] != ids:
                raise ValueError("chat template is not prefix-consistent; cannot locate turn boundaries")
            new = cur[len(ids):]
            if m["role"] == "assistant":
                gen = list(self.tok.apply_chat_template(messages[:i], add_generation_prompt=True, tokenize=True))
                n_prompt = max(0, min(len(new), len(gen) - len(ids)))        # tokens that only announce the assistant turn
                mask += [0] * n_prompt + [1] * (len(new) - n_prompt)
            else:
                mask += [0] * len(new)
            ids = cur
        return ids, mask

    def __getitem__(self, i: int) -> dict:
        ids, mask = self.tokenize(list(self.rows[i][self.key]))
        L = self.cfg.max_length
        if len(ids) > L:
            if self.cfg.truncation == "error":
                raise ValueError(f"sample {i} has {len(ids)} tokens > max_length {L}")
            ids, mask = ids[:L], mask[:L]
        n = len(ids)
        pad = getattr(self.tok, "pad_token_id", 0) or 0
        input_ids = torch.full((L,), pad, dtype=torch.long)
        input_ids[:n] = torch.tensor(ids, dtype=torch.long)
        attn = torch.zeros(L, dtype=torch.long)
        attn[:n] = 1
        loss_mask = torch.zeros(L, dtype=torch.long)
        loss_mask[:n] = torch.tensor(mask, dtype=torch.long)
        return {"input_ids": input_ids, "attention_mask": attn, "loss_mask": loss_mask}


class PairwiseRMDataset(Dataset):
    """rows {prompt, chosen, rejected} -> input_ids / attention_mask of shape [2, L] (row 0 = chosen, row 1 = rejected), right-padded —
    the input format of pairwise reward-model training (reference: verl/utils/dataset/rm_dataset.py)."""

    def __init__(self, files, tokenizer, max_length: int = 1024, prompt_key: str = "prompt", chosen_key: str = "chosen",
                 rejected_key: str = "rejected"):
        self.tok, self.L = tokenizer, max_length
        self.keys = (prompt_key, chosen_key, rejected_key)
        self.rows = _read_rows(files)

    def __len__(self) -> int:
        return len(self.rows)

    def __getitem__(self, i: int) -> dict:
        row = self.rows[i]
        p = row[self.keys[0]]
        p_ids = (list(self.tok.apply_chat_template(list(p), add_generation_prompt=True, tokenize=True)) if isinstance(p, (list, tuple))
                 else list(self.tok.encode(str(p), add_special_tokens=False)))
        pad = getattr(self.tok, "pad_token_id", 0) or 0
        ids = torch.full((2, self.L), pad, dtype=torch.long)
        attn = torch.zeros(2, self.L, dtype=torch.long)
        for j, k in enumerate(self.keys[1:]):
            seq = (p_ids + list(self.tok.encode(str(row[k]), add_special_tokens=False)) + [self.tok.eos_token_id])[: self.L]
            ids[j, :len(seq)] = torch.tensor(seq, dtype=torch.long)
            attn[j, :len(seq)] = 1
        return {"input_ids": ids, "attention_mask": attn}


def lr_at(o: OptimConfig, step: int, total: int) -> float:
    warm = o.lr_warmup_steps if o.lr_warmup_steps >= 0 else int(o.lr_warmup_steps_ratio * total)
    if warm > 0 and step < warm:
        return o.lr * (step + 1) / warm
    if o.warmup_style == "cosine" and total > warm:
        prog = (step - warm) / max(1, total - warm)
        return o.lr * (o.min_lr_ratio + (1 - o.min_lr_ratio) * 0.5 * (1 + math.cos(math.pi * 2 * o.num_cycles * prog)))
    return o.lr


class SFTTrainer:
    def __init__(self, cfg: SFTConfig, device=None, group=None):
        self.cfg, self.group = cfg, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = device or (f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}" if (cfg.trainer.device == "cuda" and torch.cuda.is_available()) else "cpu")
        self.device = torch.device(dev)
        self.model_cfg = get_model_config(cfg.model.path, **cfg.model.override_config)
        from ..models.lora import LoraConfig

        lora = LoraConfig.from_model_cfg(cfg.model.lora_rank, cfg.model.lora_alpha, cfg.model.target_modules)
        o = cfg.optim
        self.engine = Zero3Engine(self.model_cfg, self.device, group=group, comm=cfg.trainer.comm_backend, lr=o.lr, betas=tuple(o.betas),
                                  eps=o.eps, weight_decay=o.weight_decay, max_grad_norm=o.clip_grad, seed=cfg.data.seed, lora=lora)
        self.engine.act_ckpt = "auto"
        self.ckpt = ShardedCheckpointManager(self.engine, self.model_cfg)
        self.step_idx = 0

    def _loss_on(self, batch: dict, denom: float) -> torch.Tensor:
        """Sum over this micro-batch's response tokens of -log p, divided by the GLOBAL response-token count of the step."""
        ids, mask, lm = (batch[k].to(self.device) for k in ("input_ids", "attention_mask", "loss_mask"))
        inp, idx = pack_padded(ids, mask)
        labels = torch.roll(inp.input_ids, -1)                  # next token inside the packed stream
        w = torch.roll(lm.reshape(-1)[idx], -1).float()          # token t is scored by the prediction made at t-1
        ends = (inp.cu_seqlens[1:] - 1).long()
        w[ends] = 0.0                                            # the last position of a sequence predicts another sequence's token
        lp, _ = self.engine.forward_logprob_entropy(inp, labels, 1.0)
        return -(lp * w).sum() / denom

    def train_step(self, batch: dict) -> dict:
        """``batch``: this rank's sequences for one optimizer step (already sharded across data-parallel ranks)."""
        n = batch["input_ids"].shape[0]
        ntok = batch["loss_mask"][:, 1:].sum().float().to(self.device)
        if self.world > 1:
            dist.all_reduce(ntok, group=self.group)
        denom = float(ntok.clamp_min(1).item()) / self.world       # ZeRO averages gradients over ranks
        mb = self.cfg.data.micro_batch_size_per_gpu
        total = 0.0
        for s in range(0, n, mb):
            loss = self._loss_on({k: v[s:s + mb] for k, v in batch.items()}, denom)
            loss.backward()
            total += float(loss.detach())
        lr = lr_at(self.cfg.optim, self.step_idx, self.cfg.trainer.total_steps)
        st = self.engine.step(lr)
        self.step_idx += 1
        loss_t = torch.tensor([total], device=self.device)
        if self.world > 1:
            dist.all_reduce(loss_t, group=self.group)
            loss_t /= self.world
        return {"train/loss": float(loss_t.item()), "train/lr": lr, "train/grad_norm": st["grad_norm"]}

    @torch.no_grad()
    def evaluate(self, ds: Dataset, max_batches: int = 8) -> float:
        bs, tot, cnt = self.cfg.data.micro_batch_size_per_gpu, 0.0, 0.0
        for b in range(min(max_batches, (len(ds) + bs - 1) // bs)):
            rows = [ds[i] for i in range(b * bs, min(len(ds), (b + 1) * bs))]
            batch = {k: torch.stack([r[k] for r in rows]) for k in rows[0]}
            n = float(batch["loss_mask"][:, 1:].sum())
            tot += float(self._loss_on(batch, 1.0))
            cnt += n
        return tot / max(cnt, 1.0)

    def fit(self, train: Dataset, val: Dataset | None = None) -> dict:
        c = self.cfg
        per_rank = max(1, c.data.train_batch_size // self.world)
        g = torch.Generator().manual_seed(c.data.seed)
        last: dict = {}
        t0 = time.time()
        for epoch in range(c.trainer.total_epochs):
            order = torch.randperm(len(train), generator=g).tolist()
            for s in range(0, len(order) - c.data.train_batch_size + 1, c.data.train_batch_size):
                if self.step_idx >= c.trainer.total_steps:
                    break
                mine = order[s + self.rank * per_rank: s + (self.rank + 1) * per_rank]
                rows = [train[i] for i in mine]
                last = self.train_step({k: torch.stack([r[k] for r in rows]) for k in rows[0]})
                if val is not None and c.trainer.eval_freq > 0 and self.step_idx % c.trainer.eval_freq == 0:
                    last["val/loss"] = self.evaluate(val)
                if self.rank == 0:
                    print(f"step:{self.step_idx} - " + " - ".join(f"{k}:{v:.5g}" for k, v in last.items()), flush=True)
                if c.trainer.save_freq > 0 and self.step_idx % c.trainer.save_freq == 0:
                    self.ckpt.save(os.path.join(c.trainer.default_local_dir, f"global_step_{self.step_idx}"), self.step_idx)
        last["time_s"] = time.time() - t0
        return last


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("overrides", nargs="*")
    a = ap.parse_args(argv)
    cfg = SFTConfig()
    apply_overrides(cfg, a.overrides)
    world = int(os.environ.get("WORLD_SIZE", 1))
    finalize_sft_config(cfg, world)
    cuda = cfg.trainer.device == "cuda" and torch.cuda.is_available()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if cuda:
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
            dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
        else:
            dist.init_process_group("gloo")
    tr = SFTTrainer(cfg)
    tok = load_tokenizer(cfg.model.tokenizer_path or "byte") if not cfg.data.synthetic else None
    if cfg.data.multiturn_messages_key and not cfg.data.synthetic:
        train = MultiTurnSFTDataset(cfg.data.train_files, tok, cfg.data, cfg.data.multiturn_messages_key)
        val = MultiTurnSFTDataset(cfg.data.val_files, tok, cfg.data, cfg.data.multiturn_messages_key) if cfg.data.val_files else None
    else:
        train = SFTDataset(cfg.data.train_files, tok, cfg.data, vocab_size=tr.model_cfg.vocab_size)
        val = SFTDataset(cfg.data.val_files, tok, cfg.data, vocab_size=tr.model_cfg.vocab_size) if cfg.data.val_files else None
    out = tr.fit(train, val)
    if tr.rank == 0:
        print("[sft] done:", {k: round(v, 5) for k, v in out.items()}, flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
