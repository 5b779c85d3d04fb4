"""LoRA adapters for the fused-projection transformer blocks.

The reference trains LoRA through peft on top of FSDP (``lora_rank`` / ``lora_alpha`` / ``target_modules``,
3rdparty/verl/verl/workers/fsdp_workers.py:269-577) and merges the adapters layer by layer before handing weights
to the rollout engine (verl/utils/fsdp_utils.py:569-694 "layered summon").  Here an adapter is an extra, trainable
ZeRO unit per decoder block next to a FROZEN base unit (no fp32 master / moments / gradient storage for the base:
12 bytes per base parameter saved), the low-rank path is added to the fused qkv / o / gate-up / down projections,
and ``merge_layer_`` folds ``scale * B @ A`` into a weight replica for the rollout engine.

Adapters are kept per HF projection (q/k/v/o/gate/up/down) so they export 1:1 to peft names.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from ..ops import reference as ref
from .config import ModelConfig

ALL_LINEAR = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")
_PEFT_MODULE_NAMES = {"q_proj": "self_attn.q_proj", "k_proj": "self_attn.k_proj", "v_proj": "self_attn.v_proj", "o_proj": "self_attn.o_proj",
                      "gate_proj": "mlp.gate_proj", "up_proj": "mlp.up_proj", "down_proj": "mlp.down_proj"}


@dataclass
class LoraConfig:
    rank: int = 8
    alpha: float = 16.0
    target_modules: tuple = ALL_LINEAR

    @property
    def scale(self) -> float:
        return self.alpha / self.rank

    @classmethod
    def from_model_cfg(cls, rank: int, alpha: float, target_modules, exclude_modules: str | None = None) -> "LoraConfig | None":
        """``exclude_modules``: regex over HF-style module names (``model.layers.0.self_attn.q_proj``, ``visual...``); a target whose
        name matches gets no adapter (verl model.exclude_modules, e.g. ".*visual.*" — the vision tower never carries adapters here)."""
        if not rank or rank <= 0:
            return None
        if target_modules in (None, "all-linear", "all_linear"):
            tm = ALL_LINEAR
        elif isinstance(target_modules, str):
            tm = tuple(t.strip() for t in target_modules.strip("[]").split(",") if t.strip())
        else:
            tm = tuple(target_modules)
        bad = [t for t in tm if t not in ALL_LINEAR]
        if bad:
            raise ValueError(f"unsupported LoRA target module(s) {bad}; supported: {ALL_LINEAR}")
        if exclude_modules:
            import re

            pat = re.compile(exclude_modules)
            tm = tuple(t for t in tm if not pat.fullmatch(f"model.layers.0.{_PEFT_MODULE_NAMES[t]}") and not pat.fullmatch(t))
            if not tm:
                raise ValueError(f"exclude_modules={exclude_modules!r} removes every LoRA target")
        return cls(rank=int(rank), alpha=float(alpha), target_modules=tm)


def _proj_dims(cfg: ModelConfig) -> dict[str, tuple[int, int]]:
    """projection -> (out_features, in_features)"""
    h, f, d = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    return {"q_proj": (cfg.num_heads * d, h), "k_proj": (cfg.num_kv_heads * d, h), "v_proj": (cfg.num_kv_heads * d, h),
            "o_proj": (h, cfg.num_heads * d), "gate_proj": (f, h), "up_proj": (f, h), "down_proj": (h, f)}


def lora_param_specs(cfg: ModelConfig, lcfg: LoraConfig) -> list[tuple[str, tuple[int, ...]]]:
    if cfg.family != "llama":
        raise NotImplementedError("LoRA is implemented for the llama/qwen block family")
    dims = _proj_dims(cfg)
    specs = []
    for p in ALL_LINEAR:
        if p in lcfg.target_modules:
            out_f, in_f = dims[p]
            specs.append((f"lora_{p}_A", (lcfg.rank, in_f)))
            specs.append((f"lora_{p}_B", (out_f, lcfg.rank)))
    return specs


def _delta(x: torch.Tensor, w: dict, proj: str) -> torch.Tensor | None:
    a = w.get(f"lora_{proj}_A")
    if a is None:
        return None
    b = w[f"lora_{proj}_B"]
    return torch.nn.functional.linear(torch.nn.functional.linear(x, a.to(x.dtype)), b.to(x.dtype))


def has_lora(w: dict, *projs: str) -> bool:
    return any(f"lora_{p}_A" in w for p in projs)


def qkv_delta(cfg: ModelConfig, w: dict, h: torch.Tensor) -> torch.Tensor:
    """Low-rank update of the fused [q | k | v] projection output."""
    dims = _proj_dims(cfg)
    parts = []
    for p in ("q_proj", "k_proj", "v_proj"):
        d = _delta(h, w, p)
        parts.append(d if d is not None else h.new_zeros(h.shape[0], dims[p][0]))
    return torch.cat(parts, dim=-1) * w["_lora_scale"]


def single_delta(w: dict, x: torch.Tensor, proj: str) -> torch.Tensor:
    return _delta(x, w, proj) * w["_lora_scale"]


def gate_up_delta(cfg: ModelConfig, w: dict, h: torch.Tensor) -> torch.Tensor:
    """Update of the block-interleaved [128 gate | 128 up] fused projection output."""
    f = cfg.intermediate_size
    t = h.shape[0]
    dg, du = _delta(h, w, "gate_proj"), _delta(h, w, "up_proj")
    dg = dg if dg is not None else h.new_zeros(t, f)
    du = du if du is not None else h.new_zeros(t, f)
    return torch.stack([dg.view(t, f // 128, 128), du.view(t, f // 128, 128)], dim=2).reshape(t, 2 * f) * w["_lora_scale"]


@torch.no_grad()
def merge_layer_(cfg: ModelConfig, base: dict, lora: dict, scale: float) -> None:
    """In place: fold the adapters of one block into (a copy of) its fused base weights."""
    nq, nkv, d = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim

    def ba(p):
        a = lora.get(f"lora_{p}_A")
        return None if a is None else (lora[f"lora_{p}_B"].float() @ a.float()) * scale

    row = 0
    for p, rows in (("q_proj", nq * d), ("k_proj", nkv * d), ("v_proj", nkv * d)):
        m = ba(p)
        if m is not None:
            base["wqkv"][row:row + rows] += m.to(base["wqkv"].dtype)
        row += rows
    m = ba("o_proj")
    if m is not None:
        base["wo"] += m.to(base["wo"].dtype)
    g, u = ba("gate_proj"), ba("up_proj")
    if g is not None or u is not None:
        f, h = cfg.intermediate_size, cfg.hidden_size
        z = torch.zeros(f, h, device=base["wgu"].device)
        base["wgu"] += ref.interleave_gate_up(g if g is not None else z, u if u is not None else z).to(base["wgu"].dtype)
    m = ba("down_proj")
    if m is not None:
        base["wd"] += m.to(base["wd"].dtype)


_PEFT_MODULE = {"q_proj": "self_attn.q_proj", "k_proj": "self_attn.k_proj", "v_proj": "self_attn.v_proj",
                "o_proj": "self_attn.o_proj", "gate_proj": "mlp.gate_proj", "up_proj": "mlp.up_proj", "down_proj": "mlp.down_proj"}


def export_peft_state_dict(named: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """``layers.{i}.lora.lora_{proj}_{A|B}`` -> peft ``base_model.model.model.layers.{i}.<module>.lora_{A|B}.weight``"""
    out = {}
    for k, v in named.items():
        if ".lora.lora_" not in k:
            continue
        layer = k.split(".")[1]
        tail = k.split(".lora.lora_")[1]
        proj, ab = tail.rsplit("_", 1)
        out[f"base_model.model.model.layers.{layer}.{_PEFT_MODULE[proj]}.lora_{ab}.weight"] = v
    return out


def peft_adapter_config(cfg: ModelConfig, lcfg: LoraConfig) -> dict:
    return {"peft_type": "LORA", "r": lcfg.rank, "lora_alpha": lcfg.alpha, "target_modules": list(lcfg.target_modules),
            "bias": "none", "task_type": "CAUSAL_LM", "lora_dropout": 0.0}


def import_peft_state_dict(peft_sd: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """Inverse of ``export_peft_state_dict``: peft adapter tensors (any prefix up to ``layers.{i}.``, optional ``.default``
    adapter name) -> ``layers.{i}.lora.lora_{proj}_{A|B}``.  Tensors of modules this framework does not adapt are ignored."""
    import re

    inv = {v: k for k, v in _PEFT_MODULE.items()}
    out = {}
    for k, v in peft_sd.items():
        m = re.search(r"layers\.(\d+)\.((?:self_attn|mlp)\.\w+)\.lora_([AB])(?:\.\w+)?\.weight$", k)
        if m and m.group(2) in inv:
            out[f"layers.{m.group(1)}.lora.lora_{inv[m.group(2)]}_{m.group(3)}"] = v
    return out


def load_peft_adapter(engine, path: str) -> int:
    """Start the ZeRO engine's adapter units from a peft adapter directory (``adapter_model.safetensors`` / ``.bin`` next to
    ``adapter_config.json``); reference: model.lora_adapter_path, verl/workers/fsdp_workers.py (PeftModel.from_pretrained).  Rank and
    target modules must match the engine's LoRA config.  Returns the number of tensors loaded."""
    import json
    from pathlib import Path

    from .transformer import rope_feature_perm

    if rope_feature_perm(engine.cfg) is not None:
        raise NotImplementedError("peft adapters for checkpoints whose rotary layout is permuted at load time (GLM family): "
                                  "the q / k adapter rows would need the same permutation")
    p = Path(path)
    cfg_file = p / "adapter_config.json"
    if cfg_file.exists():
        ac = json.loads(cfg_file.read_text())
        if engine.lora is not None and int(ac.get("r", engine.lora.rank)) != engine.lora.rank:
            raise ValueError(f"adapter rank {ac.get('r')} != model.lora_rank {engine.lora.rank}")
    if (p / "adapter_model.safetensors").exists():
        from safetensors.torch import load_file

        sd = load_file(str(p / "adapter_model.safetensors"))
    elif (p / "adapter_model.bin").exists():
        sd = torch.load(p / "adapter_model.bin", map_location="cpu", weights_only=True)
    else:
        raise FileNotFoundError(f"no adapter_model.safetensors / adapter_model.bin in {path!r}")
    ours = import_peft_state_dict(sd)
    n = 0
    for i, unit in enumerate(engine.lora_layers):
        pre = f"layers.{i}.lora."
        named = {k[len(pre):]: v for k, v in ours.items() if k.startswith(pre)}
        if named:
            missing = [name for name, _ in unit.specs if name not in named]
            if missing:
                raise KeyError(f"adapter at {path!r} lacks {missing[:3]} for layer {i} (target modules differ?)")
            unit.load_full_state(named)
            n += len(named)
    return n
