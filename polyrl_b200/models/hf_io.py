"""HuggingFace checkpoint directories <-> the ZeRO engine / engine weights.

* ``HFCheckpoint(dir)``: lazy tensor access over ``*.safetensors`` (+ ``model.safetensors.index.json``) or ``pytorch_model*.bin``.
* ``load_into_engine(engine, dir)``: one unit (embedding, one block, head) at a time — read only that unit's HF tensors, convert
  (fuse q/k/v, interleave gate/up, transpose GPT-2 Conv1D), hand the full tensors to ``ShardedUnit.load_full_state`` which keeps
  this rank's shard.  Peak host memory = one block, every rank reads only what it needs.
* ``load_state_dict(cfg, dir)``: the whole model in our naming (rollout servers, tests).
* ``save_hf(engine_or_state, cfg, dir)``: sharded safetensors + HF ``config.json`` (checkpoint ``save_contents`` ∋ ``hf_model``).

Reference capability: HF AutoModel ``from_pretrained`` + FSDP meta-init / rank-0 broadcast load
(3rdparty/verl/verl/workers/fsdp_workers.py:269-577, verl/utils/fsdp_utils.py:259-486 parallel safetensors load,
``fsdp2_load_full_state_dict``) and the ``hf_model`` checkpoint content (fsdp_checkpoint_manager.py:228-300).
"""
from __future__ import annotations

import json
from pathlib import Path

import torch

from . import transformer as tf
from .config import ModelConfig


class HFCheckpoint:
    def __init__(self, path: str | Path):
        self.dir = Path(path)
        if not self.dir.is_dir():
            raise FileNotFoundError(f"{self.dir} is not a checkpoint directory")
        self.weight_map: dict[str, str] = {}
        idx = self.dir / "model.safetensors.index.json"
        st = sorted(self.dir.glob("*.safetensors"))
        self._bin_cache: dict[str, dict] = {}
        if idx.exists():
            self.weight_map = json.loads(idx.read_text())["weight_map"]
        elif st:
            from safetensors import safe_open

            for f in st:
                with safe_open(str(f), framework="pt") as h:
                    for k in h.keys():
                        self.weight_map[k] = f.name
        else:
            bins = sorted(self.dir.glob("pytorch_model*.bin"))
            if not bins:
                raise FileNotFoundError(f"no *.safetensors or pytorch_model*.bin in {self.dir}")
            for f in bins:
                self._bin_cache[f.name] = torch.load(f, map_location="cpu", weights_only=True)
                for k in self._bin_cache[f.name]:
                    self.weight_map[k] = f.name
        self._handles: dict[str, object] = {}

    def has_weights(self) -> bool:
        return bool(self.weight_map)

    def keys(self):
        return self.weight_map.keys()

    def get(self, name: str) -> torch.Tensor:
        if name not in self.weight_map and name.startswith("model.") and "model.language_model." + name[6:] in self.weight_map:
            name = "model.language_model." + name[6:]   # vision-language checkpoints nest the text stack
        if name not in self.weight_map and name == "lm_head.weight":
            for alias in ("model.embed_tokens.weight", "model.language_model.embed_tokens.weight"):
                if alias in self.weight_map:
                    name = alias                        # tied checkpoints omit the head
                    break
        fname = self.weight_map[name]
        if fname in self._bin_cache:
            return self._bin_cache[fname][name]
        if fname not in self._handles:
            from safetensors import safe_open

            self._handles[fname] = safe_open(str(self.dir / fname), framework="pt")
        return self._handles[fname].get_tensor(name)


def _layer_keys(cfg: ModelConfig, i: int) -> list[str]:
    if cfg.family == "gpt2":
        p = f"transformer.h.{i}."
        return [p + s for s in ("ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "attn.c_attn.weight", "attn.c_attn.bias",
                                "attn.c_proj.weight", "attn.c_proj.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight",
                                "mlp.c_proj.bias")]
    p = f"model.layers.{i}."
    keys = [p + "input_layernorm.weight", p + "post_attention_layernorm.weight", p + "self_attn.o_proj.weight"]
    if cfg.mlp_act == "xielu":               # Apertus naming: its own norm names, no gate projection, activation scalars
        keys = [p + "attention_layernorm.weight", p + "feedforward_layernorm.weight", p + "self_attn.o_proj.weight", p + "mlp.up_proj.weight",
                p + "mlp.down_proj.weight", p + "mlp.act_fn.alpha_p", p + "mlp.act_fn.alpha_n"]
        keys += [p + f"self_attn.{n}_proj.weight" for n in "qkv"] + [p + "self_attn.q_norm.weight", p + "self_attn.k_norm.weight"]
        return keys + ([p + f"self_attn.{n}_proj.bias" for n in "qkv"] if cfg.qkv_bias else [])
    if tf.is_sparse_layer(cfg, i):
        keys += [p + "mlp.gate.weight", p + "mlp.experts.gate_up_proj", p + "mlp.experts.down_proj"]
        if cfg.moe_router == "sigmoid_group":
            keys.append(p + "mlp.gate.e_score_correction_bias")
        if cfg.shared_expert_intermediate_size:
            sh = "mlp.shared_expert." if cfg.shared_expert_gate else "mlp.shared_experts."
            keys += [p + sh + f"{n}_proj.weight" for n in ("gate", "up", "down")]
            if cfg.shared_expert_gate:
                keys.append(p + "mlp.shared_expert_gate.weight")
    elif getattr(cfg, "hf_model_type", "") in ("glm", "glm4", "glm4v"):
        keys += [p + "mlp.gate_up_proj.weight", p + "mlp.down_proj.weight"]
    else:
        keys += [p + "mlp.gate_proj.weight", p + "mlp.up_proj.weight", p + "mlp.down_proj.weight"]
    if cfg.sandwich_norm:
        keys += [p + "post_self_attn_layernorm.weight", p + "post_mlp_layernorm.weight"]
    if cfg.attn_type == "mla":
        a = p + "self_attn."
        q = [a + "q_a_proj.weight", a + "q_a_layernorm.weight", a + "q_b_proj.weight"] if cfg.q_lora_rank else [a + "q_proj.weight"]
        return keys + q + [a + "kv_a_proj_with_mqa.weight", a + "kv_a_layernorm.weight", a + "kv_b_proj.weight"]
    keys += [p + f"self_attn.{n}_proj.weight" for n in "qkv"]
    if cfg.qkv_bias:
        keys += [p + f"self_attn.{n}_proj.bias" for n in "qkv"]
    if cfg.qk_norm:
        keys += [p + "self_attn.q_norm.weight", p + "self_attn.k_norm.weight"]
    return keys


def _one_layer_cfg(cfg: ModelConfig, i: int = 0) -> ModelConfig:
    """A one-block copy of the config whose block 0 has the MLP kind (sparse / dense) of block ``i``."""
    c = ModelConfig(**cfg.to_dict())
    c.num_layers = 1
    c.dense_layers = (0,) if (cfg.num_experts and i in cfg.dense_layers) else ()
    return c


def _convert_layer(cfg: ModelConfig, ck: HFCheckpoint, i: int) -> dict[str, torch.Tensor]:
    """HF tensors of block i -> our names (without the ``layers.i.`` prefix), via the single-layer converter."""
    if cfg.family == "gpt2":
        sub = {k.replace(f"transformer.h.{i}.", "transformer.h.0."): ck.get(k) for k in _layer_keys(cfg, i)}
        sub.update({"transformer.wte.weight": torch.empty(0), "transformer.wpe.weight": torch.empty(0),
                    "transformer.ln_f.weight": torch.empty(0), "transformer.ln_f.bias": torch.empty(0)})
    else:
        keys = _layer_keys(cfg, i)
        p = f"model.layers.{i}."

        def has(k: str) -> bool:
            return k in ck.weight_map or f"model.language_model.{k[6:]}" in ck.weight_map

        if tf.is_sparse_layer(cfg, i) and not has(p + "mlp.experts.gate_up_proj"):
            # checkpoints written before the stacked expert layout: one module per expert
            stacked = {p + "mlp.gate.weight", p + "mlp.experts.gate_up_proj", p + "mlp.experts.down_proj"}
            keys = [k for k in keys if k not in stacked]
            if has(p + "block_sparse_moe.gate.weight"):             # Mixtral: block_sparse_moe.experts.{e}.w1 (gate) / w3 (up) / w2 (down)
                keys += [p + "block_sparse_moe.gate.weight"] + [f"{p}block_sparse_moe.experts.{e}.{n}.weight"
                                                                for e in range(cfg.num_experts) for n in ("w1", "w2", "w3")]
            else:
                keys += [p + "mlp.gate.weight"] + [f"{p}mlp.experts.{e}.{n}_proj.weight" for e in range(cfg.num_experts)
                                                   for n in ("gate", "up", "down")]
        sub = {k.replace(f"model.layers.{i}.", "model.layers.0."): ck.get(k) for k in keys}
        sub.update({"model.embed_tokens.weight": torch.empty(0), "model.norm.weight": torch.empty(0), "lm_head.weight": torch.empty(0)})
    conv = tf.convert_hf_state_dict(_one_layer_cfg(cfg, i), sub)
    return {k[len("layers.0."):]: v for k, v in conv.items() if k.startswith("layers.0.")}


def _embed_head(cfg: ModelConfig, ck: HFCheckpoint) -> tuple[dict, dict]:
    if cfg.family == "gpt2":
        pre = "transformer." if "transformer.wte.weight" in ck.weight_map else ""
        embed = {"embed": ck.get(pre + "wte.weight"), "pos_embed": ck.get(pre + "wpe.weight")}
        head = {"norm_w": ck.get(pre + "ln_f.weight"), "norm_b": ck.get(pre + "ln_f.bias")}
        return embed, head
    embed = {"embed": ck.get("model.embed_tokens.weight")}
    head = {"norm_w": ck.get("model.norm.weight")}
    if cfg.value_head:                       # reward-model / critic checkpoints carry ``score``; language-model checkpoints do not
        if "score.weight" in ck.weight_map:
            head.update(tf.value_head_from_hf(cfg, ck.get("score.weight"), ck.get("score.bias") if "score.bias" in ck.weight_map else None))
    elif not cfg.tie_word_embeddings:
        head["lm_head"] = ck.get("lm_head.weight")
    return embed, head


def _vision_tensors(cfg: ModelConfig, ck: HFCheckpoint) -> dict[str, torch.Tensor]:
    """The tower of a vision-language checkpoint in our naming (``vision.*``)."""
    from .vision import convert_hf_vision_state_dict

    return convert_hf_vision_state_dict(cfg, {k: ck.get(k) for k in ck.keys() if k.startswith(("model.visual.", "visual."))})


def load_state_dict(cfg: ModelConfig, path: str | Path) -> dict[str, torch.Tensor]:
    ck = HFCheckpoint(path)
    embed, head = _embed_head(cfg, ck)
    out = {**embed, **head}
    if cfg.vision is not None:
        out.update(_vision_tensors(cfg, ck))
    for i in range(cfg.num_layers):
        for k, v in _convert_layer(cfg, ck, i).items():
            out[f"layers.{i}.{k}"] = v
    return out


def load_into_engine(engine, path: str | Path) -> int:
    """Fill every ZeRO unit of ``engine`` from the checkpoint; returns the number of tensors loaded.  Value-head engines keep
    their freshly initialised head projection (an LM checkpoint has none)."""
    cfg = engine.cfg
    ck = HFCheckpoint(path)
    embed, head = _embed_head(cfg, ck)
    n = 0
    engine.embed.load_full_state(embed)
    n += len(embed)
    if getattr(engine, "vision", None) is not None:
        vt = _vision_tensors(cfg, ck)
        engine.vision.load_full_state(vt, prefix="vision.")
        n += len(vt)
    for i, unit in enumerate(engine.layers):
        lw = _convert_layer(cfg, ck, i)
        unit.load_full_state(lw)
        n += len(lw)
    have = {name for name, _ in engine.head.specs}
    if have <= set(head):
        engine.head.load_full_state(head)
        n += len(head)
    else:                                     # e.g. critic: load the final norm, keep v_head
        cur = {k: v.clone() for k, v in engine.head.views(engine.head.acquire()).items()}
        engine.head.release()
        cur.update({k: v for k, v in head.items() if k in have})
        engine.head.load_full_state(cur)
        n += len([k for k in head if k in have])
    for u in engine.units:
        u.drop_gathered()
    return n


def hf_config_dict(cfg: ModelConfig) -> dict:
    """HF ``config.json`` of the model.  Value-head models (critics, reward models) are written as the family's
    ``...ForTokenClassification`` with one label, the class the reference loads them with (verl/workers/fsdp_workers.py critic:
    AutoModelForTokenClassification, num_labels = 1)."""
    d = _lm_config_dict(cfg)
    if cfg.value_head and cfg.vision is None and d.get("architectures"):
        d["architectures"] = [d["architectures"][0].replace("ForCausalLM", "ForTokenClassification").replace("LMHeadModel", "ForTokenClassification")]
        d.update(num_labels=1, id2label={"0": "LABEL_0"}, label2id={"LABEL_0": 0}, classifier_dropout=0.0)
    return d


def _lm_config_dict(cfg: ModelConfig) -> dict:
    if cfg.family == "gpt2":
        return {"model_type": "gpt2", "architectures": ["GPT2LMHeadModel"], "vocab_size": cfg.vocab_size, "n_embd": cfg.hidden_size,
                "n_inner": cfg.intermediate_size, "n_layer": cfg.num_layers, "n_head": cfg.num_heads, "n_positions": cfg.max_position_embeddings,
                "layer_norm_epsilon": cfg.rms_norm_eps, "eos_token_id": cfg.eos_token_id, "bos_token_id": cfg.bos_token_id}
    if cfg.vision is not None:
        text = {"vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size,
                "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_heads, "num_key_value_heads": cfg.num_kv_heads,
                "max_position_embeddings": cfg.max_position_embeddings, "rms_norm_eps": cfg.rms_norm_eps, "hidden_act": "silu",
                "tie_word_embeddings": cfg.tie_word_embeddings, "eos_token_id": cfg.eos_token_id, "bos_token_id": cfg.bos_token_id,
                "pad_token_id": cfg.pad_token_id,
                "rope_parameters": {"rope_type": "default", "rope_theta": cfg.rope_theta, "mrope_section": list(cfg.mrope_section)}}
        if cfg.vision.get("arch") == "glm4v":
            v = cfg.vision
            text.update(model_type="glm4v_text", attention_bias=bool(cfg.qkv_bias),
                        rope_parameters={"rope_type": "default", "rope_theta": cfg.rope_theta, "mrope_section": list(cfg.mrope_section),
                                         "partial_rotary_factor": cfg.partial_rotary_factor})
            vis = {"depth": v["depth"], "hidden_size": v["embed_dim"], "hidden_act": "silu", "intermediate_size": v["intermediate_size"],
                   "num_heads": v["num_heads"], "in_channels": v.get("in_channels", 3), "patch_size": v["patch_size"],
                   "image_size": v.get("image_size", int(round(v["num_position_embeddings"] ** 0.5)) * v["patch_size"]),
                   "spatial_merge_size": v.get("spatial_merge_size", 2), "temporal_patch_size": v.get("temporal_patch_size", 2),
                   "out_hidden_size": cfg.hidden_size, "attention_bias": bool(v.get("attention_bias", False)),
                   "rms_norm_eps": v.get("rms_norm_eps", 1e-5)}
            return {"model_type": "glm4v", "architectures": ["Glm4vForConditionalGeneration"], "text_config": text, "vision_config": vis,
                    "image_token_id": cfg.image_token_id, "video_token_id": cfg.video_token_id,
                    "image_start_token_id": cfg.vision_start_token_id, "image_end_token_id": cfg.vision_end_token_id,
                    "tie_word_embeddings": cfg.tie_word_embeddings, "torch_dtype": "bfloat16"}
        if cfg.vision.get("arch") == "qwen3":
            v = cfg.vision
            text.update(model_type="qwen3_vl_text", head_dim=cfg.head_dim, attention_bias=False,
                        rope_parameters={"rope_type": "default", "rope_theta": cfg.rope_theta, "mrope_section": list(cfg.mrope_section),
                                         "mrope_interleaved": bool(cfg.mrope_interleaved)})
            vis = {"depth": v["depth"], "hidden_size": v["embed_dim"], "hidden_act": v.get("hidden_act", "gelu_pytorch_tanh"),
                   "intermediate_size": v["intermediate_size"], "num_heads": v["num_heads"], "in_channels": v.get("in_channels", 3),
                   "patch_size": v["patch_size"], "spatial_merge_size": v.get("spatial_merge_size", 2),
                   "temporal_patch_size": v.get("temporal_patch_size", 2), "out_hidden_size": cfg.hidden_size,
                   "num_position_embeddings": v["num_position_embeddings"],
                   "deepstack_visual_indexes": list(v.get("deepstack_visual_indexes") or ())}
            if cfg.num_experts:
                text.update(model_type="qwen3_vl_moe_text", num_experts=cfg.num_experts, num_experts_per_tok=cfg.num_experts_per_tok,
                            moe_intermediate_size=cfg.moe_intermediate_size, decoder_sparse_step=1, mlp_only_layers=[])
                return {"model_type": "qwen3_vl_moe", "architectures": ["Qwen3VLMoeForConditionalGeneration"], "text_config": text,
                        "vision_config": vis, "image_token_id": cfg.image_token_id, "video_token_id": cfg.video_token_id,
                        "vision_start_token_id": cfg.vision_start_token_id, "vision_end_token_id": cfg.vision_end_token_id,
                        "tie_word_embeddings": cfg.tie_word_embeddings, "torch_dtype": "bfloat16"}
            return {"model_type": "qwen3_vl", "architectures": ["Qwen3VLForConditionalGeneration"], "text_config": text,
                    "vision_config": vis, "image_token_id": cfg.image_token_id, "video_token_id": cfg.video_token_id,
                    "vision_start_token_id": cfg.vision_start_token_id, "vision_end_token_id": cfg.vision_end_token_id,
                    "tie_word_embeddings": cfg.tie_word_embeddings, "torch_dtype": "bfloat16"}
        v25 = cfg.vision.get("arch") == "qwen2_5"
        vis = {k: v for k, v in cfg.vision.items() if k != "arch"}
        if v25:
            vis.update(hidden_size=vis.pop("embed_dim"), out_hidden_size=cfg.hidden_size, hidden_act="silu")
        else:
            vis["hidden_size"] = cfg.hidden_size
        return {"model_type": "qwen2_5_vl" if v25 else "qwen2_vl",
                "architectures": ["Qwen2_5_VLForConditionalGeneration" if v25 else "Qwen2VLForConditionalGeneration"], "text_config": text,
                "vision_config": vis, "image_token_id": cfg.image_token_id,
                "video_token_id": cfg.video_token_id, "vision_start_token_id": cfg.vision_start_token_id,
                "vision_end_token_id": cfg.vision_end_token_id,
                "tie_word_embeddings": cfg.tie_word_embeddings, "torch_dtype": "bfloat16",
                # flat copies for loaders that predate the nested layout
                "rope_scaling": {"type": "mrope", "mrope_section": list(cfg.mrope_section)}, "rope_theta": cfg.rope_theta}
    mt = "qwen3" if cfg.qk_norm else ("qwen2" if cfg.qkv_bias else "llama")
    arch = {"qwen3": "Qwen3ForCausalLM", "qwen2": "Qwen2ForCausalLM", "llama": "LlamaForCausalLM"}[mt]
    if cfg.mlp_act == "xielu":
        rp = {"rope_type": "default", "rope_theta": cfg.rope_theta, **(dict(cfg.rope_scaling) if cfg.rope_scaling else {})}
        return {"model_type": "apertus", "architectures": ["ApertusForCausalLM"], "vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size,
                "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_heads,
                "num_key_value_heads": cfg.num_kv_heads, "head_dim": cfg.head_dim, "max_position_embeddings": cfg.max_position_embeddings,
                "rope_theta": cfg.rope_theta, "rope_parameters": rp, **({"rope_scaling": dict(cfg.rope_scaling)} if cfg.rope_scaling else {}),
                "rms_norm_eps": cfg.rms_norm_eps, "tie_word_embeddings": cfg.tie_word_embeddings, "attention_bias": bool(cfg.qkv_bias),
                "hidden_act": "xielu", "qk_norm": True, "torch_dtype": "bfloat16", "eos_token_id": cfg.eos_token_id,
                "bos_token_id": cfg.bos_token_id, "pad_token_id": cfg.pad_token_id}
    if getattr(cfg, "hf_model_type", "") == "mistral" and mt == "llama":
        mt, arch = "mistral", "MistralForCausalLM"
    moe = {}
    if cfg.sliding_window:
        moe["sliding_window"] = int(cfg.sliding_window)
    if cfg.num_experts and getattr(cfg, "hf_model_type", "") == "mixtral":
        return {"model_type": "mixtral", "architectures": ["MixtralForCausalLM"], "vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size,
                "intermediate_size": cfg.moe_intermediate_size, "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_heads,
                "num_key_value_heads": cfg.num_kv_heads, "head_dim": cfg.head_dim, "max_position_embeddings": cfg.max_position_embeddings,
                "rope_theta": cfg.rope_theta, "rms_norm_eps": cfg.rms_norm_eps, "tie_word_embeddings": cfg.tie_word_embeddings,
                "num_local_experts": cfg.num_experts, "num_experts_per_tok": cfg.num_experts_per_tok, "sliding_window": cfg.sliding_window,
                "hidden_act": "silu", "torch_dtype": "bfloat16", "eos_token_id": cfg.eos_token_id, "bos_token_id": cfg.bos_token_id,
                "pad_token_id": cfg.pad_token_id}
    if getattr(cfg, "hf_model_type", "") in ("glm", "glm4"):
        gmt = cfg.hf_model_type
        return {"model_type": gmt, "architectures": ["GlmForCausalLM" if gmt == "glm" else "Glm4ForCausalLM"], "vocab_size": cfg.vocab_size,
                "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_layers,
                "num_attention_heads": cfg.num_heads, "num_key_value_heads": cfg.num_kv_heads, "head_dim": cfg.head_dim,
                "max_position_embeddings": cfg.max_position_embeddings, "rms_norm_eps": cfg.rms_norm_eps, "attention_bias": bool(cfg.qkv_bias),
                "partial_rotary_factor": cfg.partial_rotary_factor, "rope_theta": cfg.rope_theta,
                "rope_parameters": {"rope_type": "default", "rope_theta": cfg.rope_theta, "partial_rotary_factor": cfg.partial_rotary_factor},
                "tie_word_embeddings": cfg.tie_word_embeddings, "hidden_act": "silu", "torch_dtype": "bfloat16",
                "eos_token_id": [cfg.eos_token_id, *cfg.extra_eos_token_ids] if cfg.extra_eos_token_ids else cfg.eos_token_id,
                "bos_token_id": cfg.bos_token_id, "pad_token_id": cfg.pad_token_id}
    if cfg.attn_type == "mla":           # DeepSeek-V3 architecture (a dense MLA stack is written as first_k_dense_replace = all layers)
        first_dense = 0
        while first_dense in cfg.dense_layers:
            first_dense += 1
        fm = cfg.moe_intermediate_size or cfg.intermediate_size
        rp = {"rope_type": "default", "rope_theta": cfg.rope_theta, **(dict(cfg.rope_scaling) if cfg.rope_scaling else {})}
        if cfg.moe_router == "softmax" and (cfg.num_experts or getattr(cfg, "hf_model_type", "") == "deepseek_v2"):      # DeepSeek-V2 / V2-Lite
            return {"model_type": "deepseek_v2", "architectures": ["DeepseekV2ForCausalLM"], "vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size,
                    "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_heads,
                    "num_key_value_heads": cfg.num_heads, "max_position_embeddings": cfg.max_position_embeddings, "rms_norm_eps": cfg.rms_norm_eps,
                    "attention_bias": False, "q_lora_rank": cfg.q_lora_rank or None, "kv_lora_rank": cfg.kv_lora_rank,
                    "qk_nope_head_dim": cfg.qk_nope_head_dim, "qk_rope_head_dim": cfg.qk_rope_head_dim, "v_head_dim": cfg.v_head_dim,
                    "rope_theta": cfg.rope_theta, "rope_parameters": rp, **({"rope_scaling": dict(cfg.rope_scaling)} if cfg.rope_scaling else {}),
                    "n_routed_experts": cfg.num_experts or None, "num_experts_per_tok": cfg.num_experts_per_tok or None, "moe_intermediate_size": fm,
                    "n_shared_experts": (cfg.shared_expert_intermediate_size // fm) if cfg.num_experts else None,
                    "topk_method": "group_limited_greedy" if cfg.n_group > 1 else "greedy", "n_group": cfg.n_group, "topk_group": cfg.topk_group,
                    "routed_scaling_factor": cfg.routed_scaling_factor, "norm_topk_prob": False,
                    "first_k_dense_replace": first_dense if cfg.num_experts else cfg.num_layers,
                    "tie_word_embeddings": cfg.tie_word_embeddings, "hidden_act": "silu", "torch_dtype": "bfloat16",
                    "eos_token_id": cfg.eos_token_id, "bos_token_id": cfg.bos_token_id, "pad_token_id": cfg.pad_token_id}
        return {"model_type": "deepseek_v3", "architectures": ["DeepseekV3ForCausalLM"], "vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size,
                "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_heads,
                "num_key_value_heads": cfg.num_heads, "max_position_embeddings": cfg.max_position_embeddings, "rms_norm_eps": cfg.rms_norm_eps,
                "attention_bias": False, "q_lora_rank": cfg.q_lora_rank or None, "kv_lora_rank": cfg.kv_lora_rank,
                "qk_nope_head_dim": cfg.qk_nope_head_dim, "qk_rope_head_dim": cfg.qk_rope_head_dim, "v_head_dim": cfg.v_head_dim,
                "rope_interleave": bool(cfg.rope_interleaved), "rope_theta": cfg.rope_theta, "rope_parameters": rp,
                **({"rope_scaling": dict(cfg.rope_scaling)} if cfg.rope_scaling else {}),
                "n_routed_experts": cfg.num_experts or 1, "num_experts_per_tok": cfg.num_experts_per_tok or 1, "moe_intermediate_size": fm,
                "norm_topk_prob": cfg.norm_topk_prob, "n_shared_experts": (cfg.shared_expert_intermediate_size // fm) if cfg.num_experts else 0,
                "n_group": cfg.n_group, "topk_group": cfg.topk_group, "routed_scaling_factor": cfg.routed_scaling_factor,
                "first_k_dense_replace": first_dense if cfg.num_experts else cfg.num_layers,
                "tie_word_embeddings": cfg.tie_word_embeddings, "hidden_act": "silu", "torch_dtype": "bfloat16",
                "eos_token_id": [cfg.eos_token_id, *cfg.extra_eos_token_ids] if cfg.extra_eos_token_ids else cfg.eos_token_id,
                "bos_token_id": cfg.bos_token_id, "pad_token_id": cfg.pad_token_id}
    if cfg.num_experts and getattr(cfg, "hf_model_type", "") == "glm4_moe":
        first_dense = 0
        while first_dense in cfg.dense_layers:
            first_dense += 1
        return {"model_type": "glm4_moe", "architectures": ["Glm4MoeForCausalLM"], "vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size,
                "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_heads,
                "num_key_value_heads": cfg.num_kv_heads, "head_dim": cfg.head_dim, "max_position_embeddings": cfg.max_position_embeddings,
                "rms_norm_eps": cfg.rms_norm_eps, "attention_bias": bool(cfg.qkv_bias), "use_qk_norm": bool(cfg.qk_norm),
                "partial_rotary_factor": cfg.partial_rotary_factor, "rope_theta": cfg.rope_theta,
                "rope_parameters": {"rope_type": "default", "rope_theta": cfg.rope_theta, "partial_rotary_factor": cfg.partial_rotary_factor},
                "n_routed_experts": cfg.num_experts, "num_experts_per_tok": cfg.num_experts_per_tok,
                "moe_intermediate_size": cfg.moe_intermediate_size, "norm_topk_prob": cfg.norm_topk_prob,
                "n_shared_experts": cfg.shared_expert_intermediate_size // cfg.moe_intermediate_size, "n_group": cfg.n_group,
                "topk_group": cfg.topk_group, "routed_scaling_factor": cfg.routed_scaling_factor, "first_k_dense_replace": first_dense,
                "tie_word_embeddings": cfg.tie_word_embeddings, "hidden_act": "silu", "torch_dtype": "bfloat16",
                "eos_token_id": cfg.eos_token_id, "bos_token_id": cfg.bos_token_id, "pad_token_id": cfg.pad_token_id}
    if cfg.num_experts:
        mt, arch = "qwen3_moe", "Qwen3MoeForCausalLM"
        moe = {**moe, "num_experts": cfg.num_experts, "num_experts_per_tok": cfg.num_experts_per_tok, "moe_intermediate_size": cfg.moe_intermediate_size,
               "norm_topk_prob": cfg.norm_topk_prob, "decoder_sparse_step": 1, "mlp_only_layers": [int(i) for i in cfg.dense_layers]}
        if getattr(cfg, "hf_model_type", "") == "qwen2_moe":
            mt, arch = "qwen2_moe", "Qwen2MoeForCausalLM"
            moe.update(shared_expert_intermediate_size=cfg.shared_expert_intermediate_size, qkv_bias=bool(cfg.qkv_bias))
    return {**moe, "model_type": mt, "architectures": [arch], "vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size,
            "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_heads,
            "num_key_value_heads": cfg.num_kv_heads, "head_dim": cfg.head_dim, "max_position_embeddings": cfg.max_position_embeddings,
            "rope_theta": cfg.rope_theta, "rms_norm_eps": cfg.rms_norm_eps, "tie_word_embeddings": cfg.tie_word_embeddings,
            **({"rope_scaling": dict(cfg.rope_scaling)} if cfg.rope_scaling else {}),
            "attention_bias": bool(cfg.qkv_bias and mt == "llama"), "hidden_act": "silu", "torch_dtype": "bfloat16",
            "eos_token_id": cfg.eos_token_id, "bos_token_id": cfg.bos_token_id, "pad_token_id": cfg.pad_token_id}


def save_hf(state_or_engine, cfg: ModelConfig, out_dir: str | Path, max_shard_bytes: int = 4 << 30, dtype=torch.bfloat16) -> list[str]:
    """Write ``config.json`` + sharded safetensors in HF naming.  ``state_or_engine``: our-name state dict or a ZeRO engine (every
    rank must call; rank 0 writes)."""
    from safetensors.torch import save_file

    out = Path(out_dir)
    is_engine = hasattr(state_or_engine, "named_full_tensors")
    rank0 = (not is_engine) or state_or_engine.rank == 0
    sd = {n: t.detach().to("cpu", dtype).clone() for n, t in (state_or_engine.named_full_tensors() if is_engine else state_or_engine.items())
          if ".lora." not in n}
    if not rank0:
        return []
    out.mkdir(parents=True, exist_ok=True)
    hf = {k: v.contiguous() for k, v in tf.export_hf_state_dict(cfg, {k: v for k, v in sd.items() if not k.startswith("vision.")}).items()}
    if cfg.vision is not None:
        from .vision import export_hf_vision_state_dict

        hf = {("model.language_model." + k[6:] if k.startswith("model.") else k): v for k, v in hf.items()}
        hf.update({k: v.contiguous() for k, v in export_hf_vision_state_dict(cfg, sd).items()})
    if cfg.tie_word_embeddings:
        hf.pop("lm_head.weight", None)
    shards, cur, size = [], {}, 0
    for k, v in hf.items():
        nb = v.numel() * v.element_size()
        if cur and size + nb > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = v
        size += nb
    shards.append(cur)
    names, weight_map = [], {}
    for i, sh in enumerate(shards):
        name = "model.safetensors" if len(shards) == 1 else f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, str(out / name), metadata={"format": "pt"})
        names.append(name)
        weight_map.update({k: name for k in sh})
    if len(shards) > 1:
        (out / "model.safetensors.index.json").write_text(json.dumps({"metadata": {}, "weight_map": weight_map}, indent=1))
    hc = hf_config_dict(cfg)
    hc["torch_dtype"] = str(dtype).replace("torch.", "")
    (out / "config.json").write_text(json.dumps(hc, indent=2))
    return names
