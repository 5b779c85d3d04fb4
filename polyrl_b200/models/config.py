"""Model architecture configs (decoder-only transformers) and named presets.

Families: ``llama`` (Llama-2/3, Qwen2/2.5/3 — RMSNorm, RoPE, SwiGLU, GQA, optional qkv bias / qk-norm) and
``gpt2`` (LayerNorm, learned positions, GELU MLP, biases).  Shapes for the presets follow the public
HF configs (the reference loads the same families through HF AutoModel:
3rdparty/verl/verl/workers/fsdp_workers.py:269-577).
"""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass
from pathlib import Path


@dataclass
class ModelConfig:
    family: str = "llama"            # "llama" | "gpt2"
    vocab_size: int = 32000
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_layers: int = 32
    num_heads: int = 32
    num_kv_heads: int = 32
    head_dim: int = 0                # 0 -> hidden_size // num_heads
    max_position_embeddings: int = 8192
    rope_theta: float = 10000.0
    rms_norm_eps: float = 1e-6
    # long-context rotary scaling (HF rope_scaling / rope_parameters): {"rope_type": "llama3" | "linear" | "yarn", "factor": f,
    # "original_max_position_embeddings": n, "low_freq_factor", "high_freq_factor" (llama3), "beta_fast", "beta_slow", "mscale",
    # "attention_factor" (yarn)}; None / "default" = plain RoPE.  Applied once when the cos / sin tables are built.
    rope_scaling: dict | None = None
    # GLM family: only the first ``partial_rotary_factor * head_dim`` features of a head rotate; ``rope_interleaved`` says the HF
    # checkpoint pairs features (2j, 2j+1) — the converters permute the q / k rows so that OUR weights always pair (j, j + d/2) and
    # non-rotating features sit on table entries with cos = 1, sin = 0: the rope / attention kernels stay untouched.
    partial_rotary_factor: float = 1.0
    rope_interleaved: bool = False
    sandwich_norm: bool = False      # GLM-4: extra RMSNorm on the attention output and on the MLP output, before the residual adds
    qkv_bias: bool = False           # Qwen2
    qk_norm: bool = False            # Qwen3: per-head RMSNorm on q and k (weights [head_dim]) before RoPE
    tie_word_embeddings: bool = False
    pad_token_id: int = 0
    eos_token_id: int = 1
    extra_eos_token_ids: tuple = ()  # further stop ids of the HF generation config (e.g. Llama-3 <|eot_id|>)
    bos_token_id: int = 1
    name: str = "custom"
    value_head: bool = False         # critic: scalar head instead of lm_head
    # Mistral-style local attention: every query sees at most ``sliding_window`` keys.  The kernels implement full causal
    # attention, which is the same thing while sequences are no longer than the window (RL rollouts of a few K tokens against
    # Mistral's 4096 / 32768) — longer sequences are rejected instead of silently attending too far (models/transformer.py).
    sliding_window: int | None = None
    hf_model_type: str = ""          # the HF model_type this config was read from (export writes it back)
    # mixture-of-experts MLPs (Qwen3-MoE layout): every block routes each token to ``num_experts_per_tok`` of ``num_experts`` SwiGLU
    # experts of width ``moe_intermediate_size``; 0 experts = dense MLP
    num_experts: int = 0
    num_experts_per_tok: int = 0
    moe_intermediate_size: int = 0
    norm_topk_prob: bool = True
    # MoE variants.  ``dense_layers``: blocks that keep a dense SwiGLU MLP of width intermediate_size (Qwen mlp_only_layers /
    # decoder_sparse_step, DeepSeek / GLM-4.5 first_k_dense_replace).  Shared expert: a dense SwiGLU MLP of width
    # ``shared_expert_intermediate_size`` added to the routed output of every sparse block, optionally scaled by a learned sigmoid
    # gate per token (Qwen2-MoE).  ``moe_router``: "softmax" (Mixtral / Qwen) or "sigmoid_group" (DeepSeek-V3 / GLM-4.5: sigmoid
    # scores, selection on scores + a per-expert bias, experts restricted to the best ``topk_group`` of ``n_group`` groups,
    # weights scaled by ``routed_scaling_factor``).
    dense_layers: tuple = ()
    shared_expert_intermediate_size: int = 0
    shared_expert_gate: bool = False
    moe_router: str = "softmax"
    n_group: int = 1
    topk_group: int = 1
    routed_scaling_factor: float = 1.0
    # Multi-head latent attention (DeepSeek-V2/V3, Moonlight, Kimi text stacks; models/mla.py).  ``attn_type="mla"``: queries come
    # from a low-rank bottleneck (``q_lora_rank``, 0 = a plain projection), keys / values from ONE shared latent of width
    # ``kv_lora_rank`` per token plus ``qk_rope_head_dim`` rotary features shared by all heads; a head has ``qk_nope_head_dim``
    # position-free + ``qk_rope_head_dim`` rotary query / key features (``head_dim`` = their sum) and ``v_head_dim`` value features.
    # The rollout engine caches only the latent (kv_lora_rank + qk_rope_head_dim numbers per token and layer).
    attn_type: str = "mha"
    # dense MLP kind: "swiglu" (gate / up / down, every other family) or "xielu" (Apertus: up -> xIELU -> down with two learned
    # scalars per block; xIELU(x) = a_p x^2 + b x for x > 0, a_n (expm1(min(x, eps)) - x) + b x otherwise)
    mlp_act: str = "swiglu"
    q_lora_rank: int = 0
    kv_lora_rank: int = 0
    qk_nope_head_dim: int = 0
    qk_rope_head_dim: int = 0
    v_head_dim: int = 0
    # vision-language models (Qwen2-VL): a ViT tower feeds merged patch embeddings into the image-token positions and the text
    # stack rotates with 3-axis (t, h, w) M-RoPE.  ``vision`` holds the tower shape (see models/vision.py), None = text only.
    vision: dict | None = None
    mrope_section: tuple = ()        # rotary frequency split between the (t, h, w) position streams, sums to head_dim / 2
    mrope_interleaved: bool = False  # Qwen3-VL: the three streams alternate frequency by frequency instead of owning contiguous chunks
    image_token_id: int = -1
    video_token_id: int = -1
    vision_start_token_id: int = -1
    vision_end_token_id: int = -1

    def __post_init__(self):
        self.extra_eos_token_ids = tuple(self.extra_eos_token_ids or ())        # json round trip gives a list
        self.mrope_section = tuple(self.mrope_section or ())
        self.dense_layers = tuple(int(i) for i in (self.dense_layers or ()))
        if self.mlp_act not in ("swiglu", "xielu") or (self.mlp_act == "xielu" and (self.family != "llama" or self.num_experts)):
            raise ValueError("mlp_act is 'swiglu' or 'xielu' (dense llama-family stacks only)")
        if self.attn_type not in ("mha", "mla"):
            raise ValueError(f"unknown attn_type {self.attn_type!r}")
        if self.attn_type == "mla":
            if self.family != "llama" or min(self.kv_lora_rank, self.qk_nope_head_dim, self.qk_rope_head_dim, self.v_head_dim) <= 0:
                raise ValueError("MLA needs the llama family and kv_lora_rank / qk_nope_head_dim / qk_rope_head_dim / v_head_dim > 0")
            if self.qk_rope_head_dim % 2 or self.qkv_bias or self.qk_norm or self.vision is not None:
                raise ValueError("MLA: qk_rope_head_dim must be even; projection biases, per-head QK norm and vision towers are not supported")
            self.head_dim = self.qk_nope_head_dim + self.qk_rope_head_dim
            self.num_kv_heads = self.num_heads
        if self.head_dim == 0:
            self.head_dim = self.hidden_size // self.num_heads
        if self.num_heads % self.num_kv_heads != 0:
            raise ValueError("num_heads must be a multiple of num_kv_heads")
        if self.family not in ("llama", "gpt2"):
            raise ValueError(f"unknown model family {self.family!r}")
        if self.num_experts:
            if self.family != "llama" or not (0 < self.num_experts_per_tok <= self.num_experts) or self.moe_intermediate_size <= 0:
                raise ValueError("MoE needs the llama family, 0 < num_experts_per_tok <= num_experts and moe_intermediate_size > 0")
            if self.moe_intermediate_size % 128 or self.shared_expert_intermediate_size % 128:
                raise ValueError("moe / shared-expert intermediate sizes must be multiples of 128 (gate / up rows are interleaved in 128-row blocks)")
            if self.moe_router not in ("softmax", "sigmoid_group"):
                raise ValueError(f"unknown moe_router {self.moe_router!r}")
            if self.n_group > 1 and (self.num_experts % self.n_group or self.topk_group > self.n_group):
                raise ValueError("n_group must divide num_experts and topk_group <= n_group")
        if self.vision is not None:
            if self.family != "llama":
                raise ValueError("vision towers attach to the llama family only")
            rot = int(self.head_dim * self.partial_rotary_factor) if self.partial_rotary_factor < 1.0 else self.head_dim
            if not self.mrope_section or sum(self.mrope_section) != rot // 2:
                raise ValueError("mrope_section must split the rotary_dim / 2 rotary frequencies")

    @property
    def q_size(self) -> int:
        return self.num_heads * self.head_dim

    @property
    def kv_size(self) -> int:
        return self.num_kv_heads * self.head_dim

    @property
    def o_in_size(self) -> int:
        """Input width of the attention output projection (heads x value features)."""
        return self.num_heads * (self.v_head_dim if self.attn_type == "mla" else self.head_dim)

    def kv_cache_layout(self) -> tuple[int, int, bool]:
        """(heads, features per head, separate value pool) of one paged KV-cache entry.  MLA caches the shared latent + rotary key
        features once per token; values are a prefix of the same entry, so no second pool exists."""
        if self.attn_type == "mla":
            return 1, self.kv_lora_rank + self.qk_rope_head_dim, False
        return self.num_kv_heads, self.head_dim, True

    def attn_num_params(self) -> int:
        """Attention projection parameters of one block (norm gains of the MLA bottlenecks included)."""
        h = self.hidden_size
        if self.attn_type == "mla":
            qk, r = self.num_heads * self.head_dim, self.kv_lora_rank
            q = (h * self.q_lora_rank + self.q_lora_rank + self.q_lora_rank * qk) if self.q_lora_rank else h * qk
            return q + h * (r + self.qk_rope_head_dim) + r + r * self.num_heads * (self.qk_nope_head_dim + self.v_head_dim) + self.o_in_size * h
        return h * (self.q_size + 2 * self.kv_size) + self.q_size * h

    def num_params(self) -> int:
        h, f, v = self.hidden_size, self.intermediate_size, self.vocab_size
        if self.family == "llama":
            mlp = 3 * h * f if self.mlp_act == "swiglu" else 2 * h * f + 2
            if self.num_experts:
                mlp = self.num_experts * (3 * h * self.moe_intermediate_size + h) + 3 * h * self.shared_expert_intermediate_size
                mlp += (h if self.shared_expert_gate else 0) + (self.num_experts if self.moe_router == "sigmoid_group" else 0)
            per = self.attn_num_params() + mlp + 2 * h
            if self.qk_norm:
                per += 2 * self.head_dim
            if self.qkv_bias:
                per += self.q_size + 2 * self.kv_size
            if self.sandwich_norm:
                per += 2 * h
        else:
            per = 4 * h * h + 2 * h * f + 9 * h + f + 4 * h
        emb = v * h * (1 if self.tie_word_embeddings else 2)
        if self.family == "gpt2":
            emb += self.max_position_embeddings * h
        vis = 0
        if self.vision is not None:
            from .vision import vision_num_params

            vis = vision_num_params(self)
        n_dense = len([i for i in self.dense_layers if i < self.num_layers]) if self.num_experts else 0
        return per * self.num_layers - n_dense * (mlp - 3 * h * f) + emb + h + vis      # (dense blocks of MoE stacks are SwiGLU)

    def num_sparse_layers(self) -> int:
        return (self.num_layers - len([i for i in self.dense_layers if i < self.num_layers])) if self.num_experts else 0

    def num_active_params(self) -> int:
        """Parameters touched per token (MoE: only the routed experts) — the FLOPs-relevant count."""
        if not self.num_experts:
            return self.num_params()
        idle = (self.num_experts - self.num_experts_per_tok) * 3 * self.hidden_size * self.moe_intermediate_size
        return self.num_params() - idle * self.num_sparse_layers()

    def to_dict(self) -> dict:
        return asdict(self)

    def save(self, path: str | Path) -> None:
        Path(path).write_text(json.dumps(self.to_dict(), indent=2))

    @classmethod
    def load(cls, path: str | Path) -> "ModelConfig":
        return cls(**json.loads(Path(path).read_text()))

    @classmethod
    def from_hf_config(cls, hf: dict) -> "ModelConfig":
        """Translate a HuggingFace ``config.json`` dict (Llama / Mistral / Qwen2 / Qwen3 / Qwen3-MoE / Qwen2-VL / GPT-2)."""
        mt = hf.get("model_type", "llama")
        if mt == "gpt2":
            return cls(family="gpt2", vocab_size=hf["vocab_size"], hidden_size=hf["n_embd"],
                       intermediate_size=hf.get("n_inner") or 4 * hf["n_embd"], num_layers=hf["n_layer"],
                       num_heads=hf["n_head"], num_kv_heads=hf["n_head"], max_position_embeddings=hf["n_positions"],
                       rms_norm_eps=hf.get("layer_norm_epsilon", 1e-5), tie_word_embeddings=True,
                       eos_token_id=hf.get("eos_token_id", 50256), bos_token_id=hf.get("bos_token_id", 50256),
                       pad_token_id=hf.get("pad_token_id") or hf.get("eos_token_id", 50256), name=hf.get("_name_or_path", "gpt2"))
        vis_kw: dict = {}
        vl_moe_kw: dict = {}
        if mt == "qwen3_vl_moe":             # Qwen3-VL tower + DeepStack on a Qwen3-MoE text stack (always-normalised top-k routing)
            tc = dict(hf.get("text_config") or {})
            vl_moe_kw = dict(num_experts=tc.get("num_experts") or tc["num_local_experts"], num_experts_per_tok=tc["num_experts_per_tok"],
                             moe_intermediate_size=tc["moe_intermediate_size"], norm_topk_prob=True, dense_layers=_qwen_dense_layers(tc))
            mt = "qwen3_vl"
        if mt == "qwen3_vl":
            vc = dict(hf.get("vision_config") or {})
            top = hf
            hf = {**{k: v for k, v in hf.items() if k not in ("text_config", "vision_config")}, **(hf.get("text_config") or {})}
            rope = hf.get("rope_parameters") or hf.get("rope_scaling") or {}
            hf.setdefault("rope_theta", rope.get("rope_theta", 5000000.0))
            vision = dict(arch="qwen3", depth=vc["depth"], embed_dim=vc["hidden_size"], num_heads=vc["num_heads"],
                          patch_size=vc["patch_size"], spatial_merge_size=vc.get("spatial_merge_size", 2),
                          temporal_patch_size=vc.get("temporal_patch_size", 2), in_channels=vc.get("in_channels", 3),
                          intermediate_size=vc["intermediate_size"], hidden_act=vc.get("hidden_act", "gelu_pytorch_tanh"),
                          num_position_embeddings=vc.get("num_position_embeddings", 2304),
                          deepstack_visual_indexes=list(vc.get("deepstack_visual_indexes", ())))
            vis_kw = dict(vision=vision, mrope_section=tuple(rope.get("mrope_section", ())),
                          mrope_interleaved=bool(rope.get("mrope_interleaved", True)), image_token_id=top.get("image_token_id", -1),
                          video_token_id=top.get("video_token_id", -1), vision_start_token_id=top.get("vision_start_token_id", -1),
                          vision_end_token_id=top.get("vision_end_token_id", -1))
            hf["tie_word_embeddings"] = top.get("tie_word_embeddings", hf.get("tie_word_embeddings", False))
            mt = "qwen3"
        if mt == "glm4v":                    # GLM-4V tower on a GLM-4 text stack (output norms, partial interleaved rotary) with M-RoPE
            vc = dict(hf.get("vision_config") or {})
            top = hf
            hf = {**{k: v for k, v in hf.items() if k not in ("text_config", "vision_config")}, **(hf.get("text_config") or {})}
            rope = hf.get("rope_parameters") or hf.get("rope_scaling") or {}
            hf.setdefault("rope_theta", rope.get("rope_theta", 10000.0))
            side = int(vc.get("image_size", 336)) // int(vc["patch_size"])
            vision = dict(arch="glm4v", depth=vc["depth"], embed_dim=vc["hidden_size"], num_heads=vc["num_heads"], patch_size=vc["patch_size"],
                          spatial_merge_size=vc.get("spatial_merge_size", 2), temporal_patch_size=vc.get("temporal_patch_size", 2),
                          in_channels=vc.get("in_channels", 3), intermediate_size=vc["intermediate_size"],
                          attention_bias=bool(vc.get("attention_bias", False)), rms_norm_eps=vc.get("rms_norm_eps", 1e-5),
                          num_position_embeddings=side * side, image_size=int(vc.get("image_size", 336)))
            # GLM-4V wraps images in <|begin_of_image|> ... <|end_of_image|>; video frames reuse the image placeholder
            vis_kw = dict(vision=vision, mrope_section=tuple(rope.get("mrope_section", (8, 12, 12))),
                          image_token_id=top.get("image_token_id", -1), video_token_id=top.get("video_token_id", -1),
                          vision_start_token_id=top.get("image_start_token_id", -1), vision_end_token_id=top.get("image_end_token_id", -1))
            hf["tie_word_embeddings"] = top.get("tie_word_embeddings", hf.get("tie_word_embeddings", False))
            hf.setdefault("attention_bias", True)
            hf["_glm4v"] = True
            mt = "glm4"
        if mt in ("qwen2_vl", "qwen2_5_vl"):
            # transformers >= 4.5x nests the text stack under text_config; older checkpoints keep it flat
            vc = dict(hf.get("vision_config") or {})
            top = hf
            hf = {**{k: v for k, v in hf.items() if k not in ("text_config", "vision_config")}, **(hf.get("text_config") or {})}
            rope = hf.get("rope_parameters") or hf.get("rope_scaling") or {}
            hf.setdefault("rope_theta", rope.get("rope_theta", 1000000.0))
            vision = {k: vc[k] for k in ("depth", "embed_dim", "num_heads", "patch_size", "spatial_merge_size",
                                         "temporal_patch_size", "in_channels", "mlp_ratio") if k in vc}
            if mt == "qwen2_5_vl":          # tower width is called hidden_size there; windowed attention + SwiGLU MLP
                vision.update(arch="qwen2_5", embed_dim=vc["hidden_size"], intermediate_size=vc["intermediate_size"],
                              window_size=vc.get("window_size", 112), fullatt_block_indexes=list(vc.get("fullatt_block_indexes", ())))
            vis_kw = dict(
                vision=vision,
                mrope_section=tuple(rope.get("mrope_section", ())), image_token_id=top.get("image_token_id", -1),
                video_token_id=top.get("video_token_id", -1), vision_start_token_id=top.get("vision_start_token_id", -1),
                vision_end_token_id=top.get("vision_end_token_id", -1))
            hf["tie_word_embeddings"] = top.get("tie_word_embeddings", hf.get("tie_word_embeddings", False))
            mt = "qwen2"
        glm_kw: dict = {}
        if mt in ("glm", "glm4"):            # partial interleaved rotary, fused gate_up projection; GLM-4 adds the sandwich norms
            rp = hf.get("rope_parameters") or hf.get("rope_scaling") or {}
            default_prf = 1.0 if hf.get("_glm4v") else 0.5            # HF: Glm4vTextConfig rotates the whole head unless told otherwise
            glm_kw = dict(partial_rotary_factor=float(rp.get("partial_rotary_factor") or hf.get("partial_rotary_factor") or default_prf),
                          rope_interleaved=True, sandwich_norm=(mt == "glm4"))
        moe_kw: dict = {}
        if mt == "mixtral":                  # llama attention + top-k of num_local_experts SwiGLU experts of width intermediate_size
            moe_kw = dict(num_experts=hf["num_local_experts"], num_experts_per_tok=hf["num_experts_per_tok"],
                          moe_intermediate_size=hf["intermediate_size"], norm_topk_prob=True)
        if mt in ("qwen3_moe", "qwen2_moe"):
            moe_kw = dict(num_experts=hf.get("num_experts") or hf["num_local_experts"], num_experts_per_tok=hf["num_experts_per_tok"],
                          moe_intermediate_size=hf["moe_intermediate_size"], norm_topk_prob=hf.get("norm_topk_prob", False),
                          dense_layers=_qwen_dense_layers(hf))
            if mt == "qwen2_moe":            # + an always-on shared expert behind a per-token sigmoid gate; qkv bias like Qwen2
                moe_kw.update(shared_expert_intermediate_size=hf.get("shared_expert_intermediate_size", 0),
                              shared_expert_gate=bool(hf.get("shared_expert_intermediate_size", 0)))
                hf = {**hf, "attention_bias": hf.get("qkv_bias", True)}
            mt = "qwen3" if mt == "qwen3_moe" else "qwen2_moe"
        if mt == "glm4_moe":                 # GLM-4.5: DeepSeek-style routing + shared experts, first layers dense, partial (non-interleaved) rotary
            rp = hf.get("rope_parameters") or hf.get("rope_scaling") or {}
            moe_kw = dict(num_experts=hf["n_routed_experts"], num_experts_per_tok=hf["num_experts_per_tok"],
                          moe_intermediate_size=hf["moe_intermediate_size"], norm_topk_prob=hf.get("norm_topk_prob", True),
                          dense_layers=tuple(range(min(int(hf.get("first_k_dense_replace", 0)), hf["num_hidden_layers"]))),
                          shared_expert_intermediate_size=hf["moe_intermediate_size"] * int(hf.get("n_shared_experts") or 0),
                          moe_router="sigmoid_group", n_group=int(hf.get("n_group") or 1), topk_group=int(hf.get("topk_group") or 1),
                          routed_scaling_factor=float(hf.get("routed_scaling_factor", 1.0)))
            glm_kw = dict(partial_rotary_factor=float(rp.get("partial_rotary_factor") or hf.get("partial_rotary_factor") or 0.5))
        act_kw: dict = {}
        if mt == "apertus":                  # Qwen3-style attention (per-head QK norm) + a non-gated xIELU MLP
            if hf.get("hidden_act", "xielu") != "xielu":
                raise NotImplementedError(f"apertus with hidden_act={hf.get('hidden_act')!r}")
            act_kw = dict(mlp_act="xielu")
        mla_kw: dict = {}
        if mt == "deepseek_v2":              # MLA attention (complex-pair rotary = interleaved weights) + softmax routing, optionally group-limited
            if hf.get("attention_bias"):
                raise NotImplementedError("deepseek_v2 checkpoints with attention_bias are not supported")
            mla_kw = dict(attn_type="mla", q_lora_rank=int(hf.get("q_lora_rank") or 0), kv_lora_rank=int(hf["kv_lora_rank"]),
                          qk_nope_head_dim=int(hf["qk_nope_head_dim"]), qk_rope_head_dim=int(hf["qk_rope_head_dim"]),
                          v_head_dim=int(hf["v_head_dim"]), rope_interleaved=True)
            if hf.get("n_routed_experts") and int(hf.get("first_k_dense_replace", 0)) < hf["num_hidden_layers"]:
                grouped = hf.get("topk_method", "greedy") == "group_limited_greedy"
                moe_kw = dict(num_experts=hf["n_routed_experts"], num_experts_per_tok=hf["num_experts_per_tok"],
                              moe_intermediate_size=hf["moe_intermediate_size"], norm_topk_prob=False,
                              dense_layers=tuple(range(min(int(hf.get("first_k_dense_replace", 0)), hf["num_hidden_layers"]))),
                              shared_expert_intermediate_size=hf["moe_intermediate_size"] * int(hf.get("n_shared_experts") or 0),
                              moe_router="softmax", n_group=int(hf.get("n_group") or 1) if grouped else 1,
                              topk_group=int(hf.get("topk_group") or 1) if grouped else 1,
                              routed_scaling_factor=float(hf.get("routed_scaling_factor", 1.0)))
            hf = {**hf, "head_dim": 0, "num_key_value_heads": hf["num_attention_heads"]}
        if mt == "deepseek_v3":              # MLA attention + DeepSeek-V3 routing (sigmoid, bias-corrected, group-limited), shared experts
            if hf.get("attention_bias"):
                raise NotImplementedError("deepseek_v3 checkpoints with attention_bias are not supported")
            mla_kw = dict(attn_type="mla", q_lora_rank=int(hf.get("q_lora_rank") or 0), kv_lora_rank=int(hf["kv_lora_rank"]),
                          qk_nope_head_dim=int(hf["qk_nope_head_dim"]), qk_rope_head_dim=int(hf["qk_rope_head_dim"]),
                          v_head_dim=int(hf["v_head_dim"]), rope_interleaved=bool(hf.get("rope_interleave", True)))
            if hf.get("n_routed_experts") and int(hf.get("first_k_dense_replace", 0)) < hf["num_hidden_layers"]:
                moe_kw = dict(num_experts=hf["n_routed_experts"], num_experts_per_tok=hf["num_experts_per_tok"],
                              moe_intermediate_size=hf["moe_intermediate_size"], norm_topk_prob=hf.get("norm_topk_prob", True),
                              dense_layers=tuple(range(min(int(hf.get("first_k_dense_replace", 0)), hf["num_hidden_layers"]))),
                              shared_expert_intermediate_size=hf["moe_intermediate_size"] * int(hf.get("n_shared_experts") or 0),
                              moe_router="sigmoid_group", n_group=int(hf.get("n_group") or 1), topk_group=int(hf.get("topk_group") or 1),
                              routed_scaling_factor=float(hf.get("routed_scaling_factor", 1.0)))
            hf = {**hf, "head_dim": 0, "num_key_value_heads": hf["num_attention_heads"]}     # HF stores the rotary width as head_dim
        eos = hf.get("eos_token_id", 2)
        extra_eos: tuple = ()
        if isinstance(eos, list):
            eos, extra_eos = eos[0], tuple(int(e) for e in eos[1:])
        return cls(family="llama", vocab_size=hf["vocab_size"], hidden_size=hf["hidden_size"],
                   intermediate_size=hf["intermediate_size"], num_layers=hf["num_hidden_layers"],
                   num_heads=hf["num_attention_heads"], num_kv_heads=hf.get("num_key_value_heads", hf["num_attention_heads"]),
                   head_dim=hf.get("head_dim") or 0, max_position_embeddings=hf.get("max_position_embeddings", 8192),
                   rope_theta=(hf.get("rope_theta") or (hf.get("rope_parameters") or hf.get("rope_scaling") or {}).get("rope_theta") or 10000.0),
                   rope_scaling=_rope_scaling_of(hf), rms_norm_eps=hf.get("rms_norm_eps", 1e-6),
                   qkv_bias=(mt == "qwen2") or bool(hf.get("attention_bias", False)),
                   qk_norm=(mt in ("qwen3", "apertus")) or (mt == "glm4_moe" and bool(hf.get("use_qk_norm", False))),
                   tie_word_embeddings=hf.get("tie_word_embeddings", False), eos_token_id=eos, extra_eos_token_ids=extra_eos,
                   bos_token_id=hf.get("bos_token_id") or eos, pad_token_id=hf.get("pad_token_id") or eos,
                   sliding_window=(hf.get("sliding_window") if (mt in ("mistral", "mixtral") or hf.get("use_sliding_window")) else None),
                   hf_model_type=("glm4v" if hf.get("_glm4v") else mt) if mt in ("mistral", "mixtral", "glm", "glm4", "qwen2_moe", "glm4_moe", "deepseek_v2", "deepseek_v3", "apertus") else "",
                   name=hf.get("_name_or_path", mt), **vis_kw, **{**moe_kw, **vl_moe_kw}, **{**glm_kw, **mla_kw, **act_kw})


def _qwen_dense_layers(hf: dict) -> tuple:
    """Qwen MoE stacks: block i is sparse unless listed in mlp_only_layers or (i + 1) is not a multiple of decoder_sparse_step."""
    step, only = int(hf.get("decoder_sparse_step", 1) or 1), set(hf.get("mlp_only_layers") or ())
    return tuple(i for i in range(hf["num_hidden_layers"]) if i in only or (i + 1) % step != 0)


def _rope_scaling_of(hf: dict) -> dict | None:
    rp = dict(hf.get("rope_scaling") or hf.get("rope_parameters") or {})
    kind = rp.get("rope_type") or rp.get("type") or "default"
    if kind in ("default", "mrope") or not rp:
        return None
    keep = ("factor", "original_max_position_embeddings", "low_freq_factor", "high_freq_factor", "beta_fast", "beta_slow", "mscale",
            "mscale_all_dim", "attention_factor", "truncate")
    out = {"rope_type": kind, **{k: rp[k] for k in keep if rp.get(k) is not None}}
    if "original_max_position_embeddings" not in out and hf.get("original_max_position_embeddings"):
        out["original_max_position_embeddings"] = hf["original_max_position_embeddings"]
    return out


PRESETS: dict[str, dict] = {
    "llama3-8b": dict(family="llama", vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_layers=32,
                      num_heads=32, num_kv_heads=8, max_position_embeddings=8192, rope_theta=500000.0,
                      rms_norm_eps=1e-5, eos_token_id=128001, bos_token_id=128000, pad_token_id=128001),
    "llama3-70b": dict(family="llama", vocab_size=128256, hidden_size=8192, intermediate_size=28672, num_layers=80,
                       num_heads=64, num_kv_heads=8, max_position_embeddings=8192, rope_theta=500000.0,
                       rms_norm_eps=1e-5, eos_token_id=128001, bos_token_id=128000, pad_token_id=128001),
    "mistral-7b": dict(family="llama", vocab_size=32768, hidden_size=4096, intermediate_size=14336, num_layers=32, num_heads=32,
                       num_kv_heads=8, max_position_embeddings=32768, rope_theta=1000000.0, rms_norm_eps=1e-5, eos_token_id=2,
                       bos_token_id=1, pad_token_id=2, hf_model_type="mistral"),
    "qwen2-7b": dict(family="llama", vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_layers=28,
                     num_heads=28, num_kv_heads=4, max_position_embeddings=32768, rope_theta=1000000.0,
                     rms_norm_eps=1e-6, qkv_bias=True, eos_token_id=151643, bos_token_id=151643, pad_token_id=151643),
    "qwen3-1.7b": dict(family="llama", vocab_size=151936, hidden_size=2048, intermediate_size=6144, num_layers=28,
                       num_heads=16, num_kv_heads=8, head_dim=128, max_position_embeddings=40960, qk_norm=True,
                       rope_theta=1000000.0, rms_norm_eps=1e-6, tie_word_embeddings=True, eos_token_id=151645,
                       bos_token_id=151643, pad_token_id=151643),
    "qwen2.5-0.5b": dict(family="llama", vocab_size=151936, hidden_size=896, intermediate_size=4864, num_layers=24,
                         num_heads=14, num_kv_heads=2, max_position_embeddings=32768, rope_theta=1000000.0,
                         rms_norm_eps=1e-6, qkv_bias=True, tie_word_embeddings=True, eos_token_id=151645,
                         bos_token_id=151643, pad_token_id=151643),
    "gpt2-small": dict(family="gpt2", vocab_size=50257, hidden_size=768, intermediate_size=3072, num_layers=12,
                       num_heads=12, num_kv_heads=12, max_position_embeddings=1024, rms_norm_eps=1e-5,
                       tie_word_embeddings=True, eos_token_id=50256, bos_token_id=50256, pad_token_id=50256),
    # tiny configs for tests (2-layer, like the reference's test fixtures
    # 3rdparty/verl/tests/models/test_transformers_ulysses.py:49-70)
    "tiny-llama": dict(family="llama", vocab_size=512, hidden_size=256, intermediate_size=512, num_layers=2,
                       num_heads=4, num_kv_heads=2, head_dim=64, max_position_embeddings=512, rope_theta=10000.0,
                       eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "tiny-qwen3": dict(family="llama", vocab_size=512, hidden_size=256, intermediate_size=512, num_layers=2, num_heads=4,
                       num_kv_heads=2, head_dim=64, max_position_embeddings=512, rope_theta=1000000.0, rms_norm_eps=1e-6,
                       qk_norm=True, tie_word_embeddings=True, eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "tiny-glm4": dict(family="llama", vocab_size=512, hidden_size=256, intermediate_size=512, num_layers=2, num_heads=4, num_kv_heads=2,
                      head_dim=64, max_position_embeddings=512, rope_theta=10000.0, rms_norm_eps=1e-5, qkv_bias=True,
                      partial_rotary_factor=0.5, rope_interleaved=True, sandwich_norm=True, hf_model_type="glm4",
                      eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "glm-4-9b": dict(family="llama", vocab_size=151552, hidden_size=4096, intermediate_size=13696, num_layers=40, num_heads=32,
                     num_kv_heads=2, head_dim=128, max_position_embeddings=131072, rope_theta=10000.0, rms_norm_eps=1.5625e-07,
                     qkv_bias=True, partial_rotary_factor=0.5, rope_interleaved=True, hf_model_type="glm",
                     eos_token_id=151329, extra_eos_token_ids=(151336, 151338), bos_token_id=151329, pad_token_id=151329),
    "glm-4-9b-0414": dict(family="llama", vocab_size=151552, hidden_size=4096, intermediate_size=13696, num_layers=40, num_heads=32,
                          num_kv_heads=2, head_dim=128, max_position_embeddings=32768, rope_theta=10000.0, rms_norm_eps=1e-5,
                          qkv_bias=True, partial_rotary_factor=0.5, rope_interleaved=True, sandwich_norm=True, hf_model_type="glm4",
                          eos_token_id=151329, extra_eos_token_ids=(151336, 151338), bos_token_id=151329, pad_token_id=151329),
    "qwen2-vl-7b": dict(family="llama", vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_layers=28,
                        num_heads=28, num_kv_heads=4, max_position_embeddings=32768, rope_theta=1000000.0,
                        rms_norm_eps=1e-6, qkv_bias=True, eos_token_id=151645, bos_token_id=151643, pad_token_id=151643,
                        vision=dict(depth=32, embed_dim=1280, num_heads=16, patch_size=14, spatial_merge_size=2,
                                    temporal_patch_size=2, in_channels=3, mlp_ratio=4),
                        mrope_section=(16, 24, 24), image_token_id=151655, video_token_id=151656, vision_start_token_id=151652,
                        vision_end_token_id=151653),
    "qwen2.5-vl-7b": dict(family="llama", vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_layers=28,
                          num_heads=28, num_kv_heads=4, max_position_embeddings=128000, rope_theta=1000000.0,
                          rms_norm_eps=1e-6, qkv_bias=True, eos_token_id=151645, bos_token_id=151643, pad_token_id=151643,
                          vision=dict(arch="qwen2_5", depth=32, embed_dim=1280, intermediate_size=3420, num_heads=16, patch_size=14,
                                      spatial_merge_size=2, temporal_patch_size=2, in_channels=3, window_size=112,
                                      fullatt_block_indexes=[7, 15, 23, 31]),
                          mrope_section=(16, 24, 24), image_token_id=151655, video_token_id=151656, vision_start_token_id=151652,
                          vision_end_token_id=151653),
    "tiny-glm4v": dict(family="llama", vocab_size=512, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=4, num_kv_heads=2,
                       head_dim=32, max_position_embeddings=512, rope_theta=10000.0, rms_norm_eps=1e-5, qkv_bias=True,
                       partial_rotary_factor=0.5, rope_interleaved=True, sandwich_norm=True, hf_model_type="glm4v",
                       eos_token_id=1, bos_token_id=2, pad_token_id=0,
                       vision=dict(arch="glm4v", depth=2, embed_dim=64, num_heads=4, patch_size=4, spatial_merge_size=2, temporal_patch_size=2,
                                   in_channels=3, intermediate_size=96, attention_bias=False, rms_norm_eps=1e-5, num_position_embeddings=36,
                                   image_size=24),
                       mrope_section=(2, 3, 3), image_token_id=500, video_token_id=501, vision_start_token_id=499, vision_end_token_id=498),
    "glm-4.1v-9b": dict(family="llama", vocab_size=151552, hidden_size=4096, intermediate_size=13696, num_layers=40, num_heads=32,
                        num_kv_heads=2, head_dim=128, max_position_embeddings=65536, rope_theta=10000.0, rms_norm_eps=1e-5, qkv_bias=True,
                        partial_rotary_factor=0.5, rope_interleaved=True, sandwich_norm=True, hf_model_type="glm4v",
                        eos_token_id=151329, extra_eos_token_ids=(151336, 151338, 151348), bos_token_id=151329, pad_token_id=151329,
                        vision=dict(arch="glm4v", depth=24, embed_dim=1536, num_heads=12, patch_size=14, spatial_merge_size=2,
                                    temporal_patch_size=2, in_channels=3, intermediate_size=13696, attention_bias=False, rms_norm_eps=1e-5,
                                    num_position_embeddings=576, image_size=336),
                        mrope_section=(8, 12, 12), image_token_id=151343, video_token_id=151344, vision_start_token_id=151339,
                        vision_end_token_id=151340),
    "tiny-qwen2-vl": dict(family="llama", vocab_size=512, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=4,
                          num_kv_heads=2, head_dim=32, max_position_embeddings=512, rope_theta=10000.0, rms_norm_eps=1e-6,
                          qkv_bias=True, eos_token_id=1, bos_token_id=2, pad_token_id=0,
                          vision=dict(depth=2, embed_dim=64, num_heads=4, patch_size=4, spatial_merge_size=2,
                                      temporal_patch_size=2, in_channels=3, mlp_ratio=2),
                          mrope_section=(4, 6, 6), image_token_id=500, video_token_id=501, vision_start_token_id=499, vision_end_token_id=498),
    "tiny-qwen3-vl": dict(family="llama", vocab_size=512, hidden_size=128, intermediate_size=256, num_layers=4, num_heads=4,
                          num_kv_heads=2, head_dim=32, max_position_embeddings=512, rope_theta=10000.0, rms_norm_eps=1e-6,
                          qk_norm=True, eos_token_id=1, bos_token_id=2, pad_token_id=0,
                          vision=dict(arch="qwen3", depth=3, embed_dim=64, intermediate_size=96, num_heads=4, patch_size=4,
                                      spatial_merge_size=2, temporal_patch_size=2, in_channels=3, hidden_act="gelu_pytorch_tanh",
                                      num_position_embeddings=36, deepstack_visual_indexes=[0, 2]),
                          mrope_section=(6, 5, 5), mrope_interleaved=True, image_token_id=500, video_token_id=501,
                          vision_start_token_id=499, vision_end_token_id=498),
    "qwen3-vl-8b": dict(family="llama", vocab_size=151936, hidden_size=4096, intermediate_size=12288, num_layers=36, num_heads=32,
                        num_kv_heads=8, head_dim=128, max_position_embeddings=262144, rope_theta=5000000.0, rms_norm_eps=1e-6,
                        qk_norm=True, eos_token_id=151645, bos_token_id=151643, pad_token_id=151643,
                        vision=dict(arch="qwen3", depth=27, embed_dim=1152, intermediate_size=4304, num_heads=16, patch_size=16,
                                    spatial_merge_size=2, temporal_patch_size=2, in_channels=3, hidden_act="gelu_pytorch_tanh",
                                    num_position_embeddings=2304, deepstack_visual_indexes=[8, 16, 24]),
                        mrope_section=(24, 20, 20), mrope_interleaved=True, image_token_id=151655, video_token_id=151656,
                        vision_start_token_id=151652, vision_end_token_id=151653),
    # kernel-friendly VLM test config (head_dim 128 for the CUDA decode / rope kernels)
    "tiny-qwen2-vl-128": dict(family="llama", vocab_size=1024, hidden_size=512, intermediate_size=1024, num_layers=2, num_heads=4,
                              num_kv_heads=2, head_dim=128, max_position_embeddings=1024, rope_theta=10000.0, rms_norm_eps=1e-6,
                              qkv_bias=True, eos_token_id=1, bos_token_id=2, pad_token_id=0,
                              vision=dict(depth=2, embed_dim=128, num_heads=2, patch_size=14, spatial_merge_size=2,
                                          temporal_patch_size=2, in_channels=3, mlp_ratio=2),
                              mrope_section=(16, 24, 24), image_token_id=1000, video_token_id=1001, vision_start_token_id=999,
                              vision_end_token_id=998),
    "qwen3-30b-a3b": dict(family="llama", vocab_size=151936, hidden_size=2048, intermediate_size=6144, num_layers=48, num_heads=32,
                          num_kv_heads=4, head_dim=128, max_position_embeddings=40960, qk_norm=True, rope_theta=1000000.0,
                          rms_norm_eps=1e-6, num_experts=128, num_experts_per_tok=8, moe_intermediate_size=768, norm_topk_prob=True,
                          eos_token_id=151645, bos_token_id=151643, pad_token_id=151643),
    "tiny-qwen3-moe": dict(family="llama", vocab_size=512, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=4,
                           num_kv_heads=2, head_dim=32, max_position_embeddings=512, rope_theta=1000000.0, rms_norm_eps=1e-6,
                           qk_norm=True, num_experts=8, num_experts_per_tok=2, moe_intermediate_size=128, norm_topk_prob=True,
                           eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "tiny-glm4-moe": dict(family="llama", vocab_size=512, hidden_size=128, intermediate_size=256, num_layers=3, num_heads=4, num_kv_heads=2,
                          head_dim=32, max_position_embeddings=512, rope_theta=10000.0, rms_norm_eps=1e-5, qkv_bias=True, qk_norm=True,
                          partial_rotary_factor=0.5, num_experts=8, num_experts_per_tok=2, moe_intermediate_size=128, norm_topk_prob=True,
                          dense_layers=(0,), shared_expert_intermediate_size=256, moe_router="sigmoid_group", n_group=4, topk_group=2,
                          routed_scaling_factor=1.5, hf_model_type="glm4_moe", eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "tiny-qwen2-moe": dict(family="llama", vocab_size=512, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=4, num_kv_heads=2,
                           head_dim=32, max_position_embeddings=512, rope_theta=1000000.0, rms_norm_eps=1e-6, qkv_bias=True,
                           num_experts=8, num_experts_per_tok=2, moe_intermediate_size=128, norm_topk_prob=False,
                           shared_expert_intermediate_size=256, shared_expert_gate=True, hf_model_type="qwen2_moe",
                           eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "qwen1.5-moe-a2.7b": dict(family="llama", vocab_size=151936, hidden_size=2048, intermediate_size=5632, num_layers=24, num_heads=16,
                              num_kv_heads=16, max_position_embeddings=8192, rope_theta=1000000.0, rms_norm_eps=1e-6, qkv_bias=True,
                              num_experts=60, num_experts_per_tok=4, moe_intermediate_size=1408, norm_topk_prob=False,
                              shared_expert_intermediate_size=5632, shared_expert_gate=True, hf_model_type="qwen2_moe",
                              eos_token_id=151643, bos_token_id=151643, pad_token_id=151643),
    "glm-4.5-air": dict(family="llama", vocab_size=151552, hidden_size=4096, intermediate_size=10944, num_layers=46, num_heads=96,
                        num_kv_heads=8, head_dim=128, max_position_embeddings=131072, rope_theta=1000000.0, rms_norm_eps=1e-5,
                        qkv_bias=True, partial_rotary_factor=0.5, num_experts=128, num_experts_per_tok=8, moe_intermediate_size=1408,
                        norm_topk_prob=True, dense_layers=(0,), shared_expert_intermediate_size=1408, moe_router="sigmoid_group",
                        n_group=1, topk_group=1, routed_scaling_factor=1.0, hf_model_type="glm4_moe",
                        eos_token_id=151329, extra_eos_token_ids=(151336, 151338), bos_token_id=151329, pad_token_id=151329),
    # DeepSeek-V3 architecture (MLA + sigmoid group-limited routing); Moonlight-16B-A3B ships under the same model_type
    "deepseek-v3": dict(family="llama", vocab_size=129280, hidden_size=7168, intermediate_size=18432, num_layers=61, num_heads=128,
                        num_kv_heads=128, max_position_embeddings=163840, rope_theta=10000.0, rms_norm_eps=1e-6, attn_type="mla",
                        q_lora_rank=1536, kv_lora_rank=512, qk_nope_head_dim=128, qk_rope_head_dim=64, v_head_dim=128, rope_interleaved=True,
                        rope_scaling=dict(rope_type="yarn", factor=40.0, original_max_position_embeddings=4096, beta_fast=32.0, beta_slow=1.0,
                                          mscale=1.0, mscale_all_dim=1.0),
                        num_experts=256, num_experts_per_tok=8, moe_intermediate_size=2048, norm_topk_prob=True, dense_layers=(0, 1, 2),
                        shared_expert_intermediate_size=2048, moe_router="sigmoid_group", n_group=8, topk_group=4, routed_scaling_factor=2.5,
                        hf_model_type="deepseek_v3", eos_token_id=1, bos_token_id=0, pad_token_id=1),
    "moonlight-16b-a3b": dict(family="llama", vocab_size=163840, hidden_size=2048, intermediate_size=11264, num_layers=27, num_heads=16,
                              num_kv_heads=16, max_position_embeddings=8192, rope_theta=50000.0, rms_norm_eps=1e-5, attn_type="mla",
                              q_lora_rank=0, kv_lora_rank=512, qk_nope_head_dim=128, qk_rope_head_dim=64, v_head_dim=128, rope_interleaved=True,
                              num_experts=64, num_experts_per_tok=6, moe_intermediate_size=1408, norm_topk_prob=True, dense_layers=(0,),
                              shared_expert_intermediate_size=2816, moe_router="sigmoid_group", n_group=1, topk_group=1,
                              routed_scaling_factor=2.446, hf_model_type="deepseek_v3", eos_token_id=163585, bos_token_id=163584,
                              pad_token_id=163839),
    "deepseek-v2-lite": dict(family="llama", vocab_size=102400, hidden_size=2048, intermediate_size=10944, num_layers=27, num_heads=16,
                             num_kv_heads=16, max_position_embeddings=163840, rope_theta=10000.0, rms_norm_eps=1e-6, attn_type="mla",
                             q_lora_rank=0, kv_lora_rank=512, qk_nope_head_dim=128, qk_rope_head_dim=64, v_head_dim=128, rope_interleaved=True,
                             rope_scaling=dict(rope_type="yarn", factor=40.0, original_max_position_embeddings=4096, beta_fast=32.0, beta_slow=1.0,
                                               mscale=0.707, mscale_all_dim=0.707),
                             num_experts=64, num_experts_per_tok=6, moe_intermediate_size=1408, norm_topk_prob=False, dense_layers=(0,),
                             shared_expert_intermediate_size=2816, moe_router="softmax", routed_scaling_factor=1.0,
                             hf_model_type="deepseek_v2", eos_token_id=100001, bos_token_id=100000, pad_token_id=100001),
    "tiny-deepseek-v3": dict(family="llama", vocab_size=512, hidden_size=128, intermediate_size=256, num_layers=3, num_heads=4, num_kv_heads=4,
                             max_position_embeddings=512, rope_theta=10000.0, rms_norm_eps=1e-6, attn_type="mla", q_lora_rank=48,
                             kv_lora_rank=32, qk_nope_head_dim=32, qk_rope_head_dim=16, v_head_dim=32, rope_interleaved=True,
                             num_experts=8, num_experts_per_tok=2, moe_intermediate_size=128, norm_topk_prob=True, dense_layers=(0,),
                             shared_expert_intermediate_size=128, moe_router="sigmoid_group", n_group=4, topk_group=2,
                             routed_scaling_factor=2.5, hf_model_type="deepseek_v3", eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "apertus-8b": dict(family="llama", vocab_size=131072, hidden_size=4096, intermediate_size=21504, num_layers=32, num_heads=32,
                       num_kv_heads=8, head_dim=128, max_position_embeddings=65536, rope_theta=12000000.0, rms_norm_eps=1e-5, qk_norm=True,
                       mlp_act="xielu", hf_model_type="apertus",
                       rope_scaling=dict(rope_type="llama3", factor=8.0, original_max_position_embeddings=8192, low_freq_factor=1.0,
                                         high_freq_factor=4.0),
                       eos_token_id=68, bos_token_id=1, pad_token_id=3),
    "tiny-apertus": dict(family="llama", vocab_size=512, hidden_size=256, intermediate_size=512, num_layers=2, num_heads=4, num_kv_heads=2,
                         head_dim=64, max_position_embeddings=512, rope_theta=10000.0, rms_norm_eps=1e-5, qk_norm=True, mlp_act="xielu",
                         hf_model_type="apertus", eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "tiny-qwen3-moe-128": dict(family="llama", vocab_size=1024, hidden_size=512, intermediate_size=1024, num_layers=2, num_heads=4,
                               num_kv_heads=2, head_dim=128, max_position_embeddings=1024, rope_theta=1000000.0, rms_norm_eps=1e-6,
                               qk_norm=True, num_experts=8, num_experts_per_tok=2, moe_intermediate_size=256, norm_topk_prob=True,
                               eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "tiny-llama-128": dict(family="llama", vocab_size=1024, hidden_size=512, intermediate_size=1024, num_layers=2,
                           num_heads=4, num_kv_heads=2, head_dim=128, max_position_embeddings=1024,
                           rope_theta=10000.0, eos_token_id=1, bos_token_id=2, pad_token_id=0),
    "tiny-gpt2": dict(family="gpt2", vocab_size=512, hidden_size=128, intermediate_size=512, num_layers=2,
                      num_heads=4, num_kv_heads=4, max_position_embeddings=256, rms_norm_eps=1e-5,
                      tie_word_embeddings=True, eos_token_id=1, bos_token_id=2, pad_token_id=0),
}


def get_model_config(name_or_path: str, **overrides) -> ModelConfig:
    """Resolve a preset name, a directory containing ``config.json`` (HF or ours), or a json path."""
    if name_or_path in PRESETS:
        cfg = dict(PRESETS[name_or_path], name=name_or_path)
        cfg.update(overrides)
        return ModelConfig(**cfg)
    p = Path(name_or_path)
    if p.is_dir():
        p = p / "config.json"
    if p.exists():
        d = json.loads(p.read_text())
        cfg = ModelConfig(**d) if "family" in d else ModelConfig.from_hf_config(d)
        for k, v in overrides.items():
            setattr(cfg, k, v)
        return cfg
    raise ValueError(f"unknown model {name_or_path!r}; presets: {sorted(PRESETS)}")
