"""Vision-language support (Qwen2-VL, Qwen2.5-VL and Qwen3-VL families): ViT tower, patch merger, DeepStack, M-RoPE positions,
image preprocessing.

The reference trains VLMs by monkey-patching the HF Qwen2-VL classes (3rdparty/verl/verl/models/transformers/qwen2_vl.py:
``get_rope_index`` :36-150, flash-attention forward with 3-axis rotary :153-318, ``process_position_ids``) and feeds images through
the HF processor in the dataset (verl/utils/dataset/rl_dataset.py:181-230, 299-355, vision_utils.py:25-125).  Here the tower is a
functional model over plain weight dicts like the text stack (models/transformer.py), so it can be sharded by the ZeRO engine as one
more unit and run by the rollout engine at prefill:

  pixel patches [N, C*Tp*P*P] --linear--> [N, E] --depth x (LN, full attention inside every frame with 2-D rotary, LN, quick-GELU MLP)
  --> LN --> groups of merge^2 neighbouring patches concatenated --> Linear, GELU, Linear --> [N / merge^2, hidden]

and the merged rows replace the embeddings at the image-token positions of the prompt.  The text stack then rotates q/k with
M-RoPE: rotary frequency i reads the temporal, height or width position stream depending on ``mrope_section``; text tokens carry the
same value on all three streams, so decode needs only a per-sequence position offset (``delta``).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .config import ModelConfig

ParamSpec = list[tuple[str, tuple[int, ...]]]
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


# --------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------
def _v(cfg: ModelConfig) -> dict:
    if cfg.vision is None:
        raise ValueError(f"model {cfg.name!r} has no vision tower")
    v = dict(cfg.vision)
    v.setdefault("in_channels", 3)
    v.setdefault("temporal_patch_size", 2)
    v.setdefault("spatial_merge_size", 2)
    v.setdefault("mlp_ratio", 4)
    return v


def patch_dim(cfg: ModelConfig) -> int:
    v = _v(cfg)
    return v["in_channels"] * v["temporal_patch_size"] * v["patch_size"] ** 2


def is_v25(cfg: ModelConfig) -> bool:
    """Qwen2.5-VL tower: RMSNorm instead of LayerNorm, biased SwiGLU MLP, attention inside spatial windows except for the
    ``fullatt_block_indexes`` blocks."""
    return _v(cfg).get("arch", "qwen2") == "qwen2_5"


def is_v3(cfg: ModelConfig) -> bool:
    """Qwen3-VL tower: LayerNorm blocks with a plain GELU-tanh MLP, biased patch embedding, learned position embeddings that are
    bilinearly interpolated to the image grid, full attention inside a frame, and DeepStack — the outputs of the
    ``deepstack_visual_indexes`` blocks go through their own mergers and are added to the hidden states of the first decoder
    layers at the image-token positions (reference: 3rdparty/verl/verl/models/transformers/qwen3_vl.py)."""
    return _v(cfg).get("arch", "qwen2") == "qwen3"


def is_glm4v(cfg: ModelConfig) -> bool:
    """GLM-4V tower: bias-free RMSNorm / SwiGLU blocks with full attention inside a frame, RMSNorm after the patch embedding, a
    learned position table resampled BICUBICALLY at every patch centre, a strided-convolution merge of each merge^2 cell and a
    gated MLP merger (reference: 3rdparty/verl/verl/models/transformers/glm4v.py)."""
    return _v(cfg).get("arch", "qwen2") == "glm4v"


def deepstack_levels(cfg: ModelConfig) -> int:
    if cfg.vision is None or cfg.vision.get("arch") != "qwen3":
        return 0
    return len(cfg.vision.get("deepstack_visual_indexes") or ())


def embed_width(cfg: ModelConfig) -> int:
    """Columns of the tower output: the merged embedding followed by one block of ``hidden_size`` columns per DeepStack level."""
    return cfg.hidden_size * (1 + deepstack_levels(cfg))


def vision_block_specs(cfg: ModelConfig) -> ParamSpec:
    v = _v(cfg)
    e = v["embed_dim"]
    if is_v3(cfg):
        f = int(v["intermediate_size"])
        return [("ln1_w", (e,)), ("ln1_b", (e,)), ("wqkv", (3 * e, e)), ("bqkv", (3 * e,)), ("wo", (e, e)), ("bo", (e,)),
                ("ln2_w", (e,)), ("ln2_b", (e,)), ("wfc", (f, e)), ("bfc", (f,)), ("wproj", (e, f)), ("bproj", (e,))]
    if is_glm4v(cfg):
        f = int(cfg.hidden_size)                       # the block MLP is as wide as the tower OUTPUT (HF: out_hidden_size)
        bias = [("bqkv", (3 * e,))] if v.get("attention_bias") else []
        return [("ln1_w", (e,)), ("wqkv", (3 * e, e)), *bias, ("wo", (e, e)), ("ln2_w", (e,)), ("wg", (f, e)), ("wu", (f, e)),
                ("wproj", (e, f))]
    if is_v25(cfg):
        f = int(v["intermediate_size"])
        return [("ln1_w", (e,)), ("wqkv", (3 * e, e)), ("bqkv", (3 * e,)), ("wo", (e, e)), ("bo", (e,)), ("ln2_w", (e,)),
                ("wg", (f, e)), ("bg", (f,)), ("wu", (f, e)), ("bu", (f,)), ("wproj", (e, f)), ("bproj", (e,))]
    f = int(e * v["mlp_ratio"])
    return [("ln1_w", (e,)), ("ln1_b", (e,)), ("wqkv", (3 * e, e)), ("bqkv", (3 * e,)), ("wo", (e, e)), ("bo", (e,)),
            ("ln2_w", (e,)), ("ln2_b", (e,)), ("wfc", (f, e)), ("bfc", (f,)), ("wproj", (e, f)), ("bproj", (e,))]


def vision_stem_specs(cfg: ModelConfig) -> ParamSpec:
    """Patch embedding + merger (everything outside the blocks)."""
    v = _v(cfg)
    e, m = v["embed_dim"], v["embed_dim"] * v["spatial_merge_size"] ** 2
    if is_glm4v(cfg):
        o, c = cfg.hidden_size, int(v["intermediate_size"])
        return [("patch_embed", (e, patch_dim(cfg))), ("patch_embed_b", (e,)), ("post_conv_ln_w", (e,)),
                ("pos_embed", (int(v["num_position_embeddings"]), e)), ("post_ln_w", (e,)), ("downsample_w", (o, m)), ("downsample_b", (o,)),
                ("merger_proj", (o, o)), ("merger_ln_w", (o,)), ("merger_ln_b", (o,)), ("merger_wg", (c, o)), ("merger_wu", (c, o)),
                ("merger_wd", (o, c))]
    ln = [("merger_ln_w", (e,))] if is_v25(cfg) else [("merger_ln_w", (e,)), ("merger_ln_b", (e,))]
    base = [("patch_embed", (e, patch_dim(cfg))), *ln, ("merger_w1", (m, m)),
            ("merger_b1", (m,)), ("merger_w2", (cfg.hidden_size, m)), ("merger_b2", (cfg.hidden_size,))]
    if not is_v3(cfg):
        return base
    extra = [("patch_embed_b", (e,)), ("pos_embed", (int(v["num_position_embeddings"]), e))]
    for k in range(deepstack_levels(cfg)):           # DeepStack mergers normalise AFTER the merge-block shuffle (width m)
        extra += [(f"ds{k}_ln_w", (m,)), (f"ds{k}_ln_b", (m,)), (f"ds{k}_w1", (m, m)), (f"ds{k}_b1", (m,)),
                  (f"ds{k}_w2", (cfg.hidden_size, m)), (f"ds{k}_b2", (cfg.hidden_size,))]
    return base + extra


def vision_unit_specs(cfg: ModelConfig) -> ParamSpec:
    """The whole tower as ONE flat spec list (a single ZeRO unit named ``vision``): stem names + ``blocks.{i}.{name}``."""
    specs = list(vision_stem_specs(cfg))
    for i in range(_v(cfg)["depth"]):
        specs += [(f"blocks.{i}.{n}", shp) for n, shp in vision_block_specs(cfg)]
    return specs


class _TowerView:
    """``.stem`` / ``.blocks`` view over a flat name -> tensor dict (a gathered ZeRO unit, or a state dict with a prefix)."""

    def __init__(self, cfg: ModelConfig, flat: dict[str, torch.Tensor], prefix: str = ""):
        self.stem = {n: flat[prefix + n] for n, _ in vision_stem_specs(cfg)}
        self.blocks = [{n: flat[f"{prefix}blocks.{i}.{n}"] for n, _ in vision_block_specs(cfg)} for i in range(_v(cfg)["depth"])]


def tower_view(cfg: ModelConfig, flat: dict[str, torch.Tensor], prefix: str = "") -> _TowerView:
    return _TowerView(cfg, flat, prefix)


def vision_num_params(cfg: ModelConfig) -> int:
    n = sum(math.prod(s) for _, s in vision_stem_specs(cfg))
    return n + _v(cfg)["depth"] * sum(math.prod(s) for _, s in vision_block_specs(cfg))


def _init(name: str, t: torch.Tensor, g) -> None:
    if name.endswith("_w") and "ln" in name:
        t.fill_(1.0)
    elif name.startswith("b") or name.endswith("_b") or name.startswith("merger_b") or (name.startswith("ds") and "_b" in name):
        t.zero_()
    else:
        t.normal_(0.0, 0.02, generator=g)


class VisionWeights:
    """Plain (unsharded) tower weights: ``stem`` dict + ``blocks`` list of dicts."""

    def __init__(self, cfg: ModelConfig, device="cpu", dtype=torch.float32, init: bool = True, seed: int = 1):
        self.cfg = cfg
        g = torch.Generator(device=device)
        g.manual_seed(seed)

        def make(specs):
            out = {}
            for n, shp in specs:
                t = torch.empty(shp, device=device, dtype=dtype)
                if init:
                    _init(n, t, g)
                out[n] = t
            return out

        self.stem = make(vision_stem_specs(cfg))
        self.blocks = [make(vision_block_specs(cfg)) for _ in range(_v(cfg)["depth"])]

    def named_tensors(self):
        for n, t in self.stem.items():
            yield f"vision.{n}", t
        for i, b in enumerate(self.blocks):
            for n, t in b.items():
                yield f"vision.blocks.{i}.{n}", t

    def state_dict(self) -> dict[str, torch.Tensor]:
        return dict(self.named_tensors())

    def load_state_dict(self, sd: dict[str, torch.Tensor]) -> None:
        for n, t in self.named_tensors():
            t.copy_(sd[n])

    def requires_grad_(self, flag: bool = True):
        for _, t in self.named_tensors():
            t.requires_grad_(flag)
        return self


# --------------------------------------------------------------------------------------------
# tower forward
# --------------------------------------------------------------------------------------------
def _grid_list(grid_thw) -> list[tuple[int, int, int]]:
    if isinstance(grid_thw, torch.Tensor):
        grid_thw = grid_thw.tolist()
    return [tuple(int(x) for x in g) for g in grid_thw]


def vision_rotary(cfg: ModelConfig, grid_thw, device) -> tuple[torch.Tensor, torch.Tensor]:
    """2-D rotary tables for every patch, [N, head_dim / 2] cos and sin: the first half of the frequencies reads the patch row, the
    second half the patch column.  Patches are ordered merge-block by merge-block (the order the preprocessor emits)."""
    v = _v(cfg)
    m = v["spatial_merge_size"]
    hd = v["embed_dim"] // v["num_heads"]
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32, device=device) / (hd // 2)))      # [hd/4]
    rows, cols = [], []
    for t, h, w in _grid_list(grid_thw):
        r = torch.arange(h, device=device).view(h // m, m, 1, 1).expand(h // m, m, w // m, m)
        c = torch.arange(w, device=device).view(1, 1, w // m, m).expand(h // m, m, w // m, m)
        rows.append(r.permute(0, 2, 1, 3).reshape(-1).repeat(t))
        cols.append(c.permute(0, 2, 1, 3).reshape(-1).repeat(t))
    r, c = torch.cat(rows).float(), torch.cat(cols).float()
    ang = torch.cat([torch.outer(r, inv), torch.outer(c, inv)], dim=-1)          # [N, hd/2]
    return ang.cos(), ang.sin()


def _rot_half(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x:[N, heads, hd] rotate-half with per-row tables [N, hd/2] (fp32 math)."""
    xf = x.float()
    half = xf.shape[-1] // 2
    x1, x2 = xf[..., :half], xf[..., half:]
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).to(x.dtype)


def _rms(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    xf = x.float()
    return (w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps))).to(x.dtype)


def vision_block(cfg: ModelConfig, w: dict, x: torch.Tensor, cos, sin, cu: torch.Tensor, max_len: int) -> torch.Tensor:
    from ..ops.attention import attention_varlen

    v = _v(cfg)
    e, nh = v["embed_dim"], v["num_heads"]
    n = x.shape[0]
    v25 = "ln1_b" not in w                          # RMSNorm + SwiGLU blocks: Qwen2.5-VL (biased) and GLM-4V (bias-free)
    eps = float(v.get("rms_norm_eps", 1e-6))
    h = _rms(x, w["ln1_w"], eps) if v25 else F.layer_norm(x, (e,), w["ln1_w"], w["ln1_b"], 1e-6)
    qkv = F.linear(h, w["wqkv"], w.get("bqkv")).view(n, 3, nh, e // nh)
    q, k = _rot_half(qkv[:, 0], cos, sin), _rot_half(qkv[:, 1], cos, sin)
    a = attention_varlen(q.contiguous(), k.contiguous(), qkv[:, 2].contiguous(), cu, max_len, causal=False)
    x = x + F.linear(a.reshape(n, e), w["wo"], w.get("bo"))
    if v25:
        h = _rms(x, w["ln2_w"], eps)
        return x + F.linear(F.silu(F.linear(h, w["wg"], w.get("bg"))) * F.linear(h, w["wu"], w.get("bu")), w["wproj"], w.get("bproj"))
    h = F.layer_norm(x, (e,), w["ln2_w"], w["ln2_b"], 1e-6)
    h = F.linear(h, w["wfc"], w["bfc"])
    if v.get("arch") == "qwen3":
        act = v.get("hidden_act", "gelu_pytorch_tanh")
        h = F.gelu(h, approximate="tanh") if act == "gelu_pytorch_tanh" else (F.gelu(h) if act == "gelu" else F.silu(h))
        return x + F.linear(h, w["wproj"], w["bproj"])
    return x + F.linear(h * torch.sigmoid(1.702 * h), w["wproj"], w["bproj"])          # quick-GELU


def interpolated_pos_embed(cfg: ModelConfig, table: torch.Tensor, grid_thw) -> torch.Tensor:
    """Qwen3-VL: the learned [side x side] position table bilinearly resampled to every image's (h, w) patch grid, repeated over
    its frames and emitted in the merge-block patch order of the preprocessor -> [N, E]."""
    v = _v(cfg)
    m = v["spatial_merge_size"]
    side = int(round(table.shape[0] ** 0.5))
    out = []
    for t, h, w in _grid_list(grid_thw):
        ys = torch.linspace(0, side - 1, h, device=table.device)
        xs = torch.linspace(0, side - 1, w, device=table.device)
        y0, x0 = ys.long(), xs.long()
        y1, x1 = (y0 + 1).clamp(max=side - 1), (x0 + 1).clamp(max=side - 1)
        dy, dx = (ys - y0).to(table.dtype), (xs - x0).to(table.dtype)
        tab = table.view(side, side, -1)
        pe = (tab[y0][:, x0] * ((1 - dy)[:, None] * (1 - dx)[None, :]).unsqueeze(-1)
              + tab[y0][:, x1] * ((1 - dy)[:, None] * dx[None, :]).unsqueeze(-1)
              + tab[y1][:, x0] * (dy[:, None] * (1 - dx)[None, :]).unsqueeze(-1)
              + tab[y1][:, x1] * (dy[:, None] * dx[None, :]).unsqueeze(-1))                      # [h, w, E]
        pe = pe.view(h // m, m, w // m, m, -1).permute(0, 2, 1, 3, 4).reshape(h * w, -1)
        out.append(pe.repeat(t, 1))
    return torch.cat(out, dim=0)


def bicubic_pos_embed(cfg: ModelConfig, table: torch.Tensor, grid_thw) -> torch.Tensor:
    """GLM-4V: the learned [side x side] position table sampled bicubically at the centre of every patch of every image (patch
    (r, c) of an h x w grid reads the table at ((c + .5) / w, (r + .5) / h), border-clamped), in the preprocessor's patch order."""
    m = _v(cfg)["spatial_merge_size"]
    side = int(round(table.shape[0] ** 0.5))
    rows, cols, hs, ws = [], [], [], []
    for t, h, w in _grid_list(grid_thw):
        r = torch.arange(h).view(h // m, m, 1, 1).expand(h // m, m, w // m, m).permute(0, 2, 1, 3).reshape(-1).repeat(t)
        c = torch.arange(w).view(1, 1, w // m, m).expand(h // m, m, w // m, m).permute(0, 2, 1, 3).reshape(-1).repeat(t)
        rows.append(r), cols.append(c), hs.append(torch.full_like(r, h)), ws.append(torch.full_like(c, w))
    r, c, hh, ww = (torch.cat(z).to(device=table.device, dtype=torch.float32) for z in (rows, cols, hs, ws))
    grid = torch.stack(((c + 0.5) / ww * 2 - 1, (r + 0.5) / hh * 2 - 1), dim=-1).view(1, -1, 1, 2)
    tab = table.float().view(side, side, -1).permute(2, 0, 1).unsqueeze(0)
    out = F.grid_sample(tab, grid, mode="bicubic", align_corners=False, padding_mode="border")
    return out[0, :, :, 0].t().to(table.dtype)


def vision_stem_in(cfg: ModelConfig, stem: dict, pixel_values: torch.Tensor, grid_thw=None) -> torch.Tensor:
    x = F.linear(pixel_values.to(stem["patch_embed"].dtype), stem["patch_embed"], stem.get("patch_embed_b"))
    if "post_conv_ln_w" in stem:                     # GLM-4V
        if grid_thw is None:
            raise ValueError("the GLM-4V tower needs the image grids for its position embeddings")
        x = _rms(x, stem["post_conv_ln_w"], float(_v(cfg).get("rms_norm_eps", 1e-5)))
        return x + bicubic_pos_embed(cfg, stem["pos_embed"], grid_thw).to(x.dtype)
    if "pos_embed" in stem:
        if grid_thw is None:
            raise ValueError("the Qwen3-VL tower needs the image grids for its position embeddings")
        x = x + interpolated_pos_embed(cfg, stem["pos_embed"], grid_thw).to(x.dtype)
    return x


def deepstack_merge(cfg: ModelConfig, stem: dict, k: int, x: torch.Tensor) -> torch.Tensor:
    """DeepStack merger k on an intermediate block output: merge-block rows concatenated, LayerNorm over the merged width, MLP."""
    v = _v(cfg)
    mdim = v["embed_dim"] * v["spatial_merge_size"] ** 2
    h = F.layer_norm(x.view(-1, mdim), (mdim,), stem[f"ds{k}_ln_w"], stem[f"ds{k}_ln_b"], 1e-6)
    return F.linear(F.gelu(F.linear(h, stem[f"ds{k}_w1"], stem[f"ds{k}_b1"])), stem[f"ds{k}_w2"], stem[f"ds{k}_b2"])


def vision_stem_out(cfg: ModelConfig, stem: dict, x: torch.Tensor) -> torch.Tensor:
    v = _v(cfg)
    e, m2 = v["embed_dim"], v["spatial_merge_size"] ** 2
    if "downsample_w" in stem:                       # GLM-4V: norm, merge^2 cell -> one row by a strided convolution, gated merger
        x = _rms(x, stem["post_ln_w"], float(v.get("rms_norm_eps", 1e-5)))
        cells = x.view(-1, m2, e).transpose(1, 2).reshape(-1, e * m2)            # conv weight order: (channel, row, col)
        y = F.linear(cells, stem["downsample_w"], stem["downsample_b"])
        y = F.linear(y, stem["merger_proj"])
        y = F.gelu(F.layer_norm(y, (y.shape[-1],), stem["merger_ln_w"], stem["merger_ln_b"], 1e-5))
        return F.linear(F.silu(F.linear(y, stem["merger_wg"])) * F.linear(y, stem["merger_wu"]), stem["merger_wd"])
    h = _rms(x, stem["merger_ln_w"]) if "merger_ln_b" not in stem else F.layer_norm(x, (e,), stem["merger_ln_w"], stem["merger_ln_b"], 1e-6)
    h = h.view(-1, e * m2)
    return F.linear(F.gelu(F.linear(h, stem["merger_w1"], stem["merger_b1"])), stem["merger_w2"], stem["merger_b2"])


def window_partition(cfg: ModelConfig, grid_thw) -> tuple[torch.Tensor, list[int]]:
    """Qwen2.5-VL window attention layout.  Merged cells (groups of merge^2 patches) are re-ordered so that the cells of one
    ``window_size`` pixel window are contiguous; returns (order [cells] — new position -> original cell, window lengths in PATCHES).
    Windows tile every frame row-major; border windows are smaller."""
    v = _v(cfg)
    m = v["spatial_merge_size"]
    ws = max(1, int(v["window_size"]) // m // v["patch_size"])
    order, lens, base = [], [], 0
    for t, h, w in _grid_list(grid_thw):
        gh, gw = h // m, w // m
        idx = torch.arange(t * gh * gw).view(t, gh, gw) + base
        for ti in range(t):
            for y in range(0, gh, ws):
                for x in range(0, gw, ws):
                    cell = idx[ti, y:y + ws, x:x + ws].reshape(-1)
                    order.append(cell)
                    lens.append(cell.numel() * m * m)
        base += t * gh * gw
    return torch.cat(order), lens


def frame_cu_seqlens(grid_thw, device) -> tuple[torch.Tensor, int]:
    """Attention is full inside each temporal frame of each image: one varlen segment per (image, frame)."""
    lens = [h * w for t, h, w in _grid_list(grid_thw) for _ in range(t)]
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32, device=device)
    cu[1:] = torch.tensor(lens, dtype=torch.int32, device=device).cumsum(0)
    return cu, max(lens) if lens else 0


def vision_forward(cfg: ModelConfig, vw: "VisionWeights", pixel_values: torch.Tensor, grid_thw) -> torch.Tensor:
    """pixel_values [N, C*Tp*P*P] (N = sum t*h*w) -> merged image embeddings [N / merge^2, hidden_size]."""
    x = vision_stem_in(cfg, vw.stem, pixel_values, grid_thw)
    cos, sin = vision_rotary(cfg, grid_thw, x.device)
    cu, mx = frame_cu_seqlens(grid_thw, x.device)
    if is_v3(cfg):
        # the merged embedding and the DeepStack levels travel as ONE [n, hidden * (1 + levels)] tensor (embed_width): every
        # consumer that slices rows (chunked prefill, SP slicing, ZeRO unit outputs) keeps working unchanged
        taps = list(_v(cfg).get("deepstack_visual_indexes") or ())
        feats: list = [None] * len(taps)
        for i, b in enumerate(vw.blocks):
            x = vision_block(cfg, b, x, cos, sin, cu, mx)
            if i in taps:
                feats[taps.index(i)] = deepstack_merge(cfg, vw.stem, taps.index(i), x)
        return torch.cat([vision_stem_out(cfg, vw.stem, x), *feats], dim=-1)
    if not is_v25(cfg):
        for b in vw.blocks:
            x = vision_block(cfg, b, x, cos, sin, cu, mx)
        return vision_stem_out(cfg, vw.stem, x)
    # Qwen2.5-VL: window-major patch order for the whole tower (full-attention blocks are order-invariant inside a frame because
    # frames stay contiguous), original order restored on the merged rows
    m2 = _v(cfg)["spatial_merge_size"] ** 2
    order, wlens = window_partition(cfg, grid_thw)
    order = order.to(x.device)
    gather = (order[:, None] * m2 + torch.arange(m2, device=x.device)[None]).reshape(-1)
    x, cos, sin = x[gather], cos[gather], sin[gather]
    cu_w = torch.zeros(len(wlens) + 1, dtype=torch.int32, device=x.device)
    cu_w[1:] = torch.tensor(wlens, dtype=torch.int32, device=x.device).cumsum(0)
    full = set(_v(cfg).get("fullatt_block_indexes", ()))
    for i, b in enumerate(vw.blocks):
        x = vision_block(cfg, b, x, cos, sin, cu, mx) if i in full else vision_block(cfg, b, x, cos, sin, cu_w, max(wlens))
    out = vision_stem_out(cfg, vw.stem, x)
    return out[torch.argsort(order)]


def num_image_tokens(cfg: ModelConfig, grid_thw) -> list[int]:
    m2 = _v(cfg)["spatial_merge_size"] ** 2
    return [t * h * w // m2 for t, h, w in _grid_list(grid_thw)]


def merge_image_embeds(cfg: ModelConfig, x: torch.Tensor, input_ids: torch.Tensor, image_embeds: torch.Tensor) -> torch.Tensor:
    """Replace the rows of x:[T,H] at image / video placeholder tokens by the tower's output rows, in order."""
    mask = (input_ids == cfg.image_token_id) | (input_ids == cfg.video_token_id)
    idx = mask.nonzero(as_tuple=True)[0]
    if idx.numel() != image_embeds.shape[0]:
        raise ValueError(f"{idx.numel()} image placeholder tokens but {image_embeds.shape[0]} image embedding rows")
    return x.index_copy(0, idx, image_embeds[:, : x.shape[-1]].to(x.dtype))


def deepstack_add(cfg: ModelConfig, x: torch.Tensor, input_ids: torch.Tensor, image_embeds: torch.Tensor | None, layer_idx: int) -> torch.Tensor:
    """Qwen3-VL DeepStack: after decoder layer ``layer_idx`` (< number of levels) add that level's visual features to the hidden
    states at the image / video placeholder positions.  No-op for other towers, text-only batches and deeper layers."""
    if image_embeds is None or layer_idx >= deepstack_levels(cfg) or image_embeds.shape[0] == 0:
        return x
    hsz = x.shape[-1]
    mask = (input_ids == cfg.image_token_id) | (input_ids == cfg.video_token_id)
    idx = mask.nonzero(as_tuple=True)[0]
    return x.index_add(0, idx, image_embeds[:, hsz * (layer_idx + 1): hsz * (layer_idx + 2)].to(x.dtype))


# --------------------------------------------------------------------------------------------
# M-RoPE
# --------------------------------------------------------------------------------------------
def mrope_positions(cfg: ModelConfig, input_ids, grid_thw, second_per_grid_ts=None) -> tuple[torch.Tensor, int]:
    """3-axis positions of ONE sequence -> ([3, T] int64, delta) where ``delta + T`` is the position of the next generated token.
    Text runs advance all three axes together; a vision block of (t, h/m, w/m) merged cells starts at the running position p and
    uses (p + frame, p + row, p + col); the run after it resumes at the largest index used + 1.  Images have t = 1.  For videos the
    frame index advances by 1 per temporal patch (Qwen2-VL) or by ``tokens_per_second * second_per_grid_t`` (Qwen2.5-VL, with
    ``second_per_grid_ts`` = seconds covered by one temporal patch of each video) — the layout of the reference's get_rope_index
    (verl/models/transformers/qwen2_vl.py:63-150)."""
    ids = input_ids.tolist() if isinstance(input_ids, torch.Tensor) else list(input_ids)
    v = _v(cfg)
    m = v["spatial_merge_size"]
    glist = _grid_list(grid_thw) if grid_thw is not None else []
    if is_v3(cfg) or is_glm4v(cfg):   # prompts separate the frames of a video by timestamp text: one placeholder run per frame
        glist = [(1, h, w) for t, h, w in glist for _ in range(t)]
    grids = iter(glist)
    secs = iter(second_per_grid_ts if second_per_grid_ts is not None else [])
    pos = torch.empty(3, len(ids), dtype=torch.long)
    i, p = 0, 0
    n = len(ids)
    vis = (cfg.image_token_id, cfg.video_token_id)
    while i < n:
        if ids[i] in vis:
            t, h, w = next(grids)
            gh, gw = h // m, w // m
            cnt = t * gh * gw
            if i + cnt > n or any(ids[j] not in vis for j in (i, i + cnt - 1)):
                raise ValueError("vision placeholder run does not match its grid")
            step = 1.0
            if ids[i] == cfg.video_token_id and is_v25(cfg):
                step = float(v.get("tokens_per_second", 2)) * float(next(secs, 1.0))
            frame = (torch.arange(t, dtype=torch.float32) * step).long().repeat_interleave(gh * gw)
            pos[0, i:i + cnt] = p + frame
            pos[1, i:i + cnt] = p + torch.arange(gh).repeat_interleave(gw).repeat(t)
            pos[2, i:i + cnt] = p + torch.arange(gw).repeat(gh * t)
            p = int(pos[:, i:i + cnt].max()) + 1
            i += cnt
        else:
            j = i
            while j < n and ids[j] not in vis:
                j += 1
            pos[:, i:j] = torch.arange(p, p + j - i)
            p += j - i
            i = j
    delta = (int(pos.max()) + 1 - n) if n else 0
    return pos, delta


def mrope_tables(cfg: ModelConfig, pos3: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-token rotary tables [T, head_dim / 2]: frequency i reads position stream s(i) given by ``mrope_section``.  Feeding them to
    the rope op with positions arange(T) reuses the 1-D kernel unchanged."""
    sec = torch.tensor(cfg.mrope_section, device=pos3.device)
    if getattr(cfg, "mrope_interleaved", False):
        # Qwen3-VL: frequencies cycle t, h, w, t, h, w, ... (h / w only inside their first 3 * section entries), the rest reads t
        i = torch.arange(int(sec.sum()), device=pos3.device)
        stream = torch.zeros_like(i)
        stream[(i % 3 == 1) & (i < 3 * sec[1])] = 1
        stream[(i % 3 == 2) & (i < 3 * sec[2])] = 2
    else:
        stream = torch.repeat_interleave(torch.arange(3, device=pos3.device), sec)        # [d/2] -> 0/1/2
    if stream.numel() < cos.shape[1]:               # partial rotary (GLM): the pass-through columns hold cos = 1 / sin = 0 anyway
        stream = F.pad(stream, (0, cos.shape[1] - stream.numel()))
    p = pos3.to(cos.device)[stream.to(cos.device)]                                         # [d/2, T]
    f = torch.arange(cos.shape[1], device=cos.device).unsqueeze(1)                        # [d/2, 1]
    return cos[p, f].t().contiguous(), sin[p, f].t().contiguous()


# --------------------------------------------------------------------------------------------
# image preprocessing (PIL / tensor -> patches)
# --------------------------------------------------------------------------------------------
def norm_stats(cfg: ModelConfig) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-channel (mean, std) of the family's image processor: CLIP statistics for Qwen2-VL / 2.5-VL, 0.5 / 0.5 for Qwen3-VL;
    ``vision.image_mean`` / ``vision.image_std`` override."""
    v = _v(cfg)
    default_mean, default_std = ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)) if v.get("arch") == "qwen3" else (CLIP_MEAN, CLIP_STD)   # GLM-4V: CLIP
    mean, std = v.get("image_mean") or default_mean, v.get("image_std") or default_std
    return torch.tensor(mean, dtype=torch.float32).view(3, 1, 1), torch.tensor(std, dtype=torch.float32).view(3, 1, 1)


def smart_resize(height: int, width: int, factor: int, min_pixels: int, max_pixels: int, frames: int = 1) -> tuple[int, int]:
    """Closest (h, w) that are multiples of ``factor``, keep the aspect ratio and fit the pixel budget.  ``frames`` > 1 is the
    GLM-4V rule: the budget counts frames x height x width voxels (a still image counts as one temporal patch of frames), and an
    image smaller than one merge cell is first scaled up to it."""
    if frames > 1 and (height < factor or width < factor):
        scale = max(factor / height, factor / width)
        height, width = int(height * scale), int(width * scale)
    if max(height, width) / max(1, min(height, width)) > 200:
        raise ValueError("aspect ratio above 200")
    h = max(factor, round(height / factor) * factor)
    w = max(factor, round(width / factor) * factor)
    if frames * h * w > max_pixels:
        beta = math.sqrt(frames * height * width / max_pixels)
        h = max(factor, math.floor(height / beta / factor) * factor)
        w = max(factor, math.floor(width / beta / factor) * factor)
    elif frames * h * w < min_pixels:
        beta = math.sqrt(min_pixels / (frames * height * width))
        h = math.ceil(height * beta / factor) * factor
        w = math.ceil(width * beta / factor) * factor
    return h, w


def patchify(cfg: ModelConfig, frames: torch.Tensor) -> tuple[torch.Tensor, tuple[int, int, int]]:
    """frames [F, C, H, W] (normalised; H, W multiples of patch*merge) -> ([t*gh*gw, C*Tp*P*P], (t, gh, gw)).  Frames are padded by
    repeating the last one up to a multiple of the temporal patch; rows are emitted merge-block by merge-block."""
    v = _v(cfg)
    p, m, tp = v["patch_size"], v["spatial_merge_size"], v["temporal_patch_size"]
    f, c, h, w = frames.shape
    if f % tp:
        frames = torch.cat([frames, frames[-1:].expand(tp - f % tp, c, h, w)], dim=0)
    gt, gh, gw = frames.shape[0] // tp, h // p, w // p
    x = frames.reshape(gt, tp, c, gh // m, m, p, gw // m, m, p)
    x = x.permute(0, 3, 6, 4, 7, 2, 1, 5, 8)                  # gt, gh/m, gw/m, m, m, c, tp, p, p
    return x.reshape(gt * gh * gw, c * tp * p * p).contiguous(), (gt, gh, gw)


def preprocess_image(cfg: ModelConfig, image, min_pixels: int | None = None, max_pixels: int | None = None):
    """One PIL image (or uint8 tensor / array [H, W, C]) -> (pixel_values [n, patch_dim] fp32, (t, gh, gw))."""
    from PIL import Image

    v = _v(cfg)
    factor = v["patch_size"] * v["spatial_merge_size"]
    if is_glm4v(cfg):                                  # GLM-4V processor defaults (budget in frames x pixels, see smart_resize)
        min_pixels = min_pixels or 112 * 112
        max_pixels = max_pixels or factor * factor * v["temporal_patch_size"] * 6144
    min_pixels = min_pixels or 4 * factor * factor
    max_pixels = max_pixels or 1280 * factor * factor
    if not isinstance(image, Image.Image):
        import numpy as np

        image = Image.fromarray(np.asarray(image, dtype="uint8"))
    image = image.convert("RGB")
    h, w = smart_resize(image.height, image.width, factor, min_pixels, max_pixels,
                        frames=v["temporal_patch_size"] if is_glm4v(cfg) else 1)
    image = image.resize((w, h), resample=Image.BICUBIC)
    import numpy as np

    x = torch.from_numpy(np.asarray(image).copy()).permute(2, 0, 1).float() / 255.0
    mean, std = norm_stats(cfg)
    x = (x - mean) / std
    return patchify(cfg, x.unsqueeze(0))


def preprocess_video(cfg: ModelConfig, frames, min_pixels: int | None = None, max_pixels: int | None = None):
    """A video given as a sequence of frames (PIL images / uint8 arrays, all resized to one size) -> (pixel_values, (t, gh, gw)).
    Frames are grouped ``temporal_patch_size`` at a time (the last one repeated to fill the final group)."""
    import numpy as np
    from PIL import Image

    v = _v(cfg)
    factor = v["patch_size"] * v["spatial_merge_size"]
    min_pixels = min_pixels or 4 * factor * factor
    max_pixels = max_pixels or 768 * factor * factor
    imgs = [f if isinstance(f, Image.Image) else Image.fromarray(np.asarray(f, dtype="uint8")) for f in frames]
    if not imgs:
        raise ValueError("empty video")
    h, w = smart_resize(imgs[0].height, imgs[0].width, factor, min_pixels, max_pixels)
    mean, std = norm_stats(cfg)
    xs = [(torch.from_numpy(np.asarray(im.convert("RGB").resize((w, h), resample=Image.BICUBIC)).copy()).permute(2, 0, 1).float() / 255.0 - mean) / std
          for im in imgs]
    return patchify(cfg, torch.stack(xs))


def expand_image_placeholders(cfg: ModelConfig, input_ids: list[int], grid_thw) -> list[int]:
    """A prompt tokenised with ONE placeholder token per image / video -> the placeholder repeated once per merged patch."""
    counts = iter(num_image_tokens(cfg, grid_thw))
    out: list[int] = []
    for t in input_ids:
        out.extend([t] * next(counts) if t in (cfg.image_token_id, cfg.video_token_id) else [t])
    return out


# --------------------------------------------------------------------------------------------
# HF checkpoint names
# --------------------------------------------------------------------------------------------
_BLOCK_MAP = {"ln1_w": "norm1.weight", "ln1_b": "norm1.bias", "wqkv": "attn.qkv.weight", "bqkv": "attn.qkv.bias",
              "wo": "attn.proj.weight", "bo": "attn.proj.bias", "ln2_w": "norm2.weight", "ln2_b": "norm2.bias",
              "wfc": "mlp.fc1.weight", "bfc": "mlp.fc1.bias", "wproj": "mlp.fc2.weight", "bproj": "mlp.fc2.bias"}
_STEM_MAP = {"merger_ln_w": "merger.ln_q.weight", "merger_ln_b": "merger.ln_q.bias", "merger_w1": "merger.mlp.0.weight",
             "merger_b1": "merger.mlp.0.bias", "merger_w2": "merger.mlp.2.weight", "merger_b2": "merger.mlp.2.bias"}


def _visual_prefix(hf_sd: dict) -> str:
    for p in ("model.visual.", "visual."):
        if any(k.startswith(p) for k in hf_sd):
            return p
    raise KeyError("no visual.* tensors in the checkpoint")


_BLOCK_MAP_25 = {"ln1_w": "norm1.weight", "wqkv": "attn.qkv.weight", "bqkv": "attn.qkv.bias", "wo": "attn.proj.weight",
                 "bo": "attn.proj.bias", "ln2_w": "norm2.weight", "wg": "mlp.gate_proj.weight", "bg": "mlp.gate_proj.bias",
                 "wu": "mlp.up_proj.weight", "bu": "mlp.up_proj.bias", "wproj": "mlp.down_proj.weight", "bproj": "mlp.down_proj.bias"}


_BLOCK_MAP_3 = {**{k: v for k, v in _BLOCK_MAP.items() if not k.startswith(("wfc", "bfc", "wproj", "bproj"))},
                "wfc": "mlp.linear_fc1.weight", "bfc": "mlp.linear_fc1.bias", "wproj": "mlp.linear_fc2.weight", "bproj": "mlp.linear_fc2.bias"}


def _stem_map_3(cfg: ModelConfig) -> dict:
    out = {"patch_embed_b": "patch_embed.proj.bias", "pos_embed": "pos_embed.weight", "merger_ln_w": "merger.norm.weight",
           "merger_ln_b": "merger.norm.bias", "merger_w1": "merger.linear_fc1.weight", "merger_b1": "merger.linear_fc1.bias",
           "merger_w2": "merger.linear_fc2.weight", "merger_b2": "merger.linear_fc2.bias"}
    for k in range(deepstack_levels(cfg)):
        pre = f"deepstack_merger_list.{k}."
        out.update({f"ds{k}_ln_w": pre + "norm.weight", f"ds{k}_ln_b": pre + "norm.bias", f"ds{k}_w1": pre + "linear_fc1.weight",
                    f"ds{k}_b1": pre + "linear_fc1.bias", f"ds{k}_w2": pre + "linear_fc2.weight", f"ds{k}_b2": pre + "linear_fc2.bias"})
    return out


_STEM_MAP_GLM = {"patch_embed_b": "patch_embed.proj.bias", "post_conv_ln_w": "post_conv_layernorm.weight",
                 "pos_embed": "embeddings.position_embedding.weight", "post_ln_w": "post_layernorm.weight",
                 "downsample_w": "downsample.weight", "downsample_b": "downsample.bias", "merger_proj": "merger.proj.weight",
                 "merger_ln_w": "merger.post_projection_norm.weight", "merger_ln_b": "merger.post_projection_norm.bias",
                 "merger_wg": "merger.gate_proj.weight", "merger_wu": "merger.up_proj.weight", "merger_wd": "merger.down_proj.weight"}


def _maps(cfg: ModelConfig) -> tuple[dict, dict]:
    if is_glm4v(cfg):
        drop = () if _v(cfg).get("attention_bias") else ("bqkv",)
        return _STEM_MAP_GLM, {k: v for k, v in _BLOCK_MAP_25.items() if k in ("ln1_w", "wqkv", "bqkv", "wo", "ln2_w", "wg", "wu", "wproj")
                               and k not in drop}
    if is_v3(cfg):
        return _stem_map_3(cfg), _BLOCK_MAP_3
    if is_v25(cfg):
        return {k: v for k, v in _STEM_MAP.items() if k != "merger_ln_b"}, _BLOCK_MAP_25
    return _STEM_MAP, _BLOCK_MAP


def convert_hf_vision_state_dict(cfg: ModelConfig, hf_sd: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """HF ``visual.*`` names -> ours.  The Conv3d patch embedding [E, C, Tp, P, P] flattens to a linear weight."""
    p = _visual_prefix(hf_sd)
    stem_map, block_map = _maps(cfg)
    out = {"vision.patch_embed": hf_sd[p + "patch_embed.proj.weight"].reshape(_v(cfg)["embed_dim"], -1)}
    for ours, theirs in stem_map.items():
        t = hf_sd[p + theirs]
        out[f"vision.{ours}"] = t.reshape(t.shape[0], -1) if ours == "downsample_w" else t       # Conv2d [O, E, m, m] -> linear
    for i in range(_v(cfg)["depth"]):
        for ours, theirs in block_map.items():
            out[f"vision.blocks.{i}.{ours}"] = hf_sd[f"{p}blocks.{i}.{theirs}"]
    return out


def export_hf_vision_state_dict(cfg: ModelConfig, sd: dict[str, torch.Tensor], prefix: str = "model.visual.") -> dict:
    v = _v(cfg)
    out = {prefix + "patch_embed.proj.weight": sd["vision.patch_embed"].reshape(
        v["embed_dim"], v["in_channels"], v["temporal_patch_size"], v["patch_size"], v["patch_size"])}
    stem_map, block_map = _maps(cfg)
    for ours, theirs in stem_map.items():
        t = sd[f"vision.{ours}"]
        m = v["spatial_merge_size"]
        out[prefix + theirs] = t.reshape(t.shape[0], v["embed_dim"], m, m) if ours == "downsample_w" else t
    for i in range(v["depth"]):
        for ours, theirs in block_map.items():
            out[f"{prefix}blocks.{i}.{theirs}"] = sd[f"vision.blocks.{i}.{ours}"]
    return out


# --------------------------------------------------------------------------------------------
# batch / wire helpers
# --------------------------------------------------------------------------------------------
def collate_multi_modal(mm_list) -> tuple[torch.Tensor | None, list | None]:
    """Per-sample ``multi_modal_inputs`` dicts (``pixel_values`` + ``image_grid_thw``; None for text rows) of a micro-batch, in row
    order -> (all patches concatenated, all grids) matching the order of the image placeholders in the packed token stream."""
    pix, grids = [], []
    for m in mm_list:
        if m is None:
            continue
        pix.append(torch.as_tensor(m["pixel_values"]))
        grids += _grid_list(m["image_grid_thw"] if "image_grid_thw" in m else m["grid_thw"])
    if not pix:
        return None, None
    return torch.cat(pix, dim=0), grids


def encode_mm_wire(mm: dict) -> dict:
    """JSON-safe form of one sample's image inputs for the HTTP generate route (fp16 patches, base64)."""
    import base64

    pv = torch.as_tensor(mm["pixel_values"]).to(torch.float16).contiguous()
    grid = mm["image_grid_thw"] if "image_grid_thw" in mm else mm["grid_thw"]
    out = {"pixel_values_b64": base64.b64encode(pv.numpy().tobytes()).decode("ascii"), "shape": list(pv.shape),
           "grid_thw": [list(g) for g in _grid_list(grid)]}
    if mm.get("second_per_grid_ts") is not None:
        out["second_per_grid_ts"] = [float(x) for x in mm["second_per_grid_ts"]]
    return out


def decode_mm_wire(d: dict) -> dict:
    import base64

    import numpy as np

    arr = np.frombuffer(base64.b64decode(d["pixel_values_b64"]), dtype=np.float16).reshape(d["shape"])
    out = {"pixel_values": torch.from_numpy(arr.copy()).float(), "grid_thw": [tuple(g) for g in d["grid_thw"]]}
    if d.get("second_per_grid_ts") is not None:
        out["second_per_grid_ts"] = list(d["second_per_grid_ts"])
    return out
