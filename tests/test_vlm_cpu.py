"""Qwen2-VL family on CPU: the functional ViT tower + M-RoPE text stack reproduce the HF implementation's logits on a tiny random
model; the image preprocessor matches the HF processor's patches; positions follow the (t, h, w) layout.
(reference: 3rdparty/verl/verl/models/transformers/qwen2_vl.py:36-150 get_rope_index, verl/utils/dataset/vision_utils.py)"""
import numpy as np
import pytest
import torch

from polyrl_b200.models import transformer as tf
from polyrl_b200.models import vision as V
from polyrl_b200.models.config import ModelConfig, get_model_config


def _hf_tiny():
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration

    torch.manual_seed(0)
    hc = Qwen2VLConfig(
        text_config=dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=512, rms_norm_eps=1e-6, tie_word_embeddings=False,
                         rope_parameters=dict(rope_type="default", rope_theta=10000.0, mrope_section=[4, 6, 6])),
        vision_config=dict(depth=2, embed_dim=64, hidden_size=128, num_heads=4, patch_size=4, spatial_merge_size=2,
                           temporal_patch_size=2, in_channels=3, mlp_ratio=2),
        image_token_id=500, video_token_id=501, vision_start_token_id=499, vision_end_token_id=498)
    hc._attn_implementation = "eager"
    model = Qwen2VLForConditionalGeneration(hc).eval().float()
    with torch.no_grad():                      # default init leaves biases / norms trivial; randomise so the test sees them
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return hc, model


def _ours(hc, model):
    cfg = ModelConfig.from_hf_config(hc.to_dict())
    sd = model.state_dict()
    w = tf.TransformerWeights(cfg, init=False)
    w.load_state_dict(tf.convert_hf_state_dict(cfg, sd))
    vw = V.VisionWeights(cfg, init=False)
    vw.load_state_dict(V.convert_hf_vision_state_dict(cfg, sd))
    return cfg, w, vw


def test_vision_tower_and_logits_match_hf():
    hc, model = _hf_tiny()
    cfg, w, vw = _ours(hc, model)
    assert cfg.vision["depth"] == 2 and cfg.mrope_section == (4, 6, 6) and cfg.qkv_bias
    g = torch.Generator().manual_seed(1)
    grids = [(1, 4, 6), (1, 8, 4)]                                   # patches (before the 2x2 merge)
    pix = torch.randn(sum(t * h * w_ for t, h, w_ in grids), V.patch_dim(cfg), generator=g)
    grid_t = torch.tensor(grids)
    with torch.no_grad():
        theirs = model.model.visual(pix, grid_thw=grid_t).pooler_output
        mine = V.vision_forward(cfg, vw, pix, grids)
    assert mine.shape == theirs.shape == (sum(V.num_image_tokens(cfg, grids)), cfg.hidden_size)
    assert (mine - theirs).abs().max() < 2e-4

    # one sequence: text, image 0, text, image 1, text
    n0, n1 = V.num_image_tokens(cfg, grids)
    ids = [2, 11, 12, 499] + [500] * n0 + [498, 13, 14, 15, 499] + [500] * n1 + [498, 16, 17]
    ids_t = torch.tensor([ids])
    mm_type = (ids_t == 500).int()
    with torch.no_grad():
        ref_logits = model(input_ids=ids_t, pixel_values=pix, image_grid_thw=grid_t, mm_token_type_ids=mm_type).logits[0]
    pos3, delta = V.mrope_positions(cfg, ids, grids)
    hf_pos, hf_delta = model.model.get_rope_index(ids_t, mm_type, image_grid_thw=grid_t)
    assert torch.equal(pos3, hf_pos[:, 0]) and delta == int(hf_delta[0, 0])
    t = len(ids)
    inp = tf.PackedInput(torch.tensor(ids), pos3[0].clone(), torch.tensor([0, t], dtype=torch.int32), t, mrope_pos=pos3,
                         image_embeds=mine)
    with torch.no_grad():
        h = tf.forward_hidden(cfg, w, inp)
        logits = tf.logits_forward(cfg, w.head, w.embed, h)
    assert (logits - ref_logits).abs().max() < 5e-4
    # a text-only sequence through the M-RoPE path equals the 1-D rope path
    tids = torch.arange(3, 40)
    p1 = torch.arange(tids.numel())
    cu = torch.tensor([0, tids.numel()], dtype=torch.int32)
    with torch.no_grad():
        a = tf.forward_hidden(cfg, w, tf.PackedInput(tids, p1, cu, tids.numel()))
        b = tf.forward_hidden(cfg, w, tf.PackedInput(tids, p1, cu, tids.numel(), mrope_pos=p1.expand(3, -1).contiguous()))
    assert (a - b).abs().max() < 1e-5


def test_mrope_positions_layout_and_errors():
    cfg = get_model_config("tiny-qwen2-vl")
    ids = [7, 8, 500, 500, 500, 500, 500, 500, 9]               # a 1 x 4 x 6 patch image = 2 x 3 merged cells
    pos, delta = V.mrope_positions(cfg, ids, [(1, 4, 6)])
    assert pos[:, :2].tolist() == [[0, 1]] * 3
    assert pos[0, 2:8].tolist() == [2] * 6
    assert pos[1, 2:8].tolist() == [2, 2, 2, 3, 3, 3] and pos[2, 2:8].tolist() == [2, 3, 4, 2, 3, 4]
    assert pos[:, 8].tolist() == [5, 5, 5] and delta == 6 - 9            # next position 6 for a 9-token prompt
    with pytest.raises(ValueError):
        V.mrope_positions(cfg, [7, 500, 500, 9], [(1, 4, 6)])
    assert V.expand_image_placeholders(cfg, [7, 500, 9], [(1, 4, 6)]) == [7] + [500] * 6 + [9]
    x = torch.zeros(4, cfg.hidden_size)
    with pytest.raises(ValueError):
        V.merge_image_embeds(cfg, x, torch.tensor([1, 500, 500, 2]), torch.ones(3, cfg.hidden_size))
    y = V.merge_image_embeds(cfg, x, torch.tensor([1, 500, 500, 2]), torch.ones(2, cfg.hidden_size))
    assert y.sum() == 2 * cfg.hidden_size and x.sum() == 0


def test_preprocess_matches_hf_processor():
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor

    cfg = get_model_config("qwen2-vl-7b")
    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 255, size=(75, 131, 3), dtype=np.uint8))
    mine, grid = V.preprocess_image(cfg, img, min_pixels=56 * 56, max_pixels=28 * 28 * 64)
    proc = Qwen2VLImageProcessor(min_pixels=56 * 56, max_pixels=28 * 28 * 64)
    out = proc(images=[img], return_tensors="pt")
    assert tuple(out["image_grid_thw"][0].tolist()) == grid
    assert mine.shape == out["pixel_values"].shape
    assert (mine - out["pixel_values"]).abs().max() < 0.05         # resize kernels differ in the last bit of uint8 rounding
    assert (mine - out["pixel_values"]).abs().mean() < 5e-3
    # budget clamps: a huge image shrinks under max_pixels, a tiny one grows above min_pixels
    assert np.prod(V.smart_resize(4000, 3000, 28, 56 * 56, 28 * 28 * 1280)) <= 28 * 28 * 1280
    h, w = V.smart_resize(20, 30, 28, 56 * 56, 28 * 28 * 1280)
    assert h * w >= 56 * 56 and h % 28 == 0 and w % 28 == 0


def test_vision_state_dict_round_trip_and_param_count():
    cfg = get_model_config("tiny-qwen2-vl")
    vw = V.VisionWeights(cfg)
    sd = vw.state_dict()
    hf = V.export_hf_vision_state_dict(cfg, sd)
    assert hf["model.visual.patch_embed.proj.weight"].shape == (64, 3, 2, 4, 4)
    back = V.convert_hf_vision_state_dict(cfg, hf)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    assert V.vision_num_params(cfg) == sum(t.numel() for t in sd.values())
    assert cfg.num_params() == V.vision_num_params(cfg) + get_model_config("tiny-qwen2-vl", vision=None, mrope_section=()).num_params()


@pytest.mark.parametrize("family", ["qwen2_vl", "qwen3_vl", "glm4v"])
def test_engine_generates_like_hf_with_images(family):
    """Rollout engine on image prompts: greedy continuation equals HF generate (prefill with tower embeddings + M-RoPE, decode
    with the per-request rotary offset), also when the prompt is prefilled in chunks and when n samples share the prompt.
    Qwen3-VL adds DeepStack: the per-level features are added after the first decoder layers of every prefill chunk."""
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams

    hc, model = {"qwen2_vl": _hf_tiny, "qwen3_vl": _hf_tiny_q3, "glm4v": _hf_tiny_glm4v}[family]()
    cfg, w, vw = _ours(hc, model)
    g = torch.Generator().manual_seed(3)
    grids_a, grids_b = [(1, 4, 6)], [(1, 4, 4), (1, 8, 4)]
    pix_a = torch.randn(24, V.patch_dim(cfg), generator=g)
    pix_b = torch.randn(16 + 32, V.patch_dim(cfg), generator=g)
    pa = [2, 21, 499, 500, 498, 22, 23]                                   # ONE placeholder per image: the engine expands them
    pb = [2, 499, 500, 498, 31, 32, 33, 34, 499, 500, 498, 35]
    pt = [2, 40, 41, 42, 43, 44]                                          # text only, same batch
    new = 12

    def hf_greedy(ids, pix, grids):
        ids = V.expand_image_placeholders(cfg, ids, grids) if grids else ids
        ids_t = torch.tensor([ids])
        kw = {}
        if grids:
            kw = dict(pixel_values=pix, image_grid_thw=torch.tensor(grids), mm_token_type_ids=(ids_t == 500).int())
        with torch.no_grad():
            out = model.generate(input_ids=ids_t, attention_mask=torch.ones_like(ids_t), max_new_tokens=new, do_sample=False,
                                 eos_token_id=None, pad_token_id=0, **kw)
        return out[0, len(ids):].tolist()

    want = [hf_greedy(pa, pix_a, grids_a), hf_greedy(pb, pix_b, grids_b), hf_greedy(pt, None, None)]
    mm = [dict(pixel_values=pix_a, grid_thw=grids_a), dict(pixel_values=pix_b, grid_thw=grids_b), None]
    for chunk in (16384, 8):
        eng = RolloutEngine(cfg, "cpu", max_num_seqs=8, max_model_len=128, kv_cache_tokens=2048, max_prefill_tokens=chunk,
                            enable_prefix_cache=True)
        eng.set_weights(w.embed, w.layers, w.head)
        eng.set_vision_weights(vw)
        sp = SamplingParams(temperature=0.0, max_new_tokens=new, n=2, ignore_eos=True)
        groups = eng.generate([pa, pb, pt], sp, mm=mm)
        for grp, ref_out in zip(groups, want):
            for r in grp:
                assert r.output_ids == ref_out, (chunk, r.output_ids, ref_out)
        assert groups[0][0].mrope_delta == 3 - 6 and groups[2][0].mrope_delta == 0
    with pytest.raises(ValueError):
        eng.add_request([2, 499, 500, 500, 498], sp, mm=dict(pixel_values=pix_a, grid_thw=grids_a))     # 2 tokens for a 6-cell image
    text_cfg = get_model_config("tiny-llama")
    with pytest.raises(ValueError):
        RolloutEngine(text_cfg, "cpu", max_num_seqs=2, max_model_len=64, kv_cache_tokens=256).add_request([1, 2], sp, mm=mm[0])


def _vlm_batch(cfg, seed=5):
    """Two packed sequences: one with two images, one text only."""
    g = torch.Generator().manual_seed(seed)
    grids = [(1, 4, 4), (1, 4, 8)]
    pix = torch.randn(16 + 32, V.patch_dim(cfg), generator=g)
    a = V.expand_image_placeholders(cfg, [2, 11, 499, 500, 498, 12, 13, 499, 500, 498, 14, 15, 16], grids)
    b = [2, 31, 32, 33, 34, 35, 36]
    pa, _ = V.mrope_positions(cfg, a, grids)
    pb = torch.arange(len(b)).expand(3, -1)
    ids = torch.tensor(a + b)
    pos3 = torch.cat([pa, pb], dim=1).contiguous()
    cu = torch.tensor([0, len(a), len(a) + len(b)], dtype=torch.int32)
    inp = tf.PackedInput(ids, pos3[0].clone(), cu, max(len(a), len(b)), mrope_pos=pos3, pixel_values=pix, grid_thw=grids)
    labels = torch.randint(0, 400, (ids.numel(),), generator=g)
    return inp, labels


@pytest.mark.parametrize("preset", ["tiny-qwen2-vl", "tiny-qwen3-vl", "tiny-glm4v"])
def test_zero_engine_trains_the_tower_like_plain_autograd(preset):
    """ZeRO engine with a ``vision`` unit: log-probs and the gradients of tower + text parameters equal plain autograd over the
    functional model, in recompute and saved-activation modes; a frozen tower keeps its weights; HF export reloads in HF.
    (Qwen3-VL: the DeepStack levels reach the decoder outside the layer units and still carry gradients into the tower.)"""
    from polyrl_b200.ops import reference as ref
    from polyrl_b200.parallel.zero3 import Zero3Engine

    cfg = get_model_config(preset)
    inp, labels = _vlm_batch(cfg)
    eng = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=3, max_grad_norm=1e9)
    full = {n: t.clone() for n, t in eng.named_full_tensors()}
    assert any(n.startswith("vision.blocks.1.") for n in full) and eng.num_params() == cfg.num_params()
    # oracle: same weights, plain autograd
    w = tf.TransformerWeights(cfg, init=False)
    w.load_state_dict(full)
    vw = V.VisionWeights(cfg, init=False)
    vw.load_state_dict(full)
    w.requires_grad_(True), vw.requires_grad_(True)
    img = V.vision_forward(cfg, vw, inp.pixel_values, inp.grid_thw)
    o_inp = tf.PackedInput(inp.input_ids, inp.position_ids, inp.cu_seqlens, inp.max_seqlen, mrope_pos=inp.mrope_pos, image_embeds=img)
    h = tf.forward_hidden(cfg, w, o_inp)
    lp_o, _ = ref.linear_cross_entropy(h, w.head["lm_head"], labels)
    (-lp_o.mean()).backward()
    g_o = {n: t.grad.clone() for n, t in list(w.named_tensors()) + list(vw.named_tensors())}
    for mode in ("full", "none"):
        e = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=3, max_grad_norm=1e9)
        e.act_ckpt = mode
        lp, _ = e.forward_logprob_entropy(inp, labels, 1.0)
        assert (lp - lp_o.detach()).abs().max() < 1e-5
        (-lp.mean()).backward()
        for u in e.units:
            u.wait_grads()
            for (n, shp), numel, so in zip(u.specs, u.numels, u.shard_offsets):
                name = n if u in (e.embed, e.head) else f"{u.name}.{n}"
                got = u.grad_acc[so:so + numel].view(shp)
                assert (got - g_o[name]).abs().max() < 1e-5 * max(1.0, g_o[name].abs().max().item()), (mode, name)
        assert g_o["vision.patch_embed"].abs().max() > 0 and g_o["vision.blocks.0.wqkv"].abs().max() > 0
        st = e.step()
        assert not st["skipped"]
    # text-only batch on the same engine: the tower is not touched, no image inputs needed
    tids = torch.arange(3, 30)
    tin = tf.PackedInput(tids, torch.arange(27), torch.tensor([0, 27], dtype=torch.int32), 27)
    lp_t, _ = e.forward_logprob_entropy(tin, torch.arange(27), 1.0)
    assert torch.isfinite(lp_t).all()
    # frozen tower
    fz = Zero3Engine(cfg, "cpu", lr=1e-2, weight_decay=0.0, seed=3, freeze_vision_tower=True)
    assert fz.vision.master is None and fz.vision not in fz.opt_units
    lp, _ = fz.forward_logprob_entropy(inp, labels, 1.0)
    (-lp.mean()).backward()
    fz.step()
    after = dict(fz.named_full_tensors())
    assert all(torch.equal(after[n], full[n]) for n in full if n.startswith("vision."))
    assert not torch.equal(after["layers.0.wqkv"], full["layers.0.wqkv"])


@pytest.mark.parametrize("preset", ["tiny-qwen2-vl", "tiny-glm4v"])
def test_vlm_hf_checkpoint_round_trip(tmp_path, preset):
    """save_hf of a VLM engine -> a directory HF's Qwen2VLForConditionalGeneration / Glm4vForConditionalGeneration loads, with
    identical logits; and back into a fresh engine / state dict through load_into_engine / load_state_dict."""
    from transformers import Glm4vForConditionalGeneration, Qwen2VLForConditionalGeneration

    from polyrl_b200.models import hf_io
    from polyrl_b200.parallel.zero3 import Zero3Engine

    cfg = get_model_config(preset)
    eng = Zero3Engine(cfg, "cpu", seed=11)
    hf_io.save_hf(eng, cfg, tmp_path / "ck", dtype=torch.float32)
    hf_cls = Glm4vForConditionalGeneration if preset == "tiny-glm4v" else Qwen2VLForConditionalGeneration
    model = hf_cls.from_pretrained(tmp_path / "ck", dtype=torch.float32).eval()
    inp, _ = _vlm_batch(cfg)
    n0 = int(inp.cu_seqlens[1])
    ids_t = inp.input_ids[:n0].unsqueeze(0)
    with torch.no_grad():
        ref_logits = model(input_ids=ids_t, pixel_values=inp.pixel_values, image_grid_thw=torch.tensor(inp.grid_thw),
                           mm_token_type_ids=(ids_t == cfg.image_token_id).int()).logits[0]
        mine = eng.forward_logits(inp)[:n0]
    assert (mine - ref_logits).abs().max() < 5e-4
    cfg2 = get_model_config(str(tmp_path / "ck"))
    assert cfg2.vision == cfg.vision and cfg2.mrope_section == cfg.mrope_section and cfg2.image_token_id == cfg.image_token_id
    e2 = Zero3Engine(cfg2, "cpu", seed=99)
    assert hf_io.load_into_engine(e2, tmp_path / "ck") > 0
    a, b = dict(eng.named_full_tensors()), dict(e2.named_full_tensors())
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    sd = hf_io.load_state_dict(cfg2, tmp_path / "ck")
    assert all(torch.equal(sd[k], a[k]) for k in a)


def _png_bytes(rng, h, w):
    import io

    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(rng.integers(0, 255, size=(h, w, 3), dtype=np.uint8)).save(buf, format="PNG")
    return buf.getvalue()


def _vlm_rows(tmp_path, n=4):
    import json

    rng = np.random.default_rng(0)
    rows = []
    for i in range(n):
        if i % 2 == 0:
            p = tmp_path / f"img{i}.png"
            p.write_bytes(_png_bytes(rng, 16 + 8 * (i % 3), 24))
            rows.append({"prompt": [{"role": "user", "content": f"<image>what is {i}?"}], "images": [{"path": str(p)}],
                         "data_source": "geo3k", "reward_model": {"ground_truth": str(i)}})
        else:
            rows.append({"prompt": [{"role": "user", "content": f"say {i}"}], "data_source": "geo3k",
                         "reward_model": {"ground_truth": str(i)}})
    f = tmp_path / "vl.jsonl"
    f.write_text("\n".join(json.dumps(r) for r in rows))
    return f


def test_dataset_builds_multi_modal_items_and_actor_scores_them(tmp_path):
    """RLHFDataset with a VLM config: placeholders expanded, 3-axis positions, patches attached; the worker generates with its
    colocated engine and the actor's log-probs of the generated tokens agree with the engine's (same model, same images)."""
    from polyrl_b200.data.rl_dataset import ByteTokenizer, RLHFDataset
    from polyrl_b200.protocol import DataProto, collate_fn

    cfg = get_model_config("tiny-qwen2-vl")
    ds = RLHFDataset(_vlm_rows(tmp_path), ByteTokenizer(), max_prompt_length=64, model_cfg=cfg, min_pixels=64, max_pixels=8 * 8 * 16)
    it0, it1 = ds[0], ds[1]
    g = it0["multi_modal_inputs"]["image_grid_thw"]
    n_img = int(g[0].prod()) // 4
    assert it0["raw_prompt_ids"].count(cfg.image_token_id) == n_img and it0["multi_modal_inputs"]["pixel_values"].shape == (int(g[0].prod()), V.patch_dim(cfg))
    assert it0["position_ids"].shape == (3, 64) and it1["position_ids"].shape == (3, 64) and it1["multi_modal_inputs"] is None
    k = it0["raw_prompt_ids"].index(cfg.vision_start_token_id)
    assert it0["raw_prompt_ids"][k + 1] == cfg.image_token_id and it0["raw_prompt_ids"][k + 1 + n_img] == cfg.vision_end_token_id
    want, _ = V.mrope_positions(cfg, it0["raw_prompt_ids"], g)
    assert torch.equal(it0["position_ids"][:, 64 - len(it0["raw_prompt_ids"]):], want)
    with pytest.raises(ValueError):
        RLHFDataset(_vlm_rows(tmp_path), ByteTokenizer(), max_prompt_length=8, model_cfg=cfg, min_pixels=64, max_pixels=8 * 8 * 16)[0]
    batch = collate_fn([ds[i] for i in range(4)])
    tens = {k: v for k, v in batch.items() if torch.is_tensor(v)}
    dp = DataProto.from_dict(tens, {k: v for k, v in batch.items() if not torch.is_tensor(v)})
    assert dp.batch["position_ids"].shape == (4, 3, 64)

    from polyrl_b200.config.schema import load_config
    from polyrl_b200.workers.hybrid_worker import ActorRolloutRefWorker

    c = load_config(None, ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-qwen2-vl", "data.max_prompt_length=64",
                           "data.max_response_length=10", "actor_rollout_ref.rollout.max_num_seqs=16", "actor_rollout_ref.rollout.n=2",
                           "actor_rollout_ref.rollout.cuda_graph=False", "actor_rollout_ref.actor.ppo_mini_batch_size=8",
                           "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=3"])
    wk = ActorRolloutRefWorker(c, role="actor_rollout", device="cpu")
    wk.init_model()
    assert wk.rollout.vision_w is not None
    dp.meta_info.update(temperature=1.0, ignore_eos=True)
    out = wk.generate_sequences(dp)
    assert out.batch["position_ids"].shape == (8, 3, 74) and out.batch["responses"].shape == (8, 10)
    assert len(out.non_tensor_batch["multi_modal_inputs"]) == 8
    # response positions continue after the largest prompt position on all three axes
    p3 = out.batch["position_ids"]
    assert torch.equal(p3[:, 0, 64:], p3[:, 1, 64:]) and torch.equal(p3[:, 0, 64], p3[:, :, :64].amax(dim=(1, 2)) + 1)
    out.meta_info.update(temperature=1.0, micro_batch_size=3)
    lp = wk.compute_log_prob(out).batch["old_log_probs"]
    assert (lp - out.batch["rollout_log_probs"]).abs().max() < 2e-4          # engine (prefill+decode) == trainer forward
    # a policy update on the image batch moves tower and text weights, and the refreshed engine sees the new tower
    before = {n: t.clone() for n, t in wk.actor_engine.named_full_tensors()}
    out = out.union(wk.compute_log_prob(out))
    out.batch["advantages"] = torch.randn(8, 10, generator=torch.Generator().manual_seed(0)) * out.batch["response_mask"]
    out.meta_info.update({"is_opt_step": True})
    m = wk.update_actor_stream(out)
    assert "actor/pg_loss" in m.meta_info["metrics"]
    after = dict(wk.actor_engine.named_full_tensors())
    assert not torch.equal(after["vision.patch_embed"], before["vision.patch_embed"])
    assert not torch.equal(after["layers.0.wqkv"], before["layers.0.wqkv"])
    wk.update_weight_remote()
    assert wk.rollout.vision_w.stem["patch_embed"].data_ptr() == wk.actor_engine.vision.views(wk.actor_engine.vision.shard)["patch_embed"].data_ptr()


def test_image_prompts_over_http_through_the_manager(tmp_path):
    """The remote path: client payload with ``image_data`` -> C++ manager (forwards the body untouched) -> engine server -> tokens
    identical to a direct engine call on the same (fp16-rounded) patches; the client rebuilds 3-axis position ids."""
    import json

    from polyrl_b200.data.rl_dataset import ByteTokenizer, RLHFDataset
    from polyrl_b200.manager.build import build_manager
    from polyrl_b200.protocol import DataProto, collate_fn
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams
    from polyrl_b200.rollout.remote_client import build_generate_payload, postprocess_samples
    from polyrl_b200.rollout.server import EngineServer
    from tests.test_manager import Manager, _http

    cfg = get_model_config("tiny-qwen2-vl")
    w, vw = tf.TransformerWeights(cfg, seed=3), V.VisionWeights(cfg, seed=4)

    def engine():
        e = RolloutEngine(cfg, "cpu", max_num_seqs=16, max_model_len=128, kv_cache_tokens=16 * 128)
        e.set_weights(w.embed, w.layers, w.head)
        e.set_vision_weights(vw)
        return e

    ds = RLHFDataset(_vlm_rows(tmp_path), ByteTokenizer(), max_prompt_length=64, model_cfg=cfg, min_pixels=64, max_pixels=8 * 8 * 16)
    batch = collate_fn([ds[i] for i in range(4)])
    dp = DataProto.from_dict({k: v for k, v in batch.items() if torch.is_tensor(v)}, {k: v for k, v in batch.items() if not torch.is_tensor(v)})
    reqs = build_generate_payload(dp, 2, {"max_new_tokens": 6, "temperature": 0.0, "ignore_eos": True})
    assert "image_data" in reqs[0][1] and "image_data" not in reqs[1][1]
    wire = V.decode_mm_wire(json.loads(json.dumps(reqs[0][1]["image_data"])))
    assert wire["grid_thw"] == [tuple(dp.non_tensor_batch["multi_modal_inputs"][0]["image_grid_thw"][0].tolist())]
    assert (wire["pixel_values"] - dp.non_tensor_batch["multi_modal_inputs"][0]["pixel_values"]).abs().max() < 2e-3
    m = Manager(build_manager())
    srv = EngineServer(engine(), stream_interval=4).start()
    try:
        srv.register_with_manager(f"127.0.0.1:{m.port}", local=True)
        code, body = _http("POST", m.url + "/batch_generate_requests", reqs, timeout=120)
        lines = [json.loads(l) for l in body.strip().split("\n")]
        got = {l["id"]: l for l in lines[1:]}
        ref = engine()
        sp = SamplingParams(max_new_tokens=6, temperature=0.0, ignore_eos=True)
        results = []
        for i in range(4):
            assert got[i + 1]["success"], got[i + 1]
            mm_i = V.decode_mm_wire(reqs[i][1]["image_data"]) if "image_data" in reqs[i][1] else None
            want = ref.generate([dp.non_tensor_batch["raw_prompt_ids"][i]], sp, mm=[mm_i])[0][0].output_ids
            for sub in got[i + 1]["data"]:
                assert [t[1] for t in sub["meta_info"]["output_token_logprobs"]] == want
            results.append((i, got[i + 1]["data"]))
        out = postprocess_samples(dp, results, 2, 6, cfg.pad_token_id, cfg.eos_token_id)
        assert out.batch["position_ids"].shape == (8, 3, 70) and len(out.non_tensor_batch["multi_modal_inputs"]) == 8
    finally:
        srv.stop()
        m.close()


def test_qwen2_5_vl_window_tower_matches_hf(tmp_path):
    """Qwen2.5-VL: RMSNorm / SwiGLU tower with windowed attention (full attention only in ``fullatt_block_indexes``) — tower
    output, LM logits with images and the exported checkpoint all equal HF."""
    from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration

    from polyrl_b200.models import hf_io

    torch.manual_seed(0)
    hc = Qwen2_5_VLConfig(
        text_config=dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=512, rms_norm_eps=1e-6, tie_word_embeddings=False, pad_token_id=0, eos_token_id=1, bos_token_id=2,
                         rope_parameters=dict(rope_type="default", rope_theta=10000.0, mrope_section=[4, 6, 6])),
        vision_config=dict(depth=3, hidden_size=64, intermediate_size=96, out_hidden_size=128, num_heads=4, patch_size=4,
                           spatial_merge_size=2, temporal_patch_size=2, in_channels=3, window_size=16, fullatt_block_indexes=[1]),
        image_token_id=500, video_token_id=501, vision_start_token_id=499, vision_end_token_id=498)
    hc._attn_implementation = "eager"
    model = Qwen2_5_VLForConditionalGeneration(hc).eval().float()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    cfg = ModelConfig.from_hf_config(hc.to_dict())
    assert cfg.vision["arch"] == "qwen2_5" and cfg.vision["window_size"] == 16 and cfg.vision["fullatt_block_indexes"] == [1]
    sd = model.state_dict()
    w = tf.TransformerWeights(cfg, init=False)
    w.load_state_dict(tf.convert_hf_state_dict(cfg, sd))
    vw = V.VisionWeights(cfg, init=False)
    vw.load_state_dict(V.convert_hf_vision_state_dict(cfg, sd))
    g = torch.Generator().manual_seed(1)
    grids = [(1, 6, 10), (1, 8, 4)]               # 3x5 and 4x2 merged cells, window = 2x2 cells: ragged border windows
    order, wlens = V.window_partition(cfg, grids)
    assert sorted(order.tolist()) == list(range(15 + 8)) and sum(wlens) == 60 + 32 and max(wlens) == 16 and min(wlens) == 4
    pix = torch.randn(60 + 32, V.patch_dim(cfg), generator=g)
    with torch.no_grad():
        theirs = model.model.visual(pix, grid_thw=torch.tensor(grids)).pooler_output
        mine = V.vision_forward(cfg, vw, pix, grids)
    assert (mine - theirs).abs().max() < 2e-4
    n0, n1 = V.num_image_tokens(cfg, grids)
    ids = [2, 11, 499] + [500] * n0 + [498, 13, 499] + [500] * n1 + [498, 16, 17]
    ids_t = torch.tensor([ids])
    # positions are passed explicitly: an image's temporal index is its start position (second_per_grid_t = 0, as in the
    # reference's get_rope_index, verl/models/transformers/qwen2_vl.py:63-150); this transformers release scales it by
    # tokens_per_second when no second_per_grid_ts is given
    pos3, _ = V.mrope_positions(cfg, ids, grids)
    hf_kw = dict(input_ids=ids_t, pixel_values=pix, image_grid_thw=torch.tensor(grids), mm_token_type_ids=(ids_t == 500).int(),
                 position_ids=pos3[:, None, :])
    with torch.no_grad():
        ref_logits = model(**hf_kw).logits[0]
    inp = tf.PackedInput(torch.tensor(ids), pos3[0].clone(), torch.tensor([0, len(ids)], dtype=torch.int32), len(ids), mrope_pos=pos3,
                         image_embeds=mine)
    with torch.no_grad():
        logits = tf.logits_forward(cfg, w.head, w.embed, tf.forward_hidden(cfg, w, inp))
    assert (logits - ref_logits).abs().max() < 5e-4
    # export -> HF loads it
    full = {**w.state_dict(), **vw.state_dict()}
    hf_io.save_hf(full, cfg, tmp_path / "ck25", dtype=torch.float32)
    m2 = Qwen2_5_VLForConditionalGeneration.from_pretrained(tmp_path / "ck25", dtype=torch.float32).eval()
    with torch.no_grad():
        again = m2(**hf_kw).logits[0]
    assert (again - ref_logits).abs().max() < 1e-4
    assert get_model_config(str(tmp_path / "ck25")).vision == cfg.vision


def test_video_inputs_positions_tower_and_dataset(tmp_path):
    """Videos: temporal patches become extra frames of the tower grid (t > 1), M-RoPE advances the temporal axis per frame (scaled by
    tokens_per_second * seconds-per-patch for Qwen2.5-VL), the dataset turns ``<video>`` + frame lists into placeholders/patches and
    the engine generates from them exactly like a trainer-side forward."""
    import json

    from polyrl_b200.data.rl_dataset import ByteTokenizer, RLHFDataset
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams

    cfg = get_model_config("tiny-qwen2-vl")
    ids = [7, 501, 501, 501, 501, 501, 501, 501, 501, 9]                        # a 2-frame video of 2 x 2 merged cells
    pos, delta = V.mrope_positions(cfg, ids, [(2, 4, 4)])
    assert pos[0, 1:9].tolist() == [1] * 4 + [2] * 4 and pos[1, 1:9].tolist() == [1, 1, 2, 2] * 2 and pos[2, 1:9].tolist() == [1, 2] * 4
    assert pos[:, 9].tolist() == [3, 3, 3] and delta == 4 - 10
    c25 = get_model_config("tiny-qwen2-vl", vision={**cfg.vision, "arch": "qwen2_5", "intermediate_size": 96, "window_size": 16,
                                                     "fullatt_block_indexes": [1], "tokens_per_second": 2})
    p25, _ = V.mrope_positions(c25, ids, [(2, 4, 4)], second_per_grid_ts=[1.5])
    assert p25[0, 1:9].tolist() == [1] * 4 + [4] * 4 and p25[:, 9].tolist() == [5, 5, 5]          # frame step = 2 tokens/s * 1.5 s
    # tower with t = 2 equals HF's
    hc, model = _hf_tiny()
    hcfg, w, vw = _ours(hc, model)
    pix = torch.randn(32, V.patch_dim(hcfg), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert (V.vision_forward(hcfg, vw, pix, [(2, 4, 4)]) - model.model.visual(pix, grid_thw=torch.tensor([[2, 4, 4]])).pooler_output).abs().max() < 2e-4
    # dataset: 3 frames -> 2 temporal patches (last frame repeated)
    rng = np.random.default_rng(0)
    frames = []
    for k in range(3):
        p = tmp_path / f"f{k}.png"
        p.write_bytes(_png_bytes(rng, 16, 16))
        frames.append({"path": str(p)})
    row = {"prompt": [{"role": "user", "content": "<video>what moves?"}], "videos": [{"frames": frames, "fps": 4.0}],
           "data_source": "x", "reward_model": {"ground_truth": "1"}}
    f = tmp_path / "v.jsonl"
    f.write_text(json.dumps(row))
    ds = RLHFDataset(f, ByteTokenizer(), max_prompt_length=64, model_cfg=cfg, min_pixels=64, max_pixels=256)
    it = ds[0]
    g = it["multi_modal_inputs"]["image_grid_thw"][0].tolist()
    assert g[0] == 2 and it["multi_modal_inputs"]["second_per_grid_ts"] == [0.5]
    n_tok = g[0] * g[1] * g[2] // 4
    assert it["raw_prompt_ids"].count(cfg.video_token_id) == n_tok and it["raw_prompt_ids"].count(cfg.image_token_id) == 0
    assert it["multi_modal_inputs"]["pixel_values"].shape == (g[0] * g[1] * g[2], V.patch_dim(cfg))
    # engine: greedy tokens from the video prompt == argmax of a full forward with the same tower rows
    eng = RolloutEngine(hcfg, "cpu", max_num_seqs=4, max_model_len=128, kv_cache_tokens=512)
    eng.set_weights(w.embed, w.layers, w.head)
    eng.set_vision_weights(vw)
    pv = it["multi_modal_inputs"]["pixel_values"]
    prompt = [2, 499, 501, 498, 31, 32]
    r = eng.generate([prompt], SamplingParams(max_new_tokens=4, temperature=0.0, ignore_eos=True), mm=[dict(pixel_values=pv, grid_thw=[tuple(g)])])[0][0]
    full = r.prompt_ids + r.output_ids
    assert r.prompt_ids.count(501) == n_tok
    pos3, _ = V.mrope_positions(hcfg, r.prompt_ids, [tuple(g)])
    nxt = int(pos3.max()) + 1
    pos3 = torch.cat([pos3, torch.arange(nxt, nxt + 4).expand(3, -1)], dim=1)
    inp = tf.PackedInput(torch.tensor(full), pos3[0].clone(), torch.tensor([0, len(full)], dtype=torch.int32), len(full), mrope_pos=pos3,
                         image_embeds=V.vision_forward(hcfg, vw, pv, [tuple(g)]))
    with torch.no_grad():
        logits = tf.logits_forward(hcfg, w.head, w.embed, tf.forward_hidden(hcfg, w, inp))
    k = len(r.prompt_ids) - 1
    assert logits[k:k + 4].argmax(-1).tolist() == r.output_ids


def test_model_merger_handles_vlm_checkpoints(tmp_path):
    """Sharded VLM checkpoint -> merged state (tower included) -> a HF directory that Qwen2VLForConditionalGeneration loads."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    from transformers import Qwen2VLForConditionalGeneration

    from polyrl_b200.parallel.zero3 import Zero3Engine
    from polyrl_b200.utils.checkpoint import ShardedCheckpointManager
    from polyrl_b200.utils.model_merger import merge

    cfg = get_model_config("tiny-qwen2-vl")
    eng = Zero3Engine(cfg, "cpu", seed=9)
    ShardedCheckpointManager(eng, cfg).save(str(tmp_path / "actor"), global_step=1)
    mcfg, sd = merge(str(tmp_path / "actor"))
    want = dict(eng.named_full_tensors())
    assert mcfg.vision == cfg.vision and set(sd) == set(want) and all(torch.equal(sd[k].float(), want[k].float()) for k in sd)
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "polyrl_b200.utils.model_merger", "--ckpt", str(tmp_path / "actor"), "--out", str(tmp_path / "hf"),
                        "--hf-dir"], capture_output=True, text=True, env={**os.environ, "PYTHONPATH": str(root)})
    assert r.returncode == 0, r.stderr[-1500:]
    m = Qwen2VLForConditionalGeneration.from_pretrained(tmp_path / "hf", dtype=torch.float32)
    got = m.state_dict()["model.visual.blocks.1.attn.qkv.weight"]
    assert torch.allclose(got, want["vision.blocks.1.wqkv"].to(torch.bfloat16).float())


def test_launch_server_serves_a_vlm_checkpoint_directory(tmp_path):
    """Stand-alone remote instance: ``launch_server --model-path <HF dir>`` loads text stack + tower and answers an ``image_data``
    request with the same tokens as an in-process engine built from the same checkpoint."""
    import json
    import os
    import subprocess
    import sys
    import time
    import urllib.request
    from pathlib import Path

    from polyrl_b200.models import hf_io
    from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams

    cfg = get_model_config("tiny-qwen2-vl")
    w, vw = tf.TransformerWeights(cfg, seed=21), V.VisionWeights(cfg, seed=22)
    hf_io.save_hf({**w.state_dict(), **vw.state_dict()}, cfg, tmp_path / "ck", dtype=torch.float32)
    grids = [(1, 4, 6)]
    pix = torch.randn(24, V.patch_dim(cfg), generator=torch.Generator().manual_seed(0))
    wire = V.encode_mm_wire({"pixel_values": pix, "grid_thw": grids})
    prompt = [2, 21, 499, 500, 498, 22, 23]
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=4, max_model_len=64, kv_cache_tokens=256)
    eng.set_weights(w.embed, w.layers, w.head)
    eng.set_vision_weights(vw)
    want = eng.generate([prompt], SamplingParams(max_new_tokens=6, temperature=0.0, ignore_eos=True), mm=[V.decode_mm_wire(wire)])[0][0].output_ids
    root = Path(__file__).resolve().parent.parent
    port = 32481
    p = subprocess.Popen([sys.executable, "-m", "polyrl_b200.rollout.launch_server", "--model-path", str(tmp_path / "ck"), "--device", "cpu",
                          "--port", str(port), "--max-num-seqs", "4", "--max-model-len", "64"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env={**os.environ, "PYTHONPATH": str(root), "OMP_NUM_THREADS": "2"}, cwd=root)
    try:
        ok = False
        for _ in range(240):
            try:
                with urllib.request.urlopen(f"http://127.0.0.1:{port}/health_generate", timeout=2) as r:
                    ok = r.status == 200
                    break
            except Exception:
                time.sleep(0.5)
                assert p.poll() is None, p.stdout.read()[-3000:]
        assert ok
        body = {"input_ids": prompt, "image_data": wire, "sampling_params": {"max_new_tokens": 6, "temperature": 0.0, "ignore_eos": True}}
        req = urllib.request.Request(f"http://127.0.0.1:{port}/generate", method="POST", headers={"Content-Type": "application/json"},
                                     data=json.dumps(body).encode())
        with urllib.request.urlopen(req, timeout=120) as r:
            out = json.loads(r.read().decode())
        assert out["output_ids"] == want
        urllib.request.urlopen(urllib.request.Request(f"http://127.0.0.1:{port}/shutdown", method="POST", data=b"{}",
                                                      headers={"Content-Type": "application/json"}), timeout=30).read()
        assert p.wait(timeout=120) == 0
    finally:
        if p.poll() is None:
            p.terminate()


# --------------------------------------------------------------------------------------------- Qwen3-VL
def _hf_tiny_q3():
    from transformers import Qwen3VLConfig, Qwen3VLForConditionalGeneration

    torch.manual_seed(0)
    hc = Qwen3VLConfig(
        text_config=dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=4, num_attention_heads=4,
                         num_key_value_heads=2, head_dim=32, max_position_embeddings=512, rms_norm_eps=1e-6,
                         rope_parameters=dict(rope_type="default", rope_theta=10000.0, mrope_section=[6, 5, 5], mrope_interleaved=True)),
        vision_config=dict(depth=3, hidden_size=64, intermediate_size=96, num_heads=4, patch_size=4, spatial_merge_size=2,
                           temporal_patch_size=2, in_channels=3, out_hidden_size=128, num_position_embeddings=36,
                           deepstack_visual_indexes=[0, 2], hidden_act="gelu_pytorch_tanh"),
        image_token_id=500, video_token_id=501, vision_start_token_id=499, vision_end_token_id=498, tie_word_embeddings=False)
    hc._attn_implementation = "eager"
    model = Qwen3VLForConditionalGeneration(hc).eval().float()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return hc, model


def test_qwen3_vl_tower_deepstack_and_logits_match_hf():
    """Qwen3-VL: learned + interpolated position embeddings, GELU-tanh LayerNorm blocks, DeepStack features added after the first
    decoder layers, interleaved M-RoPE (reference: 3rdparty/verl/verl/models/transformers/qwen3_vl.py)."""
    hc, model = _hf_tiny_q3()
    cfg, w, vw = _ours(hc, model)
    assert V.is_v3(cfg) and V.deepstack_levels(cfg) == 2 and cfg.mrope_interleaved and cfg.qk_norm and not cfg.qkv_bias
    g = torch.Generator().manual_seed(1)
    grids = [(1, 4, 6), (1, 8, 4)]
    pix = torch.randn(sum(t * h * w_ for t, h, w_ in grids), V.patch_dim(cfg), generator=g)
    grid_t = torch.tensor(grids)
    with torch.no_grad():
        out = model.model.visual(pix, grid_thw=grid_t)
        mine = V.vision_forward(cfg, vw, pix, grids)
    n_tok = sum(V.num_image_tokens(cfg, grids))
    assert mine.shape == (n_tok, V.embed_width(cfg)) == (n_tok, 3 * cfg.hidden_size)
    assert (mine[:, : cfg.hidden_size] - out.pooler_output).abs().max() < 2e-4
    for k, f in enumerate(out.deepstack_features):
        assert (mine[:, cfg.hidden_size * (k + 1): cfg.hidden_size * (k + 2)] - f).abs().max() < 2e-4

    n0, n1 = V.num_image_tokens(cfg, grids)
    ids = [2, 11, 12, 499] + [500] * n0 + [498, 13, 14, 15, 499] + [500] * n1 + [498, 16, 17]
    ids_t = torch.tensor([ids])
    mm_type = (ids_t == 500).int()
    with torch.no_grad():
        ref_logits = model(input_ids=ids_t, pixel_values=pix, image_grid_thw=grid_t, mm_token_type_ids=mm_type).logits[0]
    pos3, delta = V.mrope_positions(cfg, ids, grids)
    hf_pos, hf_delta = model.model.get_rope_index(ids_t, mm_type, image_grid_thw=grid_t)
    assert torch.equal(pos3, hf_pos[:, 0]) and delta == int(hf_delta[0, 0])
    t = len(ids)
    inp = tf.PackedInput(torch.tensor(ids), pos3[0].clone(), torch.tensor([0, t], dtype=torch.int32), t, mrope_pos=pos3,
                         image_embeds=mine)
    with torch.no_grad():
        logits = tf.logits_forward(cfg, w.head, w.embed, tf.forward_hidden(cfg, w, inp))
    assert (logits - ref_logits).abs().max() < 5e-4
    # without the DeepStack columns the logits must differ (the levels are really consumed)
    inp2 = tf.PackedInput(torch.tensor(ids), pos3[0].clone(), torch.tensor([0, t], dtype=torch.int32), t, mrope_pos=pos3,
                          image_embeds=torch.cat([mine[:, : cfg.hidden_size], torch.zeros_like(mine[:, cfg.hidden_size:])], dim=-1))
    with torch.no_grad():
        assert (tf.logits_forward(cfg, w.head, w.embed, tf.forward_hidden(cfg, w, inp2)) - ref_logits).abs().max() > 1e-3
    # checkpoint names round-trip and the exported config parses back to the same model
    sd = model.state_dict()
    back = V.export_hf_vision_state_dict(cfg, V.convert_hf_vision_state_dict(cfg, sd))
    vis_keys = [k for k in sd if k.startswith("model.visual.")]
    assert set(back) == set(vis_keys) and all(torch.equal(back[k].reshape(sd[k].shape), sd[k]) for k in vis_keys)
    from polyrl_b200.models.hf_io import hf_config_dict

    cfg2 = ModelConfig.from_hf_config(hf_config_dict(cfg))
    assert cfg2.vision == cfg.vision and cfg2.mrope_section == cfg.mrope_section and cfg2.mrope_interleaved


def test_qwen3_vl_preprocess_matches_hf_processor():
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor

    cfg = get_model_config("qwen3-vl-8b")
    rng = np.random.default_rng(1)
    img = Image.fromarray(rng.integers(0, 255, size=(90, 150, 3), dtype=np.uint8))
    mine, grid = V.preprocess_image(cfg, img, min_pixels=64 * 64, max_pixels=32 * 32 * 64)
    proc = Qwen2VLImageProcessor(min_pixels=64 * 64, max_pixels=32 * 32 * 64, patch_size=16, merge_size=2, temporal_patch_size=2,
                                 image_mean=[0.5, 0.5, 0.5], image_std=[0.5, 0.5, 0.5])          # the Qwen3-VL processor settings
    out = proc(images=[img], return_tensors="pt")
    assert tuple(out["image_grid_thw"][0].tolist()) == grid and mine.shape == out["pixel_values"].shape
    assert (mine - out["pixel_values"]).abs().max() < 0.05 and (mine - out["pixel_values"]).abs().mean() < 5e-3


def test_video_files_are_decoded_and_sampled(tmp_path):
    """Video FILES in dataset rows (reference: vision_utils.process_video / qwen_vl_utils fetch_video): uniform sampling at the
    requested rate, frame count a multiple of the temporal patch, start / end windows, frame lists untouched."""
    cv2 = pytest.importorskip("cv2")
    from polyrl_b200.data.rl_dataset import load_video

    path = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 10.0, (32, 24))
    assert wr.isOpened()
    for i in range(40):                                              # 4 seconds at 10 fps, brightness encodes the frame index
        wr.write(np.full((24, 32, 3), 5 * i + 20, dtype=np.uint8))
    wr.release()
    frames, rate = load_video(path, fps=2.0, min_frames=4)
    assert len(frames) == 8 and abs(rate - 2.0) < 1e-6 and frames[0].size == (32, 24)
    lum = [np.asarray(f).mean() for f in frames]
    assert all(b > a for a, b in zip(lum, lum[1:])) and abs(lum[0] - 20) < 6 and abs(lum[-1] - (5 * 39 + 20)) < 6
    assert len(load_video({"path": path, "nframes": 5})[0]) == 4            # rounded to the temporal patch
    f2, r2 = load_video({"video": path, "video_start": 1.0, "video_end": 2.0, "fps": 4.0})
    assert len(f2) == 4 and 60 < np.asarray(f2[0]).mean() < 85             # starts at frame 10
    assert len(load_video([np.zeros((8, 8, 3), dtype=np.uint8)] * 3)[0]) == 3
    with pytest.raises(FileNotFoundError):
        load_video(str(tmp_path / "missing.mp4"))
    # through the tower preprocessing: the decoded clip becomes (t, gh, gw) patches
    cfg = get_model_config("tiny-qwen2-vl")
    pv, grid = V.preprocess_video(cfg, frames, min_pixels=16 * 16, max_pixels=32 * 32)
    assert grid[0] == 4 and pv.shape[0] == grid[0] * grid[1] * grid[2]


def test_qwen3_vl_moe_logits_match_hf():
    """Qwen3-VL-MoE = the Qwen3-VL tower / DeepStack / interleaved M-RoPE on a mixture-of-experts text stack (stacked expert tensors)."""
    from transformers import Qwen3VLMoeConfig, Qwen3VLMoeForConditionalGeneration

    torch.manual_seed(0)
    hc = Qwen3VLMoeConfig(
        text_config=dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, head_dim=32, max_position_embeddings=512, num_experts=4, num_experts_per_tok=2,
                         moe_intermediate_size=128, decoder_sparse_step=1, mlp_only_layers=[],
                         rope_parameters=dict(rope_type="default", rope_theta=10000.0, mrope_section=[6, 5, 5], mrope_interleaved=True)),
        vision_config=dict(depth=2, hidden_size=64, intermediate_size=96, num_heads=4, patch_size=4, spatial_merge_size=2,
                           temporal_patch_size=2, in_channels=3, out_hidden_size=128, num_position_embeddings=36, deepstack_visual_indexes=[0]),
        image_token_id=500, video_token_id=501, vision_start_token_id=499, vision_end_token_id=498, tie_word_embeddings=False)
    hc._attn_implementation = "eager"
    model = Qwen3VLMoeForConditionalGeneration(hc).eval().float()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
            elif "experts" in n:
                p.normal_(0, 0.05)                       # HF leaves the stacked expert tensors uninitialised (torch.empty)
    cfg, w, vw = _ours(hc, model)
    assert cfg.num_experts == 4 and cfg.norm_topk_prob and V.is_v3(cfg) and V.deepstack_levels(cfg) == 1
    g = torch.Generator().manual_seed(1)
    grids = [(1, 4, 4)]
    pix = torch.randn(16, V.patch_dim(cfg), generator=g)
    n0 = V.num_image_tokens(cfg, grids)[0]
    ids = [2, 11, 499] + [500] * n0 + [498, 13, 14, 15]
    ids_t = torch.tensor([ids])
    with torch.no_grad():
        ref_logits = model(input_ids=ids_t, pixel_values=pix, image_grid_thw=torch.tensor(grids), mm_token_type_ids=(ids_t == 500).int()).logits[0]
        img = V.vision_forward(cfg, vw, pix, grids)
        pos3, _ = V.mrope_positions(cfg, ids, grids)
        t = len(ids)
        inp = tf.PackedInput(torch.tensor(ids), pos3[0].clone(), torch.tensor([0, t], dtype=torch.int32), t, mrope_pos=pos3, image_embeds=img)
        logits = tf.logits_forward(cfg, w.head, w.embed, tf.forward_hidden(cfg, w, inp))
    assert (logits - ref_logits).abs().max() < 5e-4
    from polyrl_b200.models.hf_io import hf_config_dict

    cfg2 = ModelConfig.from_hf_config(hf_config_dict(cfg))
    assert cfg2.num_experts == 4 and cfg2.vision == cfg.vision


def _hf_tiny_glm4v():
    from transformers import Glm4vConfig, Glm4vForConditionalGeneration

    torch.manual_seed(0)
    hc = Glm4vConfig(
        text_config=dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=512, rms_norm_eps=1e-5, pad_token_id=0,
                         rope_parameters=dict(rope_type="default", rope_theta=10000.0, mrope_section=[2, 3, 3], partial_rotary_factor=0.5)),
        vision_config=dict(depth=2, hidden_size=64, intermediate_size=96, num_heads=4, patch_size=4, image_size=24, spatial_merge_size=2,
                           temporal_patch_size=2, in_channels=3, out_hidden_size=128, attention_bias=True, rms_norm_eps=1e-5),
        image_token_id=500, video_token_id=501, image_start_token_id=499, image_end_token_id=498, tie_word_embeddings=False)
    hc._attn_implementation = "eager"
    model = Glm4vForConditionalGeneration(hc).eval().float()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return hc, model


def test_glm4v_tower_positions_and_logits_match_hf():
    """GLM-4V: RMSNorm after the patch embedding, bicubically resampled position table, bias-free SwiGLU blocks, convolutional cell
    merge + gated merger; GLM-4 text blocks (output norms, partial interleaved rotary) under M-RoPE
    (reference: 3rdparty/verl/verl/models/transformers/glm4v.py)."""
    hc, model = _hf_tiny_glm4v()
    cfg, w, vw = _ours(hc, model)
    assert V.is_glm4v(cfg) and cfg.sandwich_norm and cfg.partial_rotary_factor == 0.5 and cfg.rope_interleaved and cfg.qkv_bias
    g = torch.Generator().manual_seed(1)
    grids = [(1, 4, 6), (1, 8, 4)]
    pix = torch.randn(sum(t * h * w_ for t, h, w_ in grids), V.patch_dim(cfg), generator=g)
    grid_t = torch.tensor(grids)
    with torch.no_grad():
        out = model.model.visual(pix, grid_thw=grid_t).pooler_output
        mine = V.vision_forward(cfg, vw, pix, grids)
    assert mine.shape == out.shape == (sum(V.num_image_tokens(cfg, grids)), cfg.hidden_size)
    assert (mine - out).abs().max() < 2e-4
    n0, n1 = V.num_image_tokens(cfg, grids)
    ids = [2, 11, 12, 499] + [500] * n0 + [498, 13, 14, 15, 499] + [500] * n1 + [498, 16, 17]
    ids_t = torch.tensor([ids])
    mm_type = (ids_t == 500).int()
    with torch.no_grad():
        ref_logits = model(input_ids=ids_t, pixel_values=pix, image_grid_thw=grid_t, mm_token_type_ids=mm_type).logits[0]
    pos3, delta = V.mrope_positions(cfg, ids, grids)
    hf_pos, hf_delta = model.model.get_rope_index(ids_t, mm_type, image_grid_thw=grid_t)
    assert torch.equal(pos3, hf_pos[:, 0]) and delta == int(hf_delta[0, 0])
    t = len(ids)
    inp = tf.PackedInput(torch.tensor(ids), pos3[0].clone(), torch.tensor([0, t], dtype=torch.int32), t, mrope_pos=pos3, image_embeds=mine)
    with torch.no_grad():
        logits = tf.logits_forward(cfg, w.head, w.embed, tf.forward_hidden(cfg, w, inp))
    assert (logits - ref_logits).abs().max() < 5e-4
    # text-only positions (all three streams equal) are NOT what the model sees on image tokens
    flat = tf.PackedInput(torch.tensor(ids), torch.arange(t), torch.tensor([0, t], dtype=torch.int32), t, image_embeds=mine)
    with torch.no_grad():
        assert (tf.logits_forward(cfg, w.head, w.embed, tf.forward_hidden(cfg, w, flat)) - ref_logits).abs().max() > 1e-3
    # checkpoint names round-trip (tower and text stack) and the exported config parses back to the same model
    sd = model.state_dict()
    back = V.export_hf_vision_state_dict(cfg, V.convert_hf_vision_state_dict(cfg, sd))
    vis_keys = [k for k in sd if k.startswith("model.visual.")]
    assert set(back) == set(vis_keys) and all(torch.equal(back[k].reshape(sd[k].shape), sd[k]) for k in vis_keys)
    txt = tf.export_hf_state_dict(cfg, w.state_dict())
    for k, v in txt.items():
        hk = k.replace("model.", "model.language_model.", 1) if k.startswith("model.") else k
        assert torch.equal(v, sd[hk]), k
    from polyrl_b200.models.hf_io import hf_config_dict

    cfg2 = ModelConfig.from_hf_config(hf_config_dict(cfg))
    assert cfg2.vision == cfg.vision and cfg2.mrope_section == cfg.mrope_section and cfg2.sandwich_norm and cfg2.partial_rotary_factor == 0.5


def test_glm4v_preprocess_matches_hf_processor():
    """GLM-4V image processor: CLIP statistics, the resize budget counts temporal-patch x height x width voxels."""
    from PIL import Image
    from transformers.models.glm4v.image_processing_glm4v import Glm4vImageProcessor

    cfg = get_model_config("glm-4.1v-9b")
    rng = np.random.default_rng(0)
    proc = Glm4vImageProcessor()
    for size in ((75, 131), (900, 40), (1900, 2400)):
        img = Image.fromarray(rng.integers(0, 255, size=(*size, 3), dtype=np.uint8))
        mine, grid = V.preprocess_image(cfg, img)
        out = proc(images=[img], return_tensors="pt")
        assert tuple(out["image_grid_thw"][0].tolist()) == grid, (size, grid, out["image_grid_thw"])
        assert mine.shape == out["pixel_values"].shape
        assert (mine - out["pixel_values"]).abs().mean() < 5e-3
