"""TP rollout plumbing on CPU/gloo (2 ranks): TP=2 engine shards (library all-reduce path) reproduce the single-engine greedy
generations, and the colocated worker with rollout.tensor_model_parallel_size=2 returns every rank its own slice."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    from polyrl_b200.models import TransformerWeights, get_model_config
    from polyrl_b200.rollout.engine import RolloutEngine
    from polyrl_b200.rollout.tp import TPContext

    cfg = get_model_config("tiny-llama")
    w = TransformerWeights(cfg, "cpu", torch.float32, seed=31)
    prompts = [[5 + i, 9, 33, 7, 2 + i, 14] for i in range(5)]
    sp = {"max_new_tokens": 10, "temperature": 0.0, "ignore_eos": True}
    tp = TPContext(cfg, None, rank, world, max_batch=8, group=dist.group.WORLD)
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=8, max_model_len=64, use_cuda_graph=False, seed=3, tp=tp)
    eng.set_weights(w.embed, w.layers, w.head)
    got = [(g[0].output_ids, g[0].logprobs) for g in eng.generate(prompts, sp)]
    ref = RolloutEngine(cfg, "cpu", max_num_seqs=8, max_model_len=64, use_cuda_graph=False, seed=3)
    ref.set_weights(w.embed, w.layers, w.head)
    want = [(g[0].output_ids, g[0].logprobs) for g in ref.generate(prompts, sp)]
    tok_ok = all(a[0] == b[0] for a, b in zip(got, want))
    lp_err = max(abs(x - y) for a, b in zip(got, want) for x, y in zip(a[1], b[1]))
    both = [None] * world
    dist.all_gather_object(both, [g[0] for g in got])
    lock = all(b == both[0] for b in both)

    # colocated worker: each rank submits DIFFERENT prompts, gets its own slice back
    from polyrl_b200.config.schema import load_config
    from polyrl_b200.protocol import DataProto
    from polyrl_b200.workers.hybrid_worker import ActorRolloutRefWorker

    c = load_config(None, ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-llama", "data.max_prompt_length=8",
                           "data.max_response_length=6", "actor_rollout_ref.rollout.n=2", "actor_rollout_ref.rollout.max_num_seqs=8",
                           "actor_rollout_ref.rollout.cuda_graph=False", "actor_rollout_ref.rollout.tensor_model_parallel_size=2",
                           "actor_rollout_ref.actor.fsdp_config.comm_backend=nccl"])
    wk = ActorRolloutRefWorker(c, role="actor_rollout", device="cpu")
    wk.init_model()
    ids = torch.full((2, 8), 7 + rank, dtype=torch.long)
    ids[1, :3] = 40 + rank
    batch = DataProto.from_dict({"input_ids": ids, "attention_mask": torch.ones_like(ids), "position_ids": torch.arange(8).repeat(2, 1)})
    batch.meta_info = {"do_sample": False}
    out = wk.generate_sequences(batch)
    c2 = load_config(None, ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-llama", "data.max_prompt_length=8",
                            "data.max_response_length=6", "actor_rollout_ref.rollout.n=2", "actor_rollout_ref.rollout.max_num_seqs=8",
                            "actor_rollout_ref.rollout.cuda_graph=False", "actor_rollout_ref.rollout.tensor_model_parallel_size=2",
                            "actor_rollout_ref.actor.fsdp_config.comm_backend=nccl",
                            "actor_rollout_ref.actor.fsdp_config.resident_params=false"])
    wk2 = ActorRolloutRefWorker(c2, role="actor_rollout", device="cpu")       # ZeRO-3 + block-by-block TP-sliced gather
    wk2.init_model()
    out2 = wk2.generate_sequences(batch)
    zero3_ok = (not wk2.actor_engine.resident) and torch.equal(out2.batch["responses"], out.batch["responses"])
    mine_ok = out.batch["responses"].shape == (4, 6) and torch.equal(out.batch["prompts"][::2], ids)
    resp = [None] * world
    dist.all_gather_object(resp, out.batch["responses"].tolist())
    differ = resp[0] != resp[1]                      # different prompts -> different slices came back
    # ---- vision-language prompts through the TP engine and the colocated TP worker (tower replicated, M-RoPE per rank)
    import numpy as np

    from polyrl_b200.models import vision as V

    vcfg = get_model_config("tiny-qwen2-vl")
    vw_, vt = TransformerWeights(vcfg, "cpu", torch.float32, seed=5), V.VisionWeights(vcfg, seed=6)
    g = torch.Generator().manual_seed(2)
    grids = [(1, 4, 6)]
    pix = torch.randn(24, V.patch_dim(vcfg), generator=g)
    vprompts, vmm = [[2, 21, 499, 500, 498, 22], [2, 40, 41, 42]], [dict(pixel_values=pix, grid_thw=grids), None]
    outs_v = []
    for ctx in (TPContext(vcfg, None, rank, world, max_batch=8, group=dist.group.WORLD), None):
        e = RolloutEngine(vcfg, "cpu", max_num_seqs=8, max_model_len=64, use_cuda_graph=False, seed=3, tp=ctx)
        e.set_weights(vw_.embed, vw_.layers, vw_.head)
        e.set_vision_weights(vt)
        outs_v.append([(q[0].output_ids, q[0].logprobs) for q in e.generate(vprompts, sp, mm=vmm)])
    vlm_tok_ok = all(a[0] == b[0] for a, b in zip(*outs_v))
    vlm_lp_err = max(abs(x - y) for a, b in zip(*outs_v) for x, y in zip(a[1], b[1]))
    cv = load_config(None, ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-qwen2-vl", "data.max_prompt_length=16",
                            "data.max_response_length=6", "actor_rollout_ref.rollout.n=2", "actor_rollout_ref.rollout.max_num_seqs=8",
                            "actor_rollout_ref.rollout.cuda_graph=False", "actor_rollout_ref.rollout.tensor_model_parallel_size=2",
                            "actor_rollout_ref.actor.fsdp_config.comm_backend=nccl"])
    wv = ActorRolloutRefWorker(cv, role="actor_rollout", device="cpu")
    wv.init_model()
    raw = V.expand_image_placeholders(vcfg, [2, 21 + rank, 499, 500, 498, 22], grids)          # rank-specific prompt, same image
    pos3, _ = V.mrope_positions(vcfg, raw, grids)
    vids = torch.zeros(1, 16, dtype=torch.long)
    vmask = torch.zeros(1, 16, dtype=torch.long)
    vpos = torch.zeros(1, 3, 16, dtype=torch.long)
    vids[0, 16 - len(raw):], vmask[0, 16 - len(raw):], vpos[0, :, 16 - len(raw):] = torch.tensor(raw), 1, pos3
    vb = DataProto.from_dict({"input_ids": vids, "attention_mask": vmask, "position_ids": vpos},
                             non_tensors={"raw_prompt_ids": np.array([raw], dtype=object),
                                          "multi_modal_inputs": np.array([{"pixel_values": pix, "image_grid_thw": torch.tensor(grids)}], dtype=object)})
    vb.meta_info = {"temperature": 1.0, "ignore_eos": True}
    vo = wv.generate_sequences(vb)
    vo.meta_info.update(temperature=1.0)
    lp_tr = wv.compute_log_prob(vo).batch["old_log_probs"]
    assert vo.batch["rollout_log_probs"].abs().min() > 1e-3 and vo.batch["response_mask"].sum() == 12
    vlm_worker_err = (lp_tr - vo.batch["rollout_log_probs"]).abs().max().item()            # TP engine == trainer forward
    # ---- mixture of experts through the TP engine: experts sharded whole over the ranks, partial outputs all-reduced
    mcfg = get_model_config("tiny-qwen3-moe")
    mw = TransformerWeights(mcfg, "cpu", torch.float32, seed=13)
    outs_m = []
    for ctx in (TPContext(mcfg, None, rank, world, max_batch=8, group=dist.group.WORLD), None):
        e = RolloutEngine(mcfg, "cpu", max_num_seqs=8, max_model_len=64, use_cuda_graph=False, seed=3, tp=ctx)
        e.set_weights(mw.embed, mw.layers, mw.head)
        if ctx is not None:
            assert e.layer_w[0]["ewgu"].shape[0] == mcfg.num_experts // world and "wgu" not in e.layer_w[0]
        outs_m.append([(q[0].output_ids, q[0].logprobs) for q in e.generate(prompts, sp)])
    moe_tok_ok = all(a[0] == b[0] for a, b in zip(*outs_m))
    moe_lp_err = max(abs(x - y) for a, b in zip(*outs_m) for x, y in zip(a[1], b[1]))
    # ---- DeepSeek-V3 blocks: latent attention with the heads split over the ranks and the latent cache replicated, on top of the
    # expert-sharded MoE; with the query bottleneck and with the plain query projection fused into the first-stage GEMM
    mla_ok, mla_err = True, 0.0
    for qlr in (48, 0):
        dcfg = get_model_config("tiny-deepseek-v3", q_lora_rank=qlr)
        dw = TransformerWeights(dcfg, "cpu", torch.float32, seed=19)
        outs_d = []
        for ctx in (TPContext(dcfg, None, rank, world, max_batch=8, group=dist.group.WORLD), None):
            e = RolloutEngine(dcfg, "cpu", max_num_seqs=8, max_model_len=64, use_cuda_graph=False, seed=3, tp=ctx)
            e.set_weights(dw.embed, dw.layers, dw.head)
            if ctx is not None:
                assert e.layer_w[1]["wkv_up"].shape[0] == dw.layers[1]["wkv_up"].shape[0] // world and e.cfg.num_heads == dcfg.num_heads // world
            outs_d.append([(q[0].output_ids, q[0].logprobs) for q in e.generate(prompts, sp)])
        mla_ok = mla_ok and all(a[0] == b[0] for a, b in zip(*outs_d))
        mla_err = max(mla_err, max(abs(x - y) for a, b in zip(*outs_d) for x, y in zip(a[1], b[1])))
    # ---- Apertus blocks: non-gated xIELU MLP split column / row wise
    acfg = get_model_config("tiny-apertus")
    aw = TransformerWeights(acfg, "cpu", torch.float32, seed=23)
    outs_a = []
    for ctx in (TPContext(acfg, None, rank, world, max_batch=8, group=dist.group.WORLD), None):
        e = RolloutEngine(acfg, "cpu", max_num_seqs=8, max_model_len=64, use_cuda_graph=False, seed=3, tp=ctx)
        e.set_weights(aw.embed, aw.layers, aw.head)
        outs_a.append([(q[0].output_ids, q[0].logprobs) for q in e.generate(prompts, sp)])
    mla_ok = mla_ok and all(a[0] == b[0] for a, b in zip(*outs_a))
    mla_err = max(mla_err, max(abs(x - y) for a, b in zip(*outs_a) for x, y in zip(a[1], b[1])))
    if rank == 0:
        print(f"TP_CPU_MLA tok_ok={int(mla_ok)} lp_err={mla_err:.2e}", flush=True)
    # ---- GLM-4 blocks (output norms after the row-parallel all-reduces, partial rotary, qkv bias) through the TP engine
    gcfg = get_model_config("tiny-glm4")
    gw = TransformerWeights(gcfg, "cpu", torch.float32, seed=17)
    for lw in gw.layers:
        for k in ("ln_attn", "ln_mlp"):
            lw[k].add_(0.2 * torch.randn(lw[k].shape, generator=torch.Generator().manual_seed(3)))
    outs_g = []
    for ctx in (TPContext(gcfg, None, rank, world, max_batch=8, group=dist.group.WORLD), None):
        e = RolloutEngine(gcfg, "cpu", max_num_seqs=8, max_model_len=64, use_cuda_graph=False, seed=3, tp=ctx)
        e.set_weights(gw.embed, gw.layers, gw.head)
        outs_g.append([(q[0].output_ids, q[0].logprobs) for q in e.generate(prompts, sp)])
    glm_tok_ok = all(a[0] == b[0] for a, b in zip(*outs_g))
    glm_lp_err = max(abs(x - y) for a, b in zip(*outs_g) for x, y in zip(a[1], b[1]))
    if rank == 0:
        print(f"TP_CPU_GLM tok_ok={int(glm_tok_ok)} lp_err={glm_lp_err:.2e}", flush=True)
    if rank == 0:
        print(f"TP_CPU_MOE tok_ok={int(moe_tok_ok)} lp_err={moe_lp_err:.2e}", flush=True)
    if rank == 0:
        print(f"TP_CPU_VLM tok_ok={int(vlm_tok_ok)} lp_err={vlm_lp_err:.2e} worker_err={vlm_worker_err:.2e}", flush=True)
    if rank == 0:
        print(f"TP_CPU_RESULTS tok_ok={int(tok_ok)} lp_err={lp_err:.2e} lock={int(lock)} slice_ok={int(mine_ok)} differ={int(differ)} zero3_tp_slices_ok={int(zero3_ok)}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
