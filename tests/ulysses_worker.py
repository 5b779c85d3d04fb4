"""Ulysses SP equivalence (torchrun, gloo/CPU or nccl/GPU): SP=world vs no-SP logprobs and parameter gradients."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    cuda = torch.cuda.is_available() and os.environ.get("ULYSSES_DEVICE", "auto") != "cpu"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if cuda:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    from polyrl_b200.models import get_model_config, pack_padded
    from polyrl_b200.parallel.zero3 import Zero3Engine

    cfg = get_model_config("tiny-llama-128" if cuda else "tiny-llama")
    torch.manual_seed(5)
    ids = torch.randint(3, cfg.vocab_size, (3, 21), device=dev)
    mask = torch.ones_like(ids)
    mask[1, 15:] = 0
    inp, _ = pack_padded(ids, mask)
    labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), device=dev)
    outs = {}
    variants = [("nosp", None), ("sp", dist.group.WORLD)]
    if cuda:
        variants.append(("sp_peer", dist.group.WORLD))        # same exchange over NVLink peer memory instead of NCCL
    for name, sp_group in variants:
        if name == "sp_peer":
            from polyrl_b200.parallel.symm import SymmetricHeap
            from polyrl_b200.parallel.ulysses import enable_peer_all_to_all

            enable_peer_all_to_all(dist.group.WORLD, SymmetricHeap(device=dev), 1 << 22)
        eng = Zero3Engine(cfg, dev, comm="nccl", lr=1e-2, weight_decay=0.0, seed=7, sp_group=sp_group)
        lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
        (-(lp.mean()) + 0.01 * ent.mean()).backward()
        eng.step()
        outs[name] = (lp.detach().float().cpu(), {n: t.float().cpu().clone() for n, t in eng.named_full_tensors()})
    lp_err = (outs["sp"][0] - outs["nosp"][0]).abs().max().item()
    w_err = max(((outs["sp"][1][k] - outs["nosp"][1][k]).norm() / (outs["nosp"][1][k].norm() + 1e-9)).item() for k in outs["sp"][1])
    extra = ""
    if "sp_peer" in outs:
        p_lp = (outs["sp_peer"][0] - outs["sp"][0]).abs().max().item()
        p_w = max(((outs["sp_peer"][1][k] - outs["sp"][1][k]).norm() / (outs["sp"][1][k].norm() + 1e-9)).item() for k in outs["sp"][1])
        extra = f" peer_lp_err={p_lp:.3e} peer_w_err={p_w:.3e}"
    # ---- vision-language batch under SP (CPU plumbing): image rows and 3-axis positions are sliced with the tokens
    if not cuda:
        from polyrl_b200.models import transformer as tf
        from polyrl_b200.models import vision as V

        for preset in ("tiny-qwen2-vl", "tiny-qwen3-vl", "tiny-glm4v"):      # Qwen3-VL: the DeepStack rows are sliced with the image rows
            vcfg = get_model_config(preset)
            g = torch.Generator().manual_seed(9)
            grids = [(1, 4, 4), (1, 4, 8)]
            pix = torch.randn(16 + 32, V.patch_dim(vcfg), generator=g)
            a = V.expand_image_placeholders(vcfg, [2, 11, 499, 500, 498, 12, 13, 499, 500, 498, 14, 15, 16], grids)
            b = [2, 31, 32, 33, 34, 35, 36, 37]
            pa, _ = V.mrope_positions(vcfg, a, grids)
            pos3 = torch.cat([pa, torch.arange(len(b)).expand(3, -1)], dim=1).contiguous()
            vin = tf.PackedInput(torch.tensor(a + b), pos3[0].clone(), torch.tensor([0, len(a), len(a) + len(b)], dtype=torch.int32),
                                 max(len(a), len(b)), mrope_pos=pos3, pixel_values=pix, grid_thw=grids)
            assert vin.input_ids.numel() % world != 0                      # exercises the tail padding of the slices
            vlabels = torch.randint(0, 400, (vin.input_ids.numel(),), generator=g)
            vouts = {}
            for name, sp_group in (("nosp", None), ("sp", dist.group.WORLD)):
                eng = Zero3Engine(vcfg, dev, comm="nccl", lr=1e-2, weight_decay=0.0, seed=7, sp_group=sp_group)
                lp, ent = eng.forward_logprob_entropy(vin, vlabels, 1.0)
                (-(lp.mean()) + 0.01 * ent.mean()).backward()
                eng.step()
                vouts[name] = (lp.detach().float(), {n: t.float().clone() for n, t in eng.named_full_tensors()})
            v_lp = (vouts["sp"][0] - vouts["nosp"][0]).abs().max().item()
            v_w = max(((vouts["sp"][1][k] - vouts["nosp"][1][k]).norm() / (vouts["nosp"][1][k].norm() + 1e-9)).item() for k in vouts["sp"][1])
            assert any(k.startswith("vision.") for k in vouts["sp"][1])
            assert v_lp < 1e-4 and v_w < 1e-3, (preset, v_lp, v_w)
            extra += f" vlm_lp_err[{preset}]={v_lp:.3e} vlm_w_err[{preset}]={v_w:.3e}"
    # ---- multi-head latent attention under SP (the reference's kimi_vl.py capability: Ulysses over the DeepSeek-V3 attention)
    if not cuda:
        mcfg = get_model_config("tiny-deepseek-v3")
        torch.manual_seed(6)
        mids = torch.randint(3, mcfg.vocab_size, (3, 19))
        mmask = torch.ones_like(mids)
        mmask[2, 11:] = 0
        minp, _ = pack_padded(mids, mmask)
        assert minp.input_ids.numel() % world != 0
        mlabels = torch.randint(0, mcfg.vocab_size, (minp.input_ids.numel(),))
        mouts = {}
        for name, sp_group in (("nosp", None), ("sp", dist.group.WORLD)):
            eng = Zero3Engine(mcfg, dev, comm="nccl", lr=1e-2, weight_decay=0.0, seed=7, sp_group=sp_group)
            lp, ent = eng.forward_logprob_entropy(minp, mlabels, 1.0)
            (-(lp.mean()) + 0.01 * ent.mean()).backward()
            eng.step()
            mouts[name] = (lp.detach().float(), {n: t.float().clone() for n, t in eng.named_full_tensors()})
        m_lp = (mouts["sp"][0] - mouts["nosp"][0]).abs().max().item()
        m_w = max(((mouts["sp"][1][k] - mouts["nosp"][1][k]).norm() / (mouts["nosp"][1][k].norm() + 1e-9)).item() for k in mouts["sp"][1])
        assert m_lp < 1e-4 and m_w < 1e-3, (m_lp, m_w)
        extra += f" mla_lp_err={m_lp:.3e} mla_w_err={m_w:.3e}"
    if rank == 0:
        print(f"ULYSSES_RESULTS lp_err={lp_err:.3e} w_err={w_err:.3e}{extra}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
