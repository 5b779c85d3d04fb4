"""Offline batch generation: sample ``n_samples`` responses for every prompt of a dataset and write them next to the rows.

Counterpart of the reference's ``verl/trainer/main_generation.py`` (Ray worker group over the rollout workers, parquet in / parquet
out).  Here: SPMD — every rank (``torchrun`` or a single process) owns one rollout engine replica, takes a contiguous slice of the
prompts and rank 0 writes the merged table.  The engine is the training-time rollout engine (paged KV, CUDA-graph decode), so the
samples are the ones the trainer would see.

    python -m polyrl_b200.trainer.main_generation --model /ckpts/qwen2.5-7b --data test.parquet --output gen.parquet \\
        --n-samples 4 --temperature 1.0 --response-length 1024 [--prompt-key prompt] [--batch-size 256]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch


def _write(rows: list[dict], path: str) -> None:
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    if path.endswith(".parquet"):
        import pyarrow as pa
        import pyarrow.parquet as pq

        pq.write_table(pa.Table.from_pylist(rows), path)
    else:
        with open(path, "w") as f:
            for r in rows:
                f.write(json.dumps(r, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o)) + "\n")


def generate(model: str, rows: list[dict], n_samples: int = 1, temperature: float = 1.0, top_p: float = 1.0, top_k: int = -1,
             response_length: int = 512, prompt_length: int = 1024, prompt_key: str = "prompt", batch_size: int = 256,
             tokenizer_path: str | None = None, device=None, seed: int = 0, weights=None) -> list[dict]:
    """-> the rows with ``responses`` (list of n strings, or token-id lists without a tokenizer) and ``response_ids`` added.
    Data parallel when ``torch.distributed`` is initialised: rank r handles rows r::world and rank 0 returns everything."""
    import torch.distributed as dist

    from ..data.rl_dataset import load_tokenizer
    from ..models import TransformerWeights, get_model_config
    from ..rollout.engine import RolloutEngine, SamplingParams

    if temperature == 0.0 and n_samples != 1:
        raise ValueError("temperature 0 is greedy: n_samples must be 1")
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = torch.device(device or (f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}" if torch.cuda.is_available() else "cpu"))
    cfg = get_model_config(model)
    tok = None
    is_dir = Path(model).is_dir()
    if tokenizer_path or is_dir:
        tok = load_tokenizer(tokenizer_path or model)
    dtype = torch.bfloat16 if dev.type == "cuda" else torch.float32
    eng = RolloutEngine(cfg, dev, max_num_seqs=min(256, max(8, batch_size)), max_model_len=prompt_length + response_length,
                        use_cuda_graph=dev.type == "cuda", seed=seed + 17 * rank)
    if weights is None:
        weights = TransformerWeights(cfg, device=dev, dtype=dtype, init=not is_dir, seed=seed)
        if is_dir:
            from ..models.hf_io import load_state_dict

            weights.load_state_dict({k: v.to(dev, dtype) for k, v in load_state_dict(cfg, model).items() if not k.startswith("vision.")})
    eng.set_weights(weights.embed, weights.layers, weights.head)

    def encode(p) -> list[int]:
        if isinstance(p, (list, tuple, np.ndarray)) and len(p) and isinstance(p[0], (int, np.integer)):
            return [int(x) for x in p][-prompt_length:]
        if tok is None:
            raise ValueError("text prompts need a tokenizer (--tokenizer or a checkpoint directory)")
        if isinstance(p, (list, np.ndarray)) and len(p) and isinstance(p[0], dict):
            p = tok.apply_chat_template(list(p), add_generation_prompt=True, tokenize=False)
        return list(tok.encode(str(p), add_special_tokens=False))[-prompt_length:]

    mine = list(range(rank, len(rows), world))
    sp = SamplingParams(max_new_tokens=response_length, temperature=temperature, top_p=top_p, top_k=top_k, n=n_samples)
    out: dict[int, dict] = {}
    for lo in range(0, len(mine), batch_size):
        idx = mine[lo:lo + batch_size]
        groups = eng.generate([encode(rows[i][prompt_key]) for i in idx], sp)
        for i, grp in zip(idx, groups):
            ids = [list(r.output_ids) for r in grp]
            out[i] = {"response_ids": ids, "responses": [tok.decode(x, skip_special_tokens=True) for x in ids] if tok is not None else ids}
    if world > 1:
        parts: list = [None] * world
        dist.all_gather_object(parts, out)
        out = {k: v for p in parts for k, v in p.items()}
    return [{**rows[i], **out[i]} for i in range(len(rows))] if rank == 0 else []


_DOTTED = {"model.path": "--model", "data.path": "--data", "data.output_path": "--output", "data.prompt_key": "--prompt-key",
           "data.n_samples": "--n-samples", "data.batch_size": "--batch-size", "rollout.temperature": "--temperature",
           "rollout.top_p": "--top-p", "rollout.top_k": "--top-k", "rollout.prompt_length": "--prompt-length",
           "rollout.response_length": "--response-length", "model.tokenizer_path": "--tokenizer"}


def _translate_dotted(argv: list[str], table: dict[str, str]) -> list[str]:
    """``key=value`` overrides in the reference's generation.yaml / evaluation.yaml naming -> this CLI's flags; other dotted keys
    (trainer.nnodes, rollout.tensor_model_parallel_size, ray settings ...) are accepted and ignored."""
    out: list[str] = []
    for a in argv:
        if not a.startswith("-") and "=" in a and "." in a.split("=", 1)[0]:
            k, v = a.split("=", 1)
            if k.lstrip("+") in table:
                out += [table[k.lstrip("+")], v]
            continue
        out.append(a)
    return out


def main(argv: list[str] | None = None) -> int:
    argv = _translate_dotted(list(sys.argv[1:] if argv is None else argv), _DOTTED)
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model", required=True, help="HF checkpoint directory or a preset name (random init)")
    ap.add_argument("--data", required=True, help="parquet / jsonl / json with the prompts")
    ap.add_argument("--output", required=True, help=".parquet or .jsonl")
    ap.add_argument("--prompt-key", default="prompt")
    ap.add_argument("--tokenizer", default=None)
    ap.add_argument("--n-samples", type=int, default=1)
    ap.add_argument("--temperature", type=float, default=1.0)
    ap.add_argument("--top-p", type=float, default=1.0)
    ap.add_argument("--top-k", type=int, default=-1)
    ap.add_argument("--prompt-length", type=int, default=1024)
    ap.add_argument("--response-length", type=int, default=512)
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    import torch.distributed as dist

    from ..data.rl_dataset import _read_rows

    if int(os.environ.get("WORLD_SIZE", 1)) > 1 and not dist.is_initialized():
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    rows = generate(a.model, _read_rows(a.data), a.n_samples, a.temperature, a.top_p, a.top_k, a.response_length, a.prompt_length,
                    a.prompt_key, a.batch_size, a.tokenizer, seed=a.seed)
    if rows:
        _write(rows, a.output)
        print(f"[generation] wrote {len(rows)} rows x {a.n_samples} samples to {a.output}", flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
