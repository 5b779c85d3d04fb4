"""Geometry3K-style VLM data -> the RL schema with images.  Input: a parquet / jsonl whose rows carry ``problem`` (text containing
``<image>``), ``answer`` and ``images`` (HF datasets image cells: {"bytes": ...} / {"path": ...}).  Output rows:
prompt chat list (the ``<image>`` marks stay in the text — RLHFDataset expands them), images, data_source, reward_model.ground_truth.
Offline counterpart of the reference's examples/data_preprocess/geo3k.py."""
import argparse
import json

import pyarrow as pa
import pyarrow.parquet as pq

INSTRUCTION = ("You FIRST think about the reasoning process as an internal monologue and then provide the final answer. "
               "The reasoning process MUST BE enclosed within <think> </think> tags. The final answer MUST BE put in \\boxed{}.")

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst", help="*.parquet (keeps image bytes) or *.jsonl (images must be {'path': ...})")
    ap.add_argument("--split", default="train")
    a = ap.parse_args()
    rows = pq.read_table(a.src).to_pylist() if a.src.endswith(".parquet") else [json.loads(l) for l in open(a.src) if l.strip()]
    out = []
    for i, r in enumerate(rows):
        out.append({"data_source": "hiyouga/geometry3k", "ability": "math", "images": r["images"],
                    "prompt": [{"role": "user", "content": r["problem"] + " " + INSTRUCTION}],
                    "reward_model": {"style": "rule", "ground_truth": r["answer"]},
                    "extra_info": {"split": a.split, "index": i, "answer": r["answer"], "question": r["problem"]}})
    if a.dst.endswith(".parquet"):
        pq.write_table(pa.Table.from_pylist(out), a.dst)
    else:
        with open(a.dst, "w") as g:
            for r in out:
                g.write(json.dumps(r) + "\n")
    print(f"wrote {len(out)} rows to {a.dst}")
