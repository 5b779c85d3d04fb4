"""OpenR1-Math-220k -> parquet/jsonl rows in the RLHF dataset schema (data_source / prompt / reward_model / extra_info).
Capability parity: examples/data_preprocess/openr1.py of the reference (same output fields; local files only — no hub access).

    python examples/preprocess_openr1.py --input openr1.jsonl --output-dir data/openr1 --val-fraction 0.02
"""
import argparse
import json
import random
from pathlib import Path

INSTRUCTION = "Let's think step by step and output the final answer within \\boxed{}."


def convert(row: dict, idx: int, split: str) -> dict | None:
    q, a = row.get("problem") or row.get("question"), row.get("answer")
    if not q or a is None:
        return None
    return {"data_source": "open-r1/OpenR1-Math-220k",
            "prompt": [{"role": "user", "content": f"{q.strip()} {INSTRUCTION}"}],
            "ability": "math",
            "reward_model": {"style": "rule", "ground_truth": str(a).strip()},
            "extra_info": {"split": split, "index": idx, "problem_type": row.get("problem_type"), "source": row.get("source")}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--input", required=True, help="jsonl (one problem per line) or parquet exported from the hub")
    ap.add_argument("--output-dir", default="data/openr1")
    ap.add_argument("--val-fraction", type=float, default=0.02)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    src = Path(a.input)
    if src.suffix == ".parquet":
        import pandas as pd

        rows = pd.read_parquet(src).to_dict("records")
    else:
        rows = [json.loads(l) for l in src.read_text().splitlines() if l.strip()]
    random.Random(a.seed).shuffle(rows)
    n_val = int(len(rows) * a.val_fraction)
    out = Path(a.output_dir)
    out.mkdir(parents=True, exist_ok=True)
    for split, part in (("test", rows[:n_val]), ("train", rows[n_val:])):
        conv = [c for i, r in enumerate(part) if (c := convert(r, i, split)) is not None]
        with open(out / f"{split}.jsonl", "w") as f:
            for c in conv:
                f.write(json.dumps(c) + "\n")
        try:
            import pandas as pd

            pd.DataFrame(conv).to_parquet(out / f"{split}.parquet")
        except Exception:
            pass
        print(f"{split}: {len(conv)} rows -> {out}")


if __name__ == "__main__":
    main()
