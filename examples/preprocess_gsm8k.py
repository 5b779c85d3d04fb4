"""Turn a local GSM8K jsonl ({"question":..., "answer": "... #### 42"}) into the RL parquet/jsonl schema
(prompt chat list, data_source, reward_model.ground_truth) — offline counterpart of the reference's
examples/data_preprocess (gsm8k.py / OpenR1 preprocessing)."""
import argparse
import json
import re

from polyrl_b200.data.rl_dataset import gsm8k_prompt

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    a = ap.parse_args()
    with open(a.src) as f, open(a.dst, "w") as g:
        for i, line in enumerate(f):
            row = json.loads(line)
            m = re.search(r"#### (\-?[0-9\.\,]+)", row["answer"])
            gt = m.group(1).replace(",", "") if m else ""
            g.write(json.dumps({"data_source": "openai/gsm8k", "prompt": gsm8k_prompt(row["question"]), "ability": "math",
                                "reward_model": {"style": "rule", "ground_truth": gt}, "extra_info": {"index": i}}) + "\n")
