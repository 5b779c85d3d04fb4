"""Offline evaluation of a generations file: score every response with the rule-based scorers (or a custom reward function)
against the row's ground truth and report the mean score per data source (+ pass@k / maj@k style bootstrap when n > 1).

Counterpart of the reference's ``verl/trainer/main_eval.py`` (Ray tasks per row); scoring is CPU work, so rows are spread over
a process pool instead.

    python -m polyrl_b200.trainer.main_eval --data gen.parquet [--response-key responses] [--data-source-key data_source]
        [--reward-model-key reward_model] [--custom-reward path.py:compute_score] [--workers 16]
"""
from __future__ import annotations

import argparse
import json
import sys
from collections import defaultdict
from concurrent.futures import ProcessPoolExecutor

import numpy as np

_FNS: dict = {}          # per worker process: reward function per --custom-reward spec (None = the built-in scorers)


def _score_row(args) -> tuple[str, list[float]]:
    data_source, responses, reward_model, extra, custom = args
    if custom not in _FNS:
        if custom:
            from ..data.reward_manager import get_custom_reward_fn

            path, _, name = custom.partition(":")
            _FNS[custom] = get_custom_reward_fn(path, name or "compute_score")
        else:
            from ..data.reward_score import default_compute_score

            _FNS[custom] = default_compute_score
    _FN = _FNS[custom]
    gt = reward_model["ground_truth"] if isinstance(reward_model, dict) else reward_model
    if isinstance(responses, str):
        responses = [responses]
    scores = []
    for r in responses:
        s = _FN(data_source, r, gt, extra)
        scores.append(float(s["score"] if isinstance(s, dict) else s))
    return data_source, scores


def evaluate(rows: list[dict], response_key: str = "responses", data_source_key: str = "data_source",
             reward_model_key: str = "reward_model", custom_reward: str | None = None, workers: int = 8) -> dict[str, float]:
    tasks = [(r.get(data_source_key, "unknown"), r[response_key], r[reward_model_key], r.get("extra_info"), custom_reward) for r in rows]
    per_source: dict[str, list[list[float]]] = defaultdict(list)
    if workers <= 1 or len(tasks) < 2:
        results = map(_score_row, tasks)
    else:
        ex = ProcessPoolExecutor(max_workers=workers)
        results = ex.map(_score_row, tasks, chunksize=max(1, len(tasks) // (4 * workers)))
    for src, scores in results:
        per_source[src].append(scores)
    out: dict[str, float] = {}
    for src, lst in per_source.items():
        out[f"test_score/{src}"] = float(np.mean([np.mean(s) for s in lst]))
        n = min(len(s) for s in lst)
        if n > 1:                                   # any-correct over the n samples of a prompt, and the best sample's score
            out[f"test_score/{src}/best@{n}"] = float(np.mean([np.max(s[:n]) for s in lst]))
            out[f"test_score/{src}/worst@{n}"] = float(np.mean([np.min(s[:n]) for s in lst]))
        out[f"num_prompts/{src}"] = float(len(lst))
    return out


def main(argv: list[str] | None = None) -> int:
    from .main_generation import _translate_dotted

    argv = list(sys.argv[1:] if argv is None else argv)
    fn_name = next((a.split("=", 1)[1] for a in argv if a.startswith("custom_reward_function.name=")), None)
    argv = _translate_dotted(argv, {"data.path": "--data", "data.response_key": "--response-key", "data.data_source_key": "--data-source-key",
                                    "data.reward_model_key": "--reward-model-key", "custom_reward_function.path": "--custom-reward"})
    if fn_name and "--custom-reward" in argv:
        i = argv.index("--custom-reward") + 1
        argv[i] = f"{argv[i]}:{fn_name}"
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--data", required=True)
    ap.add_argument("--response-key", default="responses")
    ap.add_argument("--data-source-key", default="data_source")
    ap.add_argument("--reward-model-key", default="reward_model")
    ap.add_argument("--custom-reward", default=None, help="path/to/file.py[:function]")
    ap.add_argument("--workers", type=int, default=8)
    a = ap.parse_args(argv)
    from ..data.rl_dataset import _read_rows

    metrics = evaluate(_read_rows(a.data), a.response_key, a.data_source_key, a.reward_model_key, a.custom_reward, a.workers)
    print(json.dumps(metrics, indent=1, sort_keys=True))
    return 0


if __name__ == "__main__":
    sys.exit(main())
