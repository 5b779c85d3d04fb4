"""Browse the generation dumps the trainers write (``trainer.rollout_data_dir`` / ``trainer.validation_data_dir``: one JSONL per
step with input / output / score) in a terminal (reference: scripts/rollout_viewer.py, a textual TUI): filter by step, score range
or substring, sort by score, print per-step statistics.

    python -m polyrl_b200.tools.rollout_viewer outputs/rollouts [--step 12] [--min-score 0.5] [--grep "####"] [--top 5] [--stats]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path


def load(path: str) -> list[dict]:
    p = Path(path)
    files = sorted(p.glob("*.jsonl")) if p.is_dir() else [p]
    rows = []
    for f in files:
        for line in f.read_text().splitlines():
            if line.strip():
                r = json.loads(line)
                r.setdefault("_file", f.name)
                rows.append(r)
    return rows


def select(rows: list[dict], step: int | None = None, min_score: float | None = None, max_score: float | None = None,
           grep: str | None = None) -> list[dict]:
    out = []
    for r in rows:
        if step is not None and int(r.get("step", -1)) != step:
            continue
        s = float(r.get("score", 0.0))
        if (min_score is not None and s < min_score) or (max_score is not None and s > max_score):
            continue
        if grep and grep not in str(r.get("input", r.get("input_ids", ""))) + str(r.get("output", r.get("output_ids", ""))):
            continue
        out.append(r)
    return out


def stats(rows: list[dict]) -> dict[int, dict]:
    by: dict[int, list[float]] = {}
    for r in rows:
        by.setdefault(int(r.get("step", -1)), []).append(float(r.get("score", 0.0)))
    return {k: {"n": len(v), "mean": sum(v) / len(v), "min": min(v), "max": max(v)} for k, v in sorted(by.items())}


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("path")
    ap.add_argument("--step", type=int, default=None)
    ap.add_argument("--min-score", type=float, default=None)
    ap.add_argument("--max-score", type=float, default=None)
    ap.add_argument("--grep", default=None)
    ap.add_argument("--top", type=int, default=10, help="show this many rows, best score first (negative: worst first)")
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--width", type=int, default=400, help="characters of input / output to show")
    a = ap.parse_args(argv)
    rows = load(a.path)
    if a.stats:
        for step, s in stats(rows).items():
            print(f"step {step:>6}: n={s['n']:<5} mean={s['mean']:.4f} min={s['min']:.3f} max={s['max']:.3f}")
        return 0
    sel = sorted(select(rows, a.step, a.min_score, a.max_score, a.grep), key=lambda r: float(r.get("score", 0.0)), reverse=a.top >= 0)
    for r in sel[: abs(a.top)]:
        print("=" * 100)
        print(f"step {r.get('step')}  score {float(r.get('score', 0.0)):.4f}  source {r.get('data_source')}  [{r['_file']}]")
        print("--- input\n" + str(r.get("input", r.get("input_ids")))[-a.width:])
        print("--- output\n" + str(r.get("output", r.get("output_ids")))[: a.width])
    print(f"[{len(sel)} of {len(rows)} rows match]")
    return 0


if __name__ == "__main__":
    sys.exit(main())
