"""Geometry3K reward: 0.9 * accuracy(\\boxed answer) + 0.1 * format(<think>...</think> ... \\boxed{})
(reference capability: 3rdparty/verl/verl/utils/reward_score/geo3k.py)."""
import re

from .math import is_equiv, last_boxed_only_string, remove_boxed, strip_string


def format_reward(s: str) -> float:
    return 1.0 if re.fullmatch(r"<think>.*</think>.*\\boxed\{.*\}.*", s, re.DOTALL) else 0.0


def acc_reward(s: str, ground_truth: str) -> float:
    b = last_boxed_only_string(s)
    if b is None:
        return 0.0
    ans = remove_boxed(b)
    return 1.0 if ans is not None and is_equiv(strip_string(ans), strip_string(str(ground_truth))) else 0.0


def compute_score(solution_str: str, ground_truth: str, **_) -> float:
    return 0.9 * acc_reward(solution_str, ground_truth) + 0.1 * format_reward(solution_str)
