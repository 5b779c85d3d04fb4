"""GSM8K rule-based reward: take the last ``#### <number>`` in (the tail of) the solution; exact string
match with the ground truth after removing ',' and '$' (reference behaviour:
3rdparty/verl/verl/utils/reward_score/gsm8k.py:20-72)."""
import re

_STRICT = re.compile(r"#### (\-?[0-9\.\,]+)")
_FLEX = re.compile(r"(\-?[0-9\.\,]+)")
_CLIP = 300


def extract_solution(solution_str: str, method: str = "strict"):
    if len(solution_str) > _CLIP:
        solution_str = solution_str[-_CLIP:]
    if method == "strict":
        found = _STRICT.findall(solution_str)
        return found[-1].replace(",", "").replace("$", "") if found else None
    if method == "flexible":
        for cand in reversed(_FLEX.findall(solution_str)):
            if cand not in ("", "."):
                return cand.replace(",", "")
        return None
    raise ValueError(method)


def compute_score(solution_str: str, ground_truth: str, method: str = "strict", format_score: float = 0.0, score: float = 1.0) -> float:
    ans = extract_solution(solution_str, method)
    if ans is None:
        return 0.0
    return score if ans == str(ground_truth) else format_score
