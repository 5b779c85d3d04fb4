"""Search-R1 style QA exact match: the last ``<answer>...</answer>`` span, normalised (lower-case, no punctuation / articles),
must equal one of the golden answers (reference capability: 3rdparty/verl/verl/utils/reward_score/search_r1_like_qa_em.py)."""
import re
import string


def normalize_answer(s: str) -> str:
    s = s.lower()
    s = "".join(ch for ch in s if ch not in set(string.punctuation))
    s = re.sub(r"\b(a|an|the)\b", " ", s)
    return " ".join(s.split())


def extract_solution(solution_str: str) -> str | None:
    m = re.findall(r"<answer>(.*?)</answer>", solution_str, re.DOTALL)
    return m[-1].strip() if m else None


def compute_score(solution_str: str, ground_truth, format_score: float = 0.0, score: float = 1.0, **_) -> float:
    ans = extract_solution(solution_str)
    if ans is None:
        return 0.0
    gold = ground_truth.get("target", ground_truth) if isinstance(ground_truth, dict) else ground_truth
    gold = [gold] if isinstance(gold, str) else list(gold)
    return score if any(normalize_answer(ans) == normalize_answer(str(g)) for g in gold) else format_score
