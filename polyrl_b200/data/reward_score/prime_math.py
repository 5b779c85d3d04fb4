"""PRIME-style math grading without sympy: boxed/last-number extraction, then string, fraction and numeric equivalence with
tolerance (reference capability: 3rdparty/verl/verl/utils/reward_score/prime_math/__init__.py)."""
import re
from fractions import Fraction

from .math import is_equiv, last_boxed_only_string, remove_boxed, strip_string


def _to_number(s: str):
    s = s.strip().replace(",", "").replace("$", "").rstrip("%")
    m = re.fullmatch(r"\\frac\{(-?\d+)\}\{(-?\d+)\}", s)
    if m and int(m.group(2)) != 0:
        return Fraction(int(m.group(1)), int(m.group(2)))
    m = re.fullmatch(r"(-?\d+)\s*/\s*(-?\d+)", s)
    if m and int(m.group(2)) != 0:
        return Fraction(int(m.group(1)), int(m.group(2)))
    try:
        return Fraction(s)
    except (ValueError, ZeroDivisionError):
        return None


def extract_answer(solution: str) -> str | None:
    b = last_boxed_only_string(solution)
    if b is not None:
        return remove_boxed(b)
    nums = re.findall(r"-?\d[\d,]*\.?\d*(?:/\d+)?", solution)
    return nums[-1] if nums else None


def math_equal(pred: str, ref: str, rel_tol: float = 1e-4) -> bool:
    if pred is None:
        return False
    a, b = strip_string(str(pred)), strip_string(str(ref))
    if is_equiv(a, b):
        return True
    na, nb = _to_number(a), _to_number(b)
    if na is not None and nb is not None:
        return abs(float(na) - float(nb)) <= rel_tol * max(1.0, abs(float(nb)))
    # tuples / intervals: compare element-wise
    if a[:1] in "([" and b[:1] in "([" and a[-1:] in ")]" and b[-1:] in ")]":
        pa, pb = a[1:-1].split(","), b[1:-1].split(",")
        return len(pa) == len(pb) and a[0] == b[0] and a[-1] == b[-1] and all(math_equal(x, y, rel_tol) for x, y in zip(pa, pb))
    return False


def compute_score(solution_str: str, ground_truth: str, **_) -> tuple[bool, bool, str]:
    """(is_correct, format_ok, extracted)"""
    pred = extract_answer(solution_str)
    return math_equal(pred, ground_truth), pred is not None, pred or ""
