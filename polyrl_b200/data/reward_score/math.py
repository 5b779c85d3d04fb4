"""MATH-style reward: last \\boxed{...} answer, normalised string equivalence (reference capability:
3rdparty/verl/verl/utils/reward_score/math.py)."""
import re


def last_boxed_only_string(s: str):
    idx = max(s.rfind("\\boxed"), s.rfind("\\fbox"))
    if idx < 0:
        return None
    i, depth, right = idx, 0, None
    while i < len(s):
        if s[i] == "{":
            depth += 1
        elif s[i] == "}":
            depth -= 1
            if depth == 0:
                right = i
                break
        i += 1
    return s[idx:right + 1] if right is not None else None


def remove_boxed(s: str):
    if s.startswith("\\boxed "):
        return s[len("\\boxed "):]
    left = "\\boxed{"
    if s.startswith(left) and s.endswith("}"):
        return s[len(left):-1]
    m = re.match(r"\\fbox\{(.*)\}$", s)
    return m.group(1) if m else None


def _fix_fracs(s: str) -> str:
    return re.sub(r"\\frac(\d)(\d)", r"\\frac{\1}{\2}", s)


def strip_string(s: str) -> str:
    s = str(s).replace("\n", "").replace("\\!", "").replace("\\\\", "\\")
    s = s.replace("tfrac", "frac").replace("dfrac", "frac").replace("\\left", "").replace("\\right", "")
    s = s.replace("^{\\circ}", "").replace("^\\circ", "").replace("\\$", "").replace("$", "")
    s = re.sub(r"\\text\{\s*([^}]*)\}", r"\1", s)
    s = s.replace("\\%", "").replace("%", "")
    s = s.replace(" .", " 0.").replace("{.", "{0.")
    if s.startswith("."):
        s = "0" + s
    if len(s.split("=")) == 2 and len(s.split("=")[0]) <= 2:
        s = s.split("=")[1]
    s = _fix_fracs(s.replace(" ", ""))
    if s == "0.5":
        s = "\\frac{1}{2}"
    m = re.fullmatch(r"(\d+)/(\d+)", s)
    if m:
        s = f"\\frac{{{m.group(1)}}}{{{m.group(2)}}}"
    return s.rstrip(".")


def is_equiv(a, b) -> bool:
    if a is None or b is None:
        return a is None and b is None
    try:
        return strip_string(a) == strip_string(b)
    except Exception:
        return a == b


def compute_score(solution_str: str, ground_truth: str) -> float:
    boxed = last_boxed_only_string(solution_str)
    if boxed is None:
        return 0.0
    ans = remove_boxed(boxed)
    return 1.0 if is_equiv(ans, ground_truth) else 0.0
