"""DAPO-style math reward: the final ``Answer: ...`` line (or the last \\boxed{}) is normalised and compared; the score is
+1 / -1 (reference capability: 3rdparty/verl/verl/utils/reward_score/math_dapo.py)."""
import re

from .math import last_boxed_only_string, remove_boxed

_SUBS = [("an ", ""), ("a ", ""), (".$", "$"), ("\\$", ""), (r"\ ", ""), (" ", ""), ("mbox", "text"), (",\\text{and}", ","),
         ("\\text{and}", ","), ("\\text{m}", "\\text{}")]
_REMOVED = ["square", "ways", "integers", "dollars", "mph", "inches", "hours", "km", "units", "\\ldots", "sue", "points", "feet",
            "minutes", "digits", "cents", "degrees", "cm", "gm", "pounds", "meters", "meals", "edges", "students",
            "childrentickets", "multiples", "\\text{s}", "\\text{.}", "\\text{\ns}", "\\text{}^2", "\\text{}^3", "\\text{\n}",
            "\\text{}", r"\mathrm{th}", r"^\circ", r"^{\circ}", r"\;", r",\!", "{,}", '"', "\\dots"]


def normalize_final_answer(ans: str) -> str:
    ans = ans.split("=")[-1]
    for a, b in _SUBS:
        ans = ans.replace(a, b)
    for r in _REMOVED:
        ans = ans.replace(r, "")
    ans = re.sub(r"(.*?)(\$)(.*?)(\$)(.*)", "$\\3$", ans)
    for cmd in ("text", "textbf", "overline", "boxed"):
        ans = re.sub(r"(\\%s\{)(.*?)(\})" % cmd, "\\2", ans)
    ans = re.sub(r"(frac)([^{])(.)", "frac{\\2}{\\3}", ans)
    ans = re.sub(r"(sqrt)([^{])", "sqrt{\\2}", ans)
    ans = ans.replace("$", "")
    if ans.replace(",", "").isdigit():
        ans = ans.replace(",", "")
    return ans.strip()


def extract_answer(solution: str, strict_box: bool = False) -> str | None:
    if strict_box:
        b = last_boxed_only_string(solution[-200:])
        return remove_boxed(b) if b else None
    m = re.findall(r"(?i)Answer\s*:\s*([^\n]+)", solution)
    return m[-1] if m else None


def compute_score(solution_str: str, ground_truth: str, strict_box_verify: bool = False, **_) -> dict:
    tail = solution_str[-300:]
    pred = extract_answer(tail, strict_box_verify)
    if pred is None:
        return {"score": -1.0, "acc": False, "pred": "[INVALID]"}
    if strict_box_verify:
        ok = pred == str(ground_truth)
    else:
        ok = normalize_final_answer(pred) == normalize_final_answer(str(ground_truth))
    return {"score": 1.0 if ok else -1.0, "acc": bool(ok), "pred": pred}
