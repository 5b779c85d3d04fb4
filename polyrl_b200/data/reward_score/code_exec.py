"""Test-case based code reward (PRIME / sandbox-fusion style): extract the last fenced Python block of the solution, run it in the
local sandbox against every test case and score the pass fraction (``continuous=False``: all-or-nothing).

Test-case formats (``ground_truth`` as dict or JSON string):
  * ``{"inputs": [stdin...], "outputs": [expected stdout...]}`` — stdin / stdout programs;
  * ``{"fn_name": "f", "inputs": [[args...]...], "outputs": [expected...]}`` — function-call style;
  * ``{"assert_case": ["assert f(2) == 4", ...]}`` — assertion snippets appended to the solution.
Reference capability: 3rdparty/verl/verl/utils/reward_score/prime_code/__init__.py, sandbox_fusion/__init__.py."""
from __future__ import annotations

import json
import re

from ...utils.sandbox import run_python

_FENCE = re.compile(r"```(?:python|py)?\s*\n(.*?)```", re.DOTALL)


def extract_code(solution: str) -> str | None:
    blocks = _FENCE.findall(solution)
    return blocks[-1].strip() if blocks else None


def _same(a: str, b: str) -> bool:
    na, nb = a.strip().split(), b.strip().split()
    if na == nb:
        return True
    try:
        return len(na) == len(nb) and all(abs(float(x) - float(y)) <= 1e-6 * max(1.0, abs(float(y))) for x, y in zip(na, nb))
    except ValueError:
        return False


class _Remote:
    """Runner with ``run_python``'s result surface over a sandbox-fusion ``/run_code`` endpoint (agent/remote_tools.py)."""

    def __init__(self, url: str):
        self.url = url

    def __call__(self, code: str, stdin: str = "", timeout_s: float = 5.0):
        from types import SimpleNamespace

        from ...agent.remote_tools import sandbox_fusion_run

        r = sandbox_fusion_run(self.url, code, stdin=stdin, run_timeout=timeout_s, timeout=timeout_s + 20)
        return SimpleNamespace(status="ok" if r["ok"] else "error", stdout=r["stdout"], stderr=r["stderr"])


def compute_score(solution_str: str, ground_truth, continuous: bool = True, timeout_s: float = 5.0, max_cases: int = 16,
                  sandbox_fusion_url: str | None = None, **_) -> tuple[float, dict]:
    """``sandbox_fusion_url`` (or env ``SANDBOX_FUSION_URL``) sends every run to a remote sandbox-fusion service instead of the
    local sandbox (reference: verl/utils/reward_score/sandbox_fusion/__init__.py)."""
    import os

    url = sandbox_fusion_url or os.environ.get("SANDBOX_FUSION_URL")
    run_python = _Remote(url) if url else globals()["run_python"]
    code = extract_code(solution_str)
    if code is None:
        return 0.0, {"error": "no code block"}
    tc = json.loads(ground_truth) if isinstance(ground_truth, str) else dict(ground_truth)
    results = []
    if "assert_case" in tc:
        for a in tc["assert_case"][:max_cases]:
            results.append(run_python(code + "\n\n" + a + "\n", timeout_s=timeout_s).status == "ok")
    elif tc.get("fn_name"):
        for args, want in list(zip(tc["inputs"], tc["outputs"]))[:max_cases]:
            harness = (code + "\n\nimport json as _j\n_args = _j.loads(" + repr(json.dumps(args)) + ")\n"
                       f"_out = {tc['fn_name']}(*_args)\nprint(_j.dumps(_out))\n")
            r = run_python(harness, timeout_s=timeout_s)
            ok = False
            if r.status == "ok" and r.stdout.strip():
                try:
                    got = json.loads(r.stdout.strip().splitlines()[-1])
                    ok = got == want or (isinstance(want, list) and len(want) == 1 and got == want[0])
                except json.JSONDecodeError:
                    ok = False
            results.append(ok)
    else:
        for inp, want in list(zip(tc["inputs"], tc["outputs"]))[:max_cases]:
            r = run_python(code, stdin=inp if isinstance(inp, str) else "\n".join(map(str, inp)), timeout_s=timeout_s)
            results.append(r.status == "ok" and _same(r.stdout, want if isinstance(want, str) else "\n".join(map(str, want))))
    if not results:
        return 0.0, {"error": "no test cases"}
    frac = sum(results) / len(results)
    return (frac if continuous else float(frac == 1.0)), {"passed": sum(results), "total": len(results)}
