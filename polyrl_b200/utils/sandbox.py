"""Local sandboxed Python execution for code rewards and code-interpreter tools: each run is a fresh ``python -I`` subprocess
in a temporary directory with CPU-time / address-space / file-size / process-count limits and a wall-clock timeout.

Reference capability: the sandbox-fusion client and tools (3rdparty/verl/verl/utils/reward_score/sandbox_fusion/,
verl/tools/sandbox_fusion_tools.py) and PRIME's test-case code scorer (verl/utils/reward_score/prime_code/) — those call a
remote sandbox service / fork testing utilities; with no network here the sandbox is a local subprocess jail.  It is a
resource guard for model-written code, not a security boundary against a determined attacker.
"""
from __future__ import annotations

import os
import resource
import subprocess
import sys
import tempfile
from dataclasses import dataclass


@dataclass
class RunResult:
    status: str            # ok | timeout | runtime_error | limit
    stdout: str
    stderr: str
    returncode: int | None
    wall_s: float


def _limits(cpu_s: int, mem_mb: int, fsize_mb: int = 16, nproc: int = 64):
    def apply():
        os.setsid()
        resource.setrlimit(resource.RLIMIT_CPU, (cpu_s, cpu_s + 1))
        resource.setrlimit(resource.RLIMIT_AS, (mem_mb << 20, mem_mb << 20))
        resource.setrlimit(resource.RLIMIT_FSIZE, (fsize_mb << 20, fsize_mb << 20))
        resource.setrlimit(resource.RLIMIT_CORE, (0, 0))
        try:
            resource.setrlimit(resource.RLIMIT_NPROC, (nproc, nproc))
        except (ValueError, OSError):
            pass
    return apply


def run_python(code: str, stdin: str = "", timeout_s: float = 5.0, mem_mb: int = 1024, max_output: int = 1 << 16) -> RunResult:
    import time

    t0 = time.time()
    with tempfile.TemporaryDirectory(prefix="polyrl_sbx_") as d:
        path = os.path.join(d, "main.py")
        with open(path, "w") as f:
            f.write(code)
        env = {"PATH": os.environ.get("PATH", "/usr/bin:/bin"), "HOME": d, "PYTHONDONTWRITEBYTECODE": "1", "OMP_NUM_THREADS": "1"}
        try:
            p = subprocess.run([sys.executable, "-I", "-S", path], input=stdin, capture_output=True, text=True, cwd=d, env=env,
                               timeout=timeout_s, preexec_fn=_limits(int(timeout_s) + 1, mem_mb))
        except subprocess.TimeoutExpired as e:
            return RunResult("timeout", (e.stdout or "")[:max_output] if isinstance(e.stdout, str) else "", "timeout", None, time.time() - t0)
        status = "ok" if p.returncode == 0 else ("limit" if p.returncode in (-9, -24, -25, 137) else "runtime_error")
        return RunResult(status, p.stdout[:max_output], p.stderr[-4000:], p.returncode, time.time() - t0)
