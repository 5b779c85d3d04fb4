"""torchrun launcher for tests: one retry on a different rendezvous port when the child exits non-zero (a busy port or a slow
TCPStore on a loaded CI box must not turn into a red suite); value assertions are NOT retried."""
import subprocess


def run_torchrun(cmd: list[str], **kw) -> subprocess.CompletedProcess:
    r = subprocess.run(cmd, **kw)
    if r.returncode == 0 or "--master-port" not in cmd:
        return r
    i = cmd.index("--master-port")
    retry = list(cmd)
    retry[i + 1] = str(int(cmd[i + 1]) + 1000)
    return subprocess.run(retry, **kw)
