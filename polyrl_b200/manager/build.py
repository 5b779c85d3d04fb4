"""Build the native C++ rollout manager binary (polyrl_b200/manager/bin/rollout-manager)."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "src"
BIN = HERE / "bin" / "rollout-manager"


def build_manager(force: bool = False) -> Path | None:
    srcs = sorted(SRC.glob("*.cpp"))
    if not srcs:
        return None
    hdrs = sorted(SRC.glob("*.h*"))
    BIN.parent.mkdir(parents=True, exist_ok=True)
    newest = max(p.stat().st_mtime for p in srcs + hdrs)
    if BIN.exists() and BIN.stat().st_mtime >= newest and not force:
        return BIN
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-I", str(SRC), *map(str, srcs), "-o", str(BIN)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("rollout-manager build failed:\n" + r.stdout + r.stderr)
    return BIN


if __name__ == "__main__":
    print(build_manager(force=True))
