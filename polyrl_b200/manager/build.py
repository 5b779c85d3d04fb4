"""Build the native C++ rollout manager binary (polyrl_b200/manager/bin/rollout-manager)."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "src"
BIN = HERE / "bin" / "rollout-manager"


def build_manager(force: bool = False, sanitize: str | None = None) -> Path | None:
    """``sanitize`` in {None, "thread", "address", "undefined"}: instrumented builds land next to the release binary as
    ``rollout-manager-<sanitizer>`` (the reference's Rust manager has no race tooling; tests/test_manager.py runs
    the scheduler / continuation / fault-injection scenarios under ThreadSanitizer)."""
    srcs = sorted(SRC.glob("*.cpp"))
    if not srcs:
        return None
    hdrs = sorted(SRC.glob("*.h*"))
    out = BIN if not sanitize else BIN.with_name(f"{BIN.name}-{sanitize}")
    out.parent.mkdir(parents=True, exist_ok=True)
    newest = max(p.stat().st_mtime for p in srcs + hdrs)
    if out.exists() and out.stat().st_mtime >= newest and not force:
        return out
    flags = ["-O2"] if not sanitize else ["-O1", "-g", f"-fsanitize={sanitize}", "-fno-omit-frame-pointer"]
    cmd = ["g++", *flags, "-std=c++17", "-pthread", "-I", str(SRC), *map(str, srcs), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("rollout-manager build failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    import sys

    print(build_manager(force=True, sanitize=(sys.argv[1] if len(sys.argv) > 1 else None)))
