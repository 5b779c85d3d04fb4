"""Build the native C++ rollout manager binary (polyrl_b200/manager/bin/rollout-manager).

Binaries are never committed: they are built from ``manager/src`` on first use and cached under ``manager/bin/`` (git-ignored)
next to a stamp holding the hash of the sources + flags that produced them, so a checkout with arbitrary mtimes can never run a
stale or foreign ELF."""
from __future__ import annotations

import hashlib
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "src"
BIN = HERE / "bin" / "rollout-manager"


def _digest(paths: list[Path], flags: list[str]) -> str:
    h = hashlib.sha256(" ".join(flags).encode())
    for p in paths:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


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
    flags = ["-O2"] if not sanitize else ["-O1", "-g", f"-fsanitize={sanitize}", "-fno-omit-frame-pointer"]
    flags += ["-std=c++17", "-pthread"]
    want = _digest(srcs + hdrs, flags)
    stamp = out.with_name(out.name + ".stamp")
    if out.exists() and stamp.exists() and stamp.read_text().strip() == want and not force:
        return out
    tmp = out.with_name(out.name + ".tmp")
    cmd = ["g++", *flags, "-I", str(SRC), *map(str, srcs), "-o", str(tmp)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("rollout-manager build failed:\n" + r.stdout + r.stderr)
    tmp.replace(out)
    stamp.write_text(want + "\n")
    return out


if __name__ == "__main__":
    import sys

    print(build_manager(force=True, sanitize=(sys.argv[1] if len(sys.argv) > 1 else None)))
