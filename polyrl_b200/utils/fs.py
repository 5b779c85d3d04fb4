"""Local / remote file helpers: stage model and data paths onto the node, mirror checkpoints to a second location
(reference: verl/utils/fs.py ``copy_to_local`` / ``is_non_local`` / md5-named cache, verl/utils/hdfs_io.py ``copy`` / ``makedirs`` /
``exists`` through the ``hdfs dfs`` command line).

Remote schemes go through the ``hdfs`` CLI when it is on PATH (``hdfs://``, ``viewfs://``); ``file://`` and plain paths are local.
Everything else (an NFS / Lustre mount used as ``trainer.default_hdfs_dir``) is an ordinary directory copy.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

_REMOTE = ("hdfs://", "viewfs://")


def is_non_local(path: str | os.PathLike) -> bool:
    return str(path).startswith(_REMOTE)


def _strip(path: str | os.PathLike) -> str:
    p = str(path)
    return p[len("file://"):] if p.startswith("file://") else p


def _hdfs(*args: str, check: bool = True) -> subprocess.CompletedProcess:
    exe = shutil.which("hdfs")
    if exe is None:
        raise RuntimeError(f"`hdfs` command not found but a remote path was given ({args[-1]!r})")
    r = subprocess.run([exe, "dfs", *args], capture_output=True, text=True)
    if check and r.returncode != 0:
        raise RuntimeError(f"hdfs dfs {' '.join(args)} failed: {r.stderr.strip()[-500:]}")
    return r


def exists(path: str | os.PathLike) -> bool:
    if is_non_local(path):
        return _hdfs("-test", "-e", str(path), check=False).returncode == 0
    return os.path.exists(_strip(path))


def makedirs(path: str | os.PathLike) -> None:
    if is_non_local(path):
        _hdfs("-mkdir", "-p", str(path))
    else:
        os.makedirs(_strip(path), exist_ok=True)


def copy(src: str | os.PathLike, dst: str | os.PathLike) -> str:
    """File or directory ``src`` -> ``dst`` (either side may be remote).  A directory lands AS ``dst`` (not inside it)."""
    s, d = str(src), str(dst)
    if is_non_local(s) and is_non_local(d):
        _hdfs("-cp", "-f", s, d)
    elif is_non_local(s):
        os.makedirs(os.path.dirname(_strip(d)) or ".", exist_ok=True)
        _hdfs("-get", "-f", s, _strip(d))
    elif is_non_local(d):
        makedirs(os.path.dirname(d.rstrip("/")))
        _hdfs("-put", "-f", _strip(s), d)
    else:
        s, d = _strip(s), _strip(d)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True)
        else:
            os.makedirs(os.path.dirname(d) or ".", exist_ok=True)
            shutil.copy2(s, d)
    return d


def local_cache_path(src: str, cache_dir: str | None = None) -> str:
    """Stable node-local name of a remote path: <cache>/<md5 of the full path>/<basename>."""
    root = cache_dir or os.path.join(os.environ.get("POLYRL_CACHE_DIR", os.path.join(os.path.expanduser("~"), ".cache", "polyrl_b200")), "fs")
    return os.path.join(root, hashlib.md5(src.encode()).hexdigest(), os.path.basename(src.rstrip("/")))


def copy_to_local(src: str | os.PathLike | None, cache_dir: str | None = None, use_shm: bool = False, always_recopy: bool = False) -> str | None:
    """Make ``src`` readable from this node and return the local path.  Local paths are returned unchanged (``use_shm`` first copies
    them into /dev/shm so that many ranks read the weights from memory); remote ones are fetched once into the cache."""
    if src is None:
        return None
    s = str(src)
    if not is_non_local(s):
        s = _strip(s)
        if use_shm and os.path.exists(s):
            dst = os.path.join("/dev/shm", "polyrl_b200", hashlib.md5(os.path.abspath(s).encode()).hexdigest(), os.path.basename(s.rstrip("/")))
            if always_recopy or not os.path.exists(dst):
                copy(s, dst)
            return dst
        return s
    dst = local_cache_path(s, cache_dir)
    if always_recopy or not os.path.exists(dst):
        tmp = dst + ".partial"
        if os.path.exists(tmp):
            shutil.rmtree(tmp, ignore_errors=True) if os.path.isdir(tmp) else os.remove(tmp)
        copy(s, tmp)
        os.replace(tmp, dst)
    return dst


def mirror_checkpoint(local_dir: str | os.PathLike, remote_root: str | os.PathLike | None) -> str | None:
    """Copy one ``global_step_N`` directory under ``remote_root`` (``trainer.default_hdfs_dir``) and update its latest marker."""
    if not remote_root:
        return None
    local_dir = str(local_dir).rstrip("/")
    dst = str(remote_root).rstrip("/") + "/" + os.path.basename(local_dir)
    copy(local_dir, dst)
    marker = Path(local_dir).parent / "latest_checkpointed_iteration.txt"
    if marker.exists():
        copy(str(marker), str(remote_root).rstrip("/") + "/" + marker.name)
    return dst


def stage_path(path, use_shm: bool = False):
    """``copy_to_local`` for multi-rank jobs: one rank per node (LOCAL_RANK 0) fetches a remote path, the others wait at a barrier
    and resolve the same cache location.  Lists are staged element-wise; local paths pass through untouched."""
    if path is None:
        return None
    if isinstance(path, (list, tuple)):
        return type(path)(stage_path(p, use_shm) for p in path)
    if not is_non_local(path) and not use_shm:
        return _strip(path)
    import torch.distributed as dist

    first = int(os.environ.get("LOCAL_RANK", 0)) == 0
    out = copy_to_local(path, use_shm=use_shm) if first else None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    return out if first else copy_to_local(path, use_shm=use_shm)        # already present: resolves without copying
