"""Offline generation THROUGH stand-alone rollout servers: start one engine server per replica, send every prompt of a dataset as
OpenAI chat-completion requests (``n_samples`` choices each), write the table back with a ``responses`` column.

Counterpart of the reference's ``verl/trainer/main_generation_server.py`` (stand-alone SGLang / vLLM replicas + one
``/v1/chat/completions`` call per sample).  Here the replicas are ``polyrl_b200.rollout.launch_server`` processes (one per GPU, or
``--tp-size`` GPUs each under torchrun) serving ``rollout/openai_api.py``; ``--server`` points at servers that already run.

    python -m polyrl_b200.trainer.main_generation_server --model /ckpts/qwen2.5-7b --data test.parquet --output gen.parquet \\
        --n-samples 4 --temperature 1.0 --response-length 1024 [--num-replicas 8] [--server host:port ...]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
import urllib.request
from concurrent.futures import ThreadPoolExecutor


def _post(endpoint: str, path: str, body: dict, timeout: float = 3600.0) -> dict:
    req = urllib.request.Request(f"http://{endpoint}{path}", data=json.dumps(body).encode(), headers={"Content-Type": "application/json"})
    with urllib.request.urlopen(req, timeout=timeout) as r:
        return json.loads(r.read())


def _healthy(endpoint: str) -> bool:
    try:
        with urllib.request.urlopen(f"http://{endpoint}/health", timeout=2) as r:
            return r.status == 200
    except OSError:
        return False


def start_servers(model: str, num_replicas: int, base_port: int, tokenizer: str | None, max_model_len: int, max_num_seqs: int,
                  load_format: str = "auto", device: str | None = None, wait_s: float = 900.0) -> tuple[list[subprocess.Popen], list[str]]:
    """One ``launch_server`` process per replica (replica i sees GPU i); returns (processes, endpoints) once all are healthy."""
    procs, eps = [], []
    for i in range(num_replicas):
        env = dict(os.environ)
        if device != "cpu":
            env["CUDA_VISIBLE_DEVICES"] = str(i)
        cmd = [sys.executable, "-m", "polyrl_b200.rollout.launch_server", "--model-path", model, "--port", str(base_port + i),
               "--max-model-len", str(max_model_len), "--max-num-seqs", str(max_num_seqs), "--load-format", load_format]
        if tokenizer:
            cmd += ["--tokenizer-path", tokenizer]
        if device:
            cmd += ["--device", device]
        procs.append(subprocess.Popen(cmd, env=env))
        eps.append(f"127.0.0.1:{base_port + i}")
    t0 = time.time()
    while not all(_healthy(e) for e in eps):
        if any(p.poll() is not None for p in procs) or time.time() - t0 > wait_s:
            stop_servers(procs, eps)
            raise RuntimeError("a rollout server exited or did not become healthy in time")
        time.sleep(0.5)
    return procs, eps


def stop_servers(procs: list[subprocess.Popen], endpoints: list[str]) -> None:
    for e in endpoints:
        try:
            _post(e, "/shutdown", {}, timeout=5)
        except OSError:
            pass
    for p in procs:
        try:
            p.wait(timeout=30)
        except subprocess.TimeoutExpired:
            p.kill()


def generate(endpoints: list[str], chats: list[list[dict]], model: str, n_samples: int, sampling: dict, concurrency: int = 64) -> list[list[str]]:
    """responses[i] = the n_samples texts of chat i; chats are split contiguously over the replicas like the reference does."""
    if sampling.get("temperature") == 0.0 and n_samples != 1:
        raise ValueError("temperature 0 needs n_samples == 1")
    per = (len(chats) + len(endpoints) - 1) // max(1, len(endpoints))

    def one(i: int) -> list[str]:
        out = _post(endpoints[min(i // max(per, 1), len(endpoints) - 1)], "/v1/chat/completions",
                    {"model": model, "messages": chats[i], "n": n_samples, **sampling})
        return [c["message"]["content"] or "" for c in sorted(out["choices"], key=lambda c: c["index"])]

    with ThreadPoolExecutor(max_workers=max(1, concurrency)) as ex:
        return list(ex.map(one, range(len(chats))))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model", required=True)
    ap.add_argument("--data", required=True, nargs="+")
    ap.add_argument("--output", required=True)
    ap.add_argument("--prompt-key", default="prompt")
    ap.add_argument("--n-samples", type=int, default=1)
    ap.add_argument("--temperature", type=float, default=1.0)
    ap.add_argument("--top-p", type=float, default=1.0)
    ap.add_argument("--response-length", type=int, default=512)
    ap.add_argument("--prompt-length", type=int, default=1024)
    ap.add_argument("--num-replicas", type=int, default=0, help="servers to start (default: one per visible GPU, 1 on CPU)")
    ap.add_argument("--server", action="append", default=[], help="host:port of a server that already runs (repeatable); none are started")
    ap.add_argument("--base-port", type=int, default=31000)
    ap.add_argument("--tokenizer", default=None)
    ap.add_argument("--max-num-seqs", type=int, default=256)
    ap.add_argument("--load-format", default="auto")
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    from ..data.rl_dataset import _read_rows
    from .main_generation import _write

    rows = _read_rows(a.data)
    chats = [list(r[a.prompt_key]) if not isinstance(r[a.prompt_key], str) else [{"role": "user", "content": r[a.prompt_key]}] for r in rows]
    procs: list = []
    eps = list(a.server)
    if not eps:
        import torch

        n = a.num_replicas or (torch.cuda.device_count() if (a.device or "cuda") != "cpu" and torch.cuda.is_available() else 1)
        procs, eps = start_servers(a.model, n, a.base_port, a.tokenizer, a.prompt_length + a.response_length, a.max_num_seqs, a.load_format, a.device)
    try:
        res = generate(eps, chats, a.model, a.n_samples, {"temperature": a.temperature, "top_p": a.top_p, "max_tokens": a.response_length})
    finally:
        if procs:
            stop_servers(procs, eps)
    for r, texts in zip(rows, res):
        r["responses"] = texts
    _write(rows, a.output)
    print(f"[generation-server] {len(rows)} prompts x {a.n_samples} samples through {len(eps)} server(s) -> {a.output}", flush=True)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
