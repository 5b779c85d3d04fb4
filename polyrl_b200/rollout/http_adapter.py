"""Client-side adapter for one engine server: launch (optional), health-gate, register with the manager as a LOCAL instance,
and typed wrappers over the engine's REST surface.

Capability parity: rlboost/verl_stream/workers/rollout/sglang_rollout/sglang_http_async_engine.py:43-299
(``HttpServerAsyncEngineAdapter``: subprocess launch, ``/health_generate`` polling, ``/register_local_rollout_instances``,
generate / update_weights / release+resume memory / abort / flush wrappers) and verl's ``HttpServerAdapter``
(workers/rollout/sglang_rollout/http_server_engine.py:194).  Sync and asyncio flavours share one implementation.
"""
from __future__ import annotations

import asyncio
import json
import subprocess
import sys
import time
import urllib.error
import urllib.request


class EngineHttpAdapter:
    def __init__(self, host: str = "127.0.0.1", port: int = 30000, timeout: float = 600.0, max_retries: int = 3,
                 retry_delay: float = 0.5):
        self.host, self.port, self.timeout = host, port, timeout
        self.max_retries, self.retry_delay = max_retries, retry_delay
        self.process: subprocess.Popen | None = None

    @property
    def endpoint(self) -> str:
        return f"{self.host}:{self.port}"

    # ------------------------------------------------------------------ lifecycle
    @classmethod
    def launch(cls, model_path: str, host: str = "127.0.0.1", port: int = 30000, extra_args: list[str] | None = None,
               wait_s: float = 300.0, **kw) -> "EngineHttpAdapter":
        """Start ``python -m polyrl_b200.rollout.launch_server`` as a subprocess and wait until it answers health checks."""
        self = cls(host, port, **kw)
        cmd = [sys.executable, "-m", "polyrl_b200.rollout.launch_server", "--model-path", model_path, "--host", host, "--port",
               str(port), *(extra_args or [])]
        self.process = subprocess.Popen(cmd)
        self.wait_healthy(wait_s)
        return self

    def wait_healthy(self, timeout: float = 300.0, interval: float = 0.5) -> None:
        t0 = time.time()
        while time.time() - t0 < timeout:
            if self.process is not None and self.process.poll() is not None:
                raise RuntimeError(f"engine server exited with code {self.process.returncode} during start-up")
            try:
                if self._request("GET", "/health_generate", None, retries=1).get("http_status", 200) == 200:
                    return
            except Exception:
                pass
            time.sleep(interval)
        raise TimeoutError(f"engine at {self.endpoint} not healthy after {timeout} s")

    def register_local(self, manager_endpoint: str) -> dict:
        """Tell the rollout manager this engine sits on the trainer's GPUs (time-boxed local window applies)."""
        req = urllib.request.Request(f"http://{manager_endpoint}/register_local_rollout_instances",
                                     data=json.dumps({"instances": [{"host": self.host, "port": self.port}]}).encode(),
                                     headers={"Content-Type": "application/json"}, method="POST")
        with urllib.request.urlopen(req, timeout=30) as r:
            return json.loads(r.read().decode() or "{}")

    def shutdown(self, graceful: bool = True) -> None:
        try:
            self._request("POST", f"/shutdown?graceful={'true' if graceful else 'false'}", {}, retries=1)
        except Exception:
            pass
        if self.process is not None:
            try:
                self.process.wait(timeout=10)
            except subprocess.TimeoutExpired:
                self.process.terminate()

    # ------------------------------------------------------------------ REST wrappers
    def _request(self, method: str, path: str, body, retries: int | None = None) -> dict:
        last = None
        for attempt in range(retries or self.max_retries):
            try:
                data = None if body is None else json.dumps(body).encode()
                req = urllib.request.Request(f"http://{self.endpoint}{path}", data=data, method=method,
                                             headers={"Content-Type": "application/json"})
                with urllib.request.urlopen(req, timeout=self.timeout) as r:
                    raw = r.read().decode()
                    out = json.loads(raw) if raw else {}
                    if isinstance(out, dict):
                        out.setdefault("http_status", r.status)
                    return out
            except (urllib.error.URLError, ConnectionError, TimeoutError) as e:
                last = e
                time.sleep(self.retry_delay * (attempt + 1))
        raise RuntimeError(f"{method} {path} on {self.endpoint} failed after retries: {last!r}")

    def generate(self, input_ids, sampling_params: dict | None = None, return_logprob: bool = True):
        return self._request("POST", "/generate", {"input_ids": input_ids, "sampling_params": sampling_params or {},
                                                   "return_logprob": return_logprob, "stream": False})

    def server_info(self) -> dict:
        return self._request("GET", "/get_server_info", None)

    def update_weights_from_agent(self, weight_version: int, tensors_meta=None, bootstrap: bool = False) -> dict:
        return self._request("POST", "/update_weights_from_agent", {"weight_version": weight_version, "bootstrap": bootstrap,
                                                                    "tensors_meta": tensors_meta or []})

    def update_weights_from_tensor(self, payload: dict | None = None) -> dict:
        return self._request("POST", "/update_weights_from_tensor", payload or {})

    def release_memory_occupation(self, tags=("kv_cache",)) -> dict:
        return self._request("POST", "/release_memory_occupation", {"tags": list(tags)})

    def resume_memory_occupation(self, tags=("kv_cache",)) -> dict:
        return self._request("POST", "/resume_memory_occupation", {"tags": list(tags)})

    def abort_request(self, rid: str = "", abort_all: bool = False) -> dict:
        return self._request("POST", "/abort_request", {"rid": rid, "abort_all": abort_all})

    def flush_cache(self) -> dict:
        return self._request("POST", "/flush_cache", {})


class AsyncEngineHttpAdapter(EngineHttpAdapter):
    """asyncio flavour: every wrapper is awaitable (requests run on the default executor)."""

    def __getattribute__(self, name):
        attr = super().__getattribute__(name)
        if name in {"generate", "server_info", "update_weights_from_agent", "update_weights_from_tensor", "release_memory_occupation",
                    "resume_memory_occupation", "abort_request", "flush_cache"}:
            async def call(*a, **k):
                return await asyncio.get_running_loop().run_in_executor(None, lambda: attr(*a, **k))
            return call
        return attr
