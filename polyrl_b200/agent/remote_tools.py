"""Tools backed by remote services and image tools: MCP servers (JSON-RPC 2.0 over HTTP), a sandbox-fusion style code execution
service, Geometry3K answer checking, image zoom-in.  Capability parity with the reference's verl/tools/{mcp_base_tool,
mcp_search_tool,sandbox_fusion_tools,geo3k_tool,image_zoom_in_tool}.py and tools/utils/mcp_clients/.  Everything speaks plain
HTTP through the standard library (there is no network in CI: the tests run tiny local servers)."""
from __future__ import annotations

import http.client
import itertools
import json
import threading
import time
from urllib.parse import urlparse

from .tools import BaseTool, ToolResponse, ToolSchema, register_tool


def _post_json(url: str, body: dict, timeout: float = 30.0, headers: dict | None = None) -> tuple[int, dict | str]:
    u = urlparse(url)
    conn = (http.client.HTTPSConnection if u.scheme == "https" else http.client.HTTPConnection)(u.hostname, u.port, timeout=timeout)
    try:
        conn.request("POST", u.path or "/", body=json.dumps(body),
                     headers={"Content-Type": "application/json", "Accept": "application/json, text/event-stream", **(headers or {})})
        r = conn.getresponse()
        data = r.read().decode()
        if "text/event-stream" in (r.getheader("Content-Type") or ""):        # streamable-HTTP servers may answer as one SSE event
            for line in data.splitlines():
                if line.startswith("data:"):
                    data = line[5:].strip()
        try:
            return r.status, json.loads(data)
        except json.JSONDecodeError:
            return r.status, data
    finally:
        conn.close()


class TokenBucket:
    """Client-side rate limit shared by the instances of a tool (reference: mcp_clients TokenBucket / sandbox rate_limit)."""

    def __init__(self, rate_per_s: float, burst: int | None = None):
        self.rate, self.cap = float(rate_per_s), float(burst or max(1, int(rate_per_s)))
        self.tokens, self.t = self.cap, time.monotonic()
        self.lock = threading.Lock()

    def acquire(self) -> None:
        while True:
            with self.lock:
                now = time.monotonic()
                self.tokens = min(self.cap, self.tokens + (now - self.t) * self.rate)
                self.t = now
                if self.tokens >= 1:
                    self.tokens -= 1
                    return
                wait = (1 - self.tokens) / self.rate
            time.sleep(wait)


class MCPClient:
    """Minimal Model Context Protocol client: ``initialize`` once, then ``tools/list`` and ``tools/call``."""

    def __init__(self, url: str, timeout: float = 30.0, headers: dict | None = None):
        self.url, self.timeout, self.headers = url, timeout, dict(headers or {})
        self._ids = itertools.count(1)
        self._ready = False
        self._lock = threading.Lock()

    def _rpc(self, method: str, params: dict | None = None) -> dict:
        status, out = _post_json(self.url, {"jsonrpc": "2.0", "id": next(self._ids), "method": method, "params": params or {}},
                                 self.timeout, self.headers)
        if status != 200 or not isinstance(out, dict):
            raise RuntimeError(f"MCP server {self.url}: HTTP {status}: {str(out)[:200]}")
        if "error" in out:
            raise RuntimeError(f"MCP error {out['error'].get('code')}: {out['error'].get('message')}")
        return out.get("result", {})

    def _ensure(self) -> None:
        with self._lock:
            if not self._ready:
                self._rpc("initialize", {"protocolVersion": "2025-03-26", "capabilities": {},
                                         "clientInfo": {"name": "polyrl_b200", "version": "0"}})
                self._ready = True

    def list_tools(self) -> list[dict]:
        self._ensure()
        return self._rpc("tools/list").get("tools", [])

    def call_tool(self, name: str, arguments: dict) -> tuple[str, bool]:
        self._ensure()
        res = self._rpc("tools/call", {"name": name, "arguments": arguments})
        text = "\n".join(c.get("text", "") for c in res.get("content", []) if c.get("type") == "text")
        return text, bool(res.get("isError", False))


@register_tool("mcp")
class MCPTool(BaseTool):
    """One tool of an MCP server.  config: ``{"url": ..., "tool_name": ... (default: the schema name), "rate_limit": calls/s,
    "timeout": s, "headers": {...}}``.  ``MCPTool.discover(url)`` builds one instance per tool the server lists."""

    def __init__(self, config=None, schema=None, client: MCPClient | None = None):
        super().__init__(config, schema)
        c = self.config
        self.client = client or MCPClient(c["url"], c.get("timeout", 30.0), c.get("headers"))
        self.remote_name = c.get("tool_name") or self.name
        self.bucket = TokenBucket(c["rate_limit"]) if c.get("rate_limit") else None

    @classmethod
    def discover(cls, url: str, **config) -> list["MCPTool"]:
        client = MCPClient(url, config.get("timeout", 30.0), config.get("headers"))
        out = []
        for t in client.list_tools():
            schema = ToolSchema(name=t["name"], description=t.get("description", ""),
                                parameters=t.get("inputSchema") or {"type": "object", "properties": {}, "required": []})
            out.append(cls({"url": url, **config}, schema, client))
        return out

    def _render(self, text: str) -> str:
        return text

    def execute(self, instance_id, parameters, **kwargs):
        if self.bucket is not None:
            self.bucket.acquire()
        try:
            text, is_err = self.client.call_tool(self.remote_name, parameters)
        except Exception as e:                         # the model sees the failure and may retry
            return ToolResponse(text=f"tool call failed: {e}"), 0.0, {"mcp_error": 1}
        return ToolResponse(text=self._render(text)), 0.0, {"mcp_error": int(is_err)}


@register_tool("mcp_search")
class MCPSearchTool(MCPTool):
    """Search servers answer with JSON ``{"results": [{"title", "url", "content"}...]}``; shown to the model as a numbered list."""

    def _render(self, text: str) -> str:
        try:
            hits = json.loads(text).get("results", [])
        except (json.JSONDecodeError, AttributeError):
            return text
        return "\n".join(f"[{i + 1}] {h.get('title', '')} ({h.get('url', '')})\n{h.get('content', '')}" for i, h in enumerate(hits)) or "no results"


def sandbox_fusion_run(url: str, code: str, stdin: str = "", language: str = "python", run_timeout: float = 10.0,
                       memory_limit_mb: int = 1024, timeout: float = 30.0) -> dict:
    """POST ``{url}`` (a sandbox-fusion ``/run_code`` endpoint) -> ``{"ok", "stdout", "stderr", "status"}``."""
    status, out = _post_json(url, {"code": code, "stdin": stdin, "language": language, "run_timeout": run_timeout,
                                   "compile_timeout": run_timeout, "memory_limit_MB": memory_limit_mb}, timeout)
    if status != 200 or not isinstance(out, dict):
        return {"ok": False, "stdout": "", "stderr": f"sandbox HTTP {status}: {str(out)[:200]}", "status": "SandboxError"}
    run = out.get("run_result") or {}
    ok = out.get("status") == "Success" and run.get("return_code", 1) == 0
    return {"ok": ok, "stdout": run.get("stdout", "") or "", "stderr": run.get("stderr", "") or "", "status": out.get("status", "")}


@register_tool("sandbox_fusion")
class SandboxFusionTool(BaseTool):
    """``code_interpreter(code)`` executed by a remote sandbox-fusion service (config: ``sandbox_fusion_url``, ``rate_limit``,
    ``default_timeout``, ``default_language``, ``memory_limit_mb``).  The local-sandbox variant is ``code_interpreter``."""

    @classmethod
    def default_schema(cls) -> ToolSchema:
        return ToolSchema(name="code_interpreter", description="Execute code remotely and return what it prints.",
                          parameters={"type": "object", "required": ["code"],
                                      "properties": {"code": {"type": "string", "description": "source code"}}})

    def __init__(self, config=None, schema=None):
        super().__init__(config, schema)
        self.url = self.config["sandbox_fusion_url"]
        self.bucket = TokenBucket(self.config["rate_limit"]) if self.config.get("rate_limit") else None

    def execute(self, instance_id, parameters, **kwargs):
        if self.bucket is not None:
            self.bucket.acquire()
        c = self.config
        res = sandbox_fusion_run(self.url, str(parameters.get("code", "")), language=parameters.get("language", c.get("default_language", "python")),
                                 run_timeout=float(parameters.get("timeout", c.get("default_timeout", 10))),
                                 memory_limit_mb=int(c.get("memory_limit_mb", 1024)))
        text = res["stdout"] if res["ok"] else (res["stdout"] + ("\n" if res["stdout"] else "") + (res["stderr"] or res["status"]))
        return ToolResponse(text=text or "(no output)"), 0.0, {"sandbox_ok": int(res["ok"])}


@register_tool("calc_geo3k_reward")
class Geo3kTool(BaseTool):
    """``calc_geo3k_reward(answer)``: checks a boxed answer against the ground truth with the geo3k scorer."""

    @classmethod
    def default_schema(cls) -> ToolSchema:
        return ToolSchema(name="calc_geo3k_reward", description="Check a candidate final answer for the geometry problem.",
                          parameters={"type": "object", "required": ["answer"],
                                      "properties": {"answer": {"type": "string", "description": "the final answer"}}})

    def create(self, instance_id=None, ground_truth: str | None = None, **kwargs):
        iid, resp = super().create(instance_id, **kwargs)
        self._instances[iid] = {"ground_truth": ground_truth, "response": "", "reward": 0.0}
        return iid, resp

    def execute(self, instance_id, parameters, **kwargs):
        st = self._instances[instance_id]
        ans = str(parameters.get("answer", "")).strip()
        st["response"] = ans if "\\boxed" in ans else f"\\boxed{{{ans}}}"
        reward = self.calc_reward(instance_id)
        step = -0.05 if reward <= st["reward"] else 0.0
        st["reward"] = max(st["reward"], reward)
        return ToolResponse(text=f"Current parsed answer={ans!r} reward={reward}"), step, {}

    def calc_reward(self, instance_id, **kwargs) -> float:
        from ..data.reward_score import geo3k

        st = self._instances[instance_id]
        return float(geo3k.acc_reward(st["response"], st["ground_truth"]) if hasattr(geo3k, "acc_reward")
                     else geo3k.compute_score(st["response"], st["ground_truth"]))


@register_tool("image_zoom_in_tool")
class ImageZoomInTool(BaseTool):
    """``image_zoom_in_tool(bbox_2d, label?)``: crop the trajectory's image to ``[x1, y1, x2, y2]`` (absolute pixels), enlarged
    to at least ``min_size`` pixels per side; the crop comes back as an image payload for vision-language policies."""

    @classmethod
    def default_schema(cls) -> ToolSchema:
        return ToolSchema(name="image_zoom_in_tool", description="Zoom into a region of the image.",
                          parameters={"type": "object", "required": ["bbox_2d"],
                                      "properties": {"bbox_2d": {"type": "array", "items": {"type": "number"}, "description": "[x1, y1, x2, y2]"},
                                                     "label": {"type": "string"}}})

    def create(self, instance_id=None, image=None, **kwargs):
        from ..data.rl_dataset import load_image

        iid, resp = super().create(instance_id, **kwargs)
        self._instances[iid] = {"image": load_image(image) if image is not None else None}
        return iid, resp

    def execute(self, instance_id, parameters, **kwargs):
        img = self._instances[instance_id].get("image")
        box = parameters.get("bbox_2d")
        if img is None:
            return ToolResponse(text="no image is attached to this conversation"), 0.0, {"zoom_ok": 0}
        if not (isinstance(box, (list, tuple)) and len(box) == 4):
            return ToolResponse(text="bbox_2d must be [x1, y1, x2, y2]"), -0.05, {"zoom_ok": 0}
        x1, y1, x2, y2 = (float(v) for v in box)
        x1, x2 = sorted((max(0.0, min(x1, img.width)), max(0.0, min(x2, img.width))))
        y1, y2 = sorted((max(0.0, min(y1, img.height)), max(0.0, min(y2, img.height))))
        if x2 - x1 < 1 or y2 - y1 < 1:
            return ToolResponse(text="the box is empty after clipping to the image"), -0.05, {"zoom_ok": 0}
        crop = img.crop((int(x1), int(y1), int(round(x2)), int(round(y2))))
        m = int(self.config.get("min_size", 28))
        if min(crop.size) < m:
            s = m / min(crop.size)
            crop = crop.resize((max(m, int(crop.width * s)), max(m, int(crop.height * s))))
        label = parameters.get("label")
        return ToolResponse(text=f"Zoomed in on {label or 'the region'}.", image=crop), 0.0, {"zoom_ok": 1}
