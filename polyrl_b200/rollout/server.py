"""HTTP front-end of the native rollout engine, wire-compatible with the SGLang endpoints the rollout
manager and the reference trainer use (reference patches:
rlboost/sglang/patches.py:390-566, rlboost/verl_stream/workers/rollout/sglang_rollout/sglang_http_async_engine.py:155-299):

  POST /generate (stream => SSE ``data: {...}`` cumulative chunks, ``data: [DONE]``), GET /health_generate,
  GET /get_server_info (``#running_req`` / ``#queue_req`` / ``last_gen_throughput``), POST /abort_request,
  POST /shutdown, POST /update_weights_from_agent, POST /update_weights_from_tensor,
  POST /release_memory_occupation, /resume_memory_occupation, /flush_cache.

One background thread owns the engine (``step()`` loop); HTTP handler threads only enqueue requests and read
per-request event queues.  ``register_with_manager`` self-registers the instance (remote or local).
"""
from __future__ import annotations

import json
import os
import queue
import threading
import time
import urllib.request
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

from .engine import Request, RolloutEngine, SamplingParams


class _Tracked:
    def __init__(self, req: Request, index: int):
        self.req, self.index = req, index
        self.sent = 0


class EngineServer:
    def __init__(self, engine: RolloutEngine, host: str = "127.0.0.1", port: int = 0, stream_interval: int = 128,
                 weight_update_fn=None, tp_cmd_group=None, tokenizer=None, model_name: str | None = None, tool_parser: str | None = None):
        self.engine = engine
        # OpenAI-compatible text front-end (/v1/completions, /v1/chat/completions, /v1/models): needs a tokenizer; the token-id
        # /generate route the manager and the trainers use does not
        self.tokenizer, self.model_name, self.tool_parser = tokenizer, model_name or engine.full_cfg.name, tool_parser
        # tensor-parallel serving: this object lives on TP rank 0; every scheduler iteration its command list is
        # broadcast over ``tp_cmd_group`` (a gloo group of the TP ranks) so that followers (``run_tp_follower``) apply the
        # same requests / aborts / weight updates and then take the same engine step (reference: SGLang broadcasts
        # scheduler inputs to its TP workers; weight fan-out rlboost/sglang/patches.py:185-241)
        self.tp_cmd_group = tp_cmd_group
        self._rid = 0
        # cumulative SSE chunks cost O(tokens) each to build and to parse: every 128 tokens is plenty for the manager's
        # partial-result bookkeeping (an abort flushes the latest state anyway); POLYRL_STREAM_INTERVAL overrides
        self.stream_interval = int(os.environ.get("POLYRL_STREAM_INTERVAL", stream_interval))
        self.weight_update_fn = weight_update_fn        # callable(payload dict) -> (ok, message)
        self._lock = threading.Lock()
        self._inbox: list = []                          # (kind, payload) commands for the engine thread
        self._streams: dict[str, tuple[queue.Queue, list[_Tracked]]] = {}
        self._stop = threading.Event()
        self._paused = threading.Event()
        self.weight_version = 0
        srv = self

        class H(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, *a):
                return

            def _json(self, code, obj):
                data = json.dumps(obj).encode()
                self.send_response(code)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)

            def do_GET(self):
                path = self.path.split("?")[0]
                if path in ("/health", "/health_generate"):
                    self._json(200, {"status": "ok"})
                elif path == "/get_server_info":
                    info = srv.engine.server_info()
                    self._json(200, {"internal_states": [info], "weight_version": srv.weight_version})
                elif path == "/v1/models":
                    self._json(200, {"object": "list", "data": [{"id": srv.model_name, "object": "model", "owned_by": "polyrl_b200",
                                                                  "max_model_len": srv.engine.max_model_len}]})
                else:
                    self._json(404, {"error": "not found"})

            def do_POST(self):
                n = int(self.headers.get("Content-Length", 0))
                raw = self.rfile.read(n) if n else b""
                try:
                    body = json.loads(raw) if raw else {}
                except Exception:
                    self._json(400, {"error": "invalid json"})
                    return
                path = self.path.split("?")[0]
                if path == "/generate":
                    srv._handle_generate(self, body)
                elif path in ("/v1/chat/completions", "/v1/completions"):
                    from .openai_api import handle_openai

                    handle_openai(srv, self, body, chat=path.endswith("chat/completions"))
                elif path == "/abort_request":
                    srv._command("abort", body)
                    self._json(200, {"success": True})
                elif path == "/shutdown":
                    self._json(200, {"success": True})
                    srv._stop.set()
                elif path == "/update_weights_from_agent":
                    ok, msg = srv._update_weights(body)
                    self._json(200 if ok else 500, {"success": ok, "message": msg})
                elif path == "/update_weights_from_tensor":
                    self._json(501, {"success": False, "message": "in-process engines alias trainer weights; no tensor upload needed"})
                elif path == "/release_memory_occupation":
                    srv._command("release", body, wait=True)
                    self._json(200, {"success": True})
                elif path == "/resume_memory_occupation":
                    srv._command("resume", body, wait=True)
                    self._json(200, {"success": True})
                elif path == "/flush_cache":
                    self._json(200, {"success": True})
                else:
                    self._json(404, {"error": "not found"})

        self.httpd = ThreadingHTTPServer((host, port), H)
        self.httpd.daemon_threads = True
        self.host, self.port = host, self.httpd.server_address[1]
        self._http_thread = threading.Thread(target=self.httpd.serve_forever, daemon=True)
        self._engine_thread = threading.Thread(target=self._engine_loop, daemon=True)

    @property
    def endpoint(self) -> str:
        return f"{self.host}:{self.port}"

    def start(self) -> "EngineServer":
        self._http_thread.start()
        self._engine_thread.start()
        return self

    def stop(self) -> None:
        self._stop.set()
        self.httpd.shutdown()
        self.httpd.server_close()

    def serve_until_shutdown(self, poll_s: float = 0.2) -> None:
        while not self._stop.is_set():
            time.sleep(poll_s)
        self.stop()

    # ------------------------------------------------------------------ manager registration
    def register_with_manager(self, manager_endpoint: str, local: bool = False, handshake_port: int = 21000, timeout: float = 600) -> dict:
        url = f"http://{manager_endpoint}"
        if local:
            body, path = [[self.host, self.port]], "/register_local_rollout_instances"
        else:
            body, path = {"host": self.host, "port": self.port, "mooncake_handshake_port": handshake_port}, "/register_rollout_instance"
        req = urllib.request.Request(url + path, data=json.dumps(body).encode(), method="POST",
                                     headers={"Content-Type": "application/json"})
        with urllib.request.urlopen(req, timeout=timeout) as r:
            return json.loads(r.read().decode())

    # ------------------------------------------------------------------ commands executed on the engine thread
    def _command(self, kind: str, payload, wait: bool = False):
        done = threading.Event() if wait else None
        with self._lock:
            self._inbox.append((kind, payload, done))
        if done is not None:
            done.wait(timeout=120)

    def _update_weights(self, body: dict):
        """The payload data already sits in this GPU's weight buffers (peer-memory push) or is fetched by
        ``weight_update_fn``; here we only fence: drain the engine, apply, bump the version."""
        done = threading.Event()
        result = {}
        with self._lock:
            self._inbox.append(("update_weights", (body, result), done))
        done.wait(timeout=600)
        return result.get("ok", False), result.get("msg", "timeout")

    def submit(self, prompts: list[list[int]], sp: SamplingParams, mm=None) -> tuple[str, "queue.Queue"]:
        """Queue token-id prompts for the engine thread; events (index, payload, finished) arrive on the returned queue."""
        q: queue.Queue = queue.Queue()
        key = f"g{id(q)}"
        with self._lock:
            self._inbox.append(("generate", (key, prompts, sp, q, mm), None))
        return key, q

    def _handle_generate(self, h, body: dict):
        ids = body.get("input_ids")
        if ids is None:
            h._json(400, {"error": "input_ids required (text prompts need a tokenizer front-end)"})
            return
        batch = bool(ids) and isinstance(ids[0], list)
        prompts = ids if batch else [ids]
        sp = SamplingParams.from_dict(body.get("sampling_params") or {})
        stream = bool(body.get("stream", False))
        mm = None
        if body.get("image_data") is not None:          # one entry for the whole request, or one per prompt
            from ..models.vision import decode_mm_wire

            img = body["image_data"]
            mm = [decode_mm_wire(d) if d is not None else None for d in img] if isinstance(img, list) else [decode_mm_wire(img)] * len(prompts)
        key, q = self.submit(prompts, sp, mm)
        n = len(prompts)
        latest = [None] * n
        if stream:
            h.send_response(200)
            h.send_header("Content-Type", "text/event-stream")
            h.send_header("Transfer-Encoding", "chunked")
            h.end_headers()

        def write(s: str):
            d = s.encode()
            h.wfile.write(f"{len(d):x}\r\n".encode() + d + b"\r\n")
            h.wfile.flush()

        finished = 0
        try:
            while finished < n:
                ev = q.get(timeout=3600)
                idx, payload, fin = ev
                latest[idx] = payload
                if fin:
                    finished += 1
                if stream:
                    write("data: " + json.dumps(payload) + "\n\n")
            if stream:
                write("data: [DONE]\n\n")
                h.wfile.write(b"0\r\n\r\n")
                h.wfile.flush()
            else:
                h._json(200, latest if batch else latest[0])
        except (BrokenPipeError, ConnectionResetError, OSError, queue.Empty):
            self._command("abort_key", key)

    # ------------------------------------------------------------------ engine thread
    @staticmethod
    def _payload(t: _Tracked) -> dict:
        r = t.req
        fr = None
        if r.finished:
            fr = {"type": r.finish_reason}
            if r.finish_reason == "length":
                fr["length"] = len(r.output_ids)
        return {"text": "", "index": t.index, "output_ids": list(r.output_ids),
                "meta_info": {"id": r.rid, "finish_reason": fr, "prompt_tokens": len(r.prompt_ids),
                              "completion_tokens": len(r.output_ids),
                              "output_token_logprobs": [[lp, tok, None] for lp, tok in zip(r.logprobs, r.output_ids)]}}

    def _engine_loop(self) -> None:
        eng = self.engine
        if eng.is_cuda:
            import torch

            torch.cuda.set_device(eng.device)       # the current CUDA device is per THREAD: without this the engine thread issues
                                                    # its barrier kernels / NCCL calls on cuda:0 (found on the 2-GPU disaggregated run)
        while not self._stop.is_set():
            with self._lock:
                cmds, self._inbox = self._inbox, []
            wire = []                                   # serialisable mirror of the commands (TP followers)
            for kind, payload, done in cmds:
                try:
                    if kind == "generate":
                        key, prompts, sp, q, mm = payload
                        rids = []
                        for _ in prompts:
                            self._rid += 1
                            rids.append(f"s{self._rid}")
                        wire.append(("generate", {"prompts": prompts, "sp": sp.__dict__, "rids": rids, "mm": mm}))
                        tracked = [_Tracked(r, i) for i, r in enumerate(_add_group(eng, prompts, sp, rids, mm))]
                        self._streams[key] = (q, tracked)
                    elif kind == "abort_key":
                        ent = self._streams.pop(payload, None)
                        rids = [t.req.rid for t in ent[1]] if ent else []
                        wire.append(("abort_rids", rids))
                        _apply_wire(eng, "abort_rids", rids, None)
                    elif kind == "update_weights":
                        body, result = payload
                        if self.tp_cmd_group is not None:
                            _tp_broadcast(self.tp_cmd_group, wire + [("update_weights", body)], 0)
                            wire = []
                        while eng.has_work() and not body.get("bootstrap"):
                            if self.tp_cmd_group is not None:
                                _tp_broadcast(self.tp_cmd_group, [], 1)
                            eng.step()              # drain in-flight requests on the old weights
                            self._flush_streams()
                        if self.tp_cmd_group is not None:
                            _tp_broadcast(self.tp_cmd_group, [("apply_weights", body)], 0)
                        ok, msg = True, "ok"
                        if self.weight_update_fn is not None:
                            ok, msg = self.weight_update_fn(body)
                        if ok and not body.get("bootstrap"):
                            self.weight_version = int(body.get("weight_version", self.weight_version + 1))
                            eng.weight_version = self.weight_version
                        result["ok"], result["msg"] = ok, msg
                    else:                           # abort / release / resume
                        wire.append((kind, payload))
                        _apply_wire(eng, kind, payload, None)
                except Exception as e:  # keep the engine thread alive
                    if kind == "update_weights":
                        payload[1]["ok"], payload[1]["msg"] = False, repr(e)
                finally:
                    if done is not None:
                        done.set()
            do_step = eng.layer_w is not None and eng.has_work()
            if self.tp_cmd_group is not None and (wire or do_step):
                _tp_broadcast(self.tp_cmd_group, wire, int(do_step))
            if do_step:
                eng.step()
                self._flush_streams()
            else:
                time.sleep(0.002)
        if self.tp_cmd_group is not None:
            _tp_broadcast(self.tp_cmd_group, [("shutdown", None)], 0)

    def _flush_streams(self) -> None:
        for key in list(self._streams):
            q, tracked = self._streams[key]
            alive = False
            for t in tracked:
                n = len(t.req.output_ids)
                if t.req.finished and t.sent >= 0:
                    q.put((t.index, self._payload(t), True))
                    t.sent = -1
                elif t.sent >= 0:
                    alive = True
                    if n - t.sent >= self.stream_interval:
                        q.put((t.index, self._payload(t), False))
                        t.sent = n
            if not alive:
                self._streams.pop(key, None)


# ---------------------------------------------------------------------------------------------- TP lock-step helpers
def _add_group(eng: RolloutEngine, prompts, sp: SamplingParams, rids, mm=None) -> list[Request]:
    """Requests of one HTTP call; identical prompts (with the same image inputs) share their prompt KV as an n-sample group."""
    same_mm = mm is None or all(m is mm[0] for m in mm)
    tag = object() if len(prompts) > 1 and same_mm and all(p == prompts[0] for p in prompts) else None
    return [eng.add_request(list(map(int, p)), sp, rid=rid, group=tag, mm=mm[i] if mm is not None else None)
            for i, (p, rid) in enumerate(zip(prompts, rids))]


def _apply_wire(eng: RolloutEngine, kind: str, payload, weight_update_fn) -> bool:
    """Apply one broadcast command on an engine replica.  Returns False on shutdown."""
    if kind == "generate":
        _add_group(eng, payload["prompts"], SamplingParams(**{**payload["sp"], "stop_token_ids": tuple(payload["sp"].get("stop_token_ids", ()))}),
                   payload["rids"], payload.get("mm"))
    elif kind == "abort_rids":
        for rid in payload:
            eng.abort(rid)
    elif kind == "abort":
        eng.abort(payload.get("rid") or None, abort_all=bool(payload.get("abort_all", not payload.get("rid"))))
    elif kind == "release":
        eng.abort(abort_all=True)
        eng.release_memory(tuple((payload or {}).get("tags") or ("kv_cache",)))
    elif kind == "resume":
        eng.resume_memory(tuple((payload or {}).get("tags") or ("kv_cache",)))
    elif kind == "apply_weights":
        if weight_update_fn is not None:
            weight_update_fn(payload)
        if not payload.get("bootstrap"):
            eng.weight_version = int(payload.get("weight_version", eng.weight_version + 1))
    elif kind == "shutdown":
        return False
    return True


def _tp_broadcast(group, wire: list, do_step: int, src_is_me: bool = True):
    """One scheduler tick on the command channel: [commands, do_step] from TP rank 0 to the followers."""
    import torch.distributed as dist

    box = [(wire, do_step)] if src_is_me else [None]
    dist.broadcast_object_list(box, src=dist.get_process_group_ranks(group)[0], group=group)
    return box[0]


def run_tp_follower(engine: RolloutEngine, tp_cmd_group, weight_update_fn=None) -> None:
    """Main loop of TP ranks > 0: mirror rank 0's commands, then take the same engine step."""
    if engine.is_cuda:
        import torch

        torch.cuda.set_device(engine.device)
    while True:
        wire, do_step = _tp_broadcast(tp_cmd_group, None, 0, src_is_me=False)
        alive = True
        for kind, payload in wire:
            if kind == "update_weights":
                continue                    # rank 0 drains with explicit do_step ticks, then sends apply_weights
            alive = _apply_wire(engine, kind, payload, weight_update_fn) and alive
        if not alive:
            return
        if do_step:
            engine.step()
