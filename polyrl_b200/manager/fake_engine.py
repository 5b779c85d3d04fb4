"""In-process fake rollout engine speaking the SGLang-compatible HTTP surface the manager relies on
(SURVEY Appendix B): ``/generate`` (SSE, cumulative chunks), ``/health_generate``, ``/get_server_info``,
``/abort_request``, ``/shutdown``, ``/update_weights_from_agent``.  Deterministic "model": next token =
(sum(input_ids so far) * 31 + 7) % vocab — so a continuation resumed on another engine must reproduce
exactly the tokens an uninterrupted run would have produced.  Fault injection: ``die_after_tokens`` closes
the socket mid-stream (engine crash), ``abort`` flags emit finish_reason abort.  Used by the C++ manager
tests (the reference has none) and by CPU plumbing tests of the streaming trainer.
"""
from __future__ import annotations

import json
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

VOCAB = 1000


def next_token(ids: list[int]) -> int:
    return (sum(ids) * 31 + 7) % VOCAB


def reference_generation(prompt: list[int], max_new_tokens: int) -> list[int]:
    ids, out = list(prompt), []
    for _ in range(max_new_tokens):
        t = next_token(ids)
        ids.append(t)
        out.append(t)
    return out


class FakeEngine:
    def __init__(self, token_delay_s: float = 0.002, stream_interval: int = 4, die_after_tokens: int | None = None,
                 healthy: bool = True, port: int = 0):
        self.token_delay_s = token_delay_s
        self.stream_interval = stream_interval
        self.die_after_tokens = die_after_tokens
        self.healthy = healthy
        self.abort_all = threading.Event()
        self.running = 0
        self.lock = threading.Lock()
        self.weight_updates: list[dict] = []
        self.generate_calls = 0
        self.shutdown_called = threading.Event()
        eng = self

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
                if self.path.startswith("/health_generate") or self.path.startswith("/health"):
                    self._json(200 if eng.healthy else 503, {"ok": eng.healthy})
                elif self.path.startswith("/get_server_info"):
                    with eng.lock:
                        r = eng.running
                    self._json(200, {"internal_states": [{"#running_req": r, "#queue_req": 0, "last_gen_throughput": 100.0}]})
                else:
                    self._json(404, {})

            def do_POST(self):
                n = int(self.headers.get("Content-Length", 0))
                body = json.loads(self.rfile.read(n) or b"{}") if n else {}
                path = self.path.split("?")[0]
                if path == "/generate":
                    eng._generate(self, body)
                elif path == "/abort_request":
                    eng.abort_all.set()
                    self._json(200, {"success": True})
                elif path == "/shutdown":
                    eng.shutdown_called.set()
                    self._json(200, {"success": True})
                elif path == "/update_weights_from_agent":
                    eng.weight_updates.append(body)
                    self._json(200, {"success": True, "message": "ok"})
                elif path in ("/flush_cache", "/release_memory_occupation", "/resume_memory_occupation"):
                    self._json(200, {"success": True})
                else:
                    self._json(404, {})

        self.server = ThreadingHTTPServer(("127.0.0.1", port), H)
        self.server.daemon_threads = True
        self.port = self.server.server_address[1]
        self.thread = threading.Thread(target=self.server.serve_forever, daemon=True)
        self.thread.start()

    @property
    def endpoint(self) -> str:
        return f"127.0.0.1:{self.port}"

    def stop(self):
        self.server.shutdown()
        self.server.server_close()

    def _generate(self, h, body):
        self.generate_calls += 1
        ids = body["input_ids"]
        batch = bool(ids) and isinstance(ids[0], list)
        prompts = [list(p) for p in ids] if batch else [list(ids)]
        max_new = int(body.get("sampling_params", {}).get("max_new_tokens", 16))
        with self.lock:
            self.running += 1
        h.send_response(200)
        h.send_header("Content-Type", "text/event-stream")
        h.send_header("Transfer-Encoding", "chunked")
        h.end_headers()

        def chunk(s: str):
            d = s.encode()
            h.wfile.write(f"{len(d):x}\r\n".encode() + d + b"\r\n")
            h.wfile.flush()

        outs = [[] for _ in prompts]
        finish = "length"
        try:
            for step in range(max_new):
                if self.abort_all.is_set():
                    finish = "abort"
                    break
                for i, p in enumerate(prompts):
                    t = next_token(p + outs[i])
                    outs[i].append(t)
                time.sleep(self.token_delay_s)
                if self.die_after_tokens is not None and step + 1 >= self.die_after_tokens:
                    self._emit(chunk, outs, batch, None)
                    h.wfile.flush()
                    h.connection.shutdown(2)       # crash: no [DONE]
                    h.close_connection = True
                    return
                if (step + 1) % self.stream_interval == 0:
                    self._emit(chunk, outs, batch, None)
            self._emit(chunk, outs, batch, finish)
            chunk("data: [DONE]\n\n")
            h.wfile.write(b"0\r\n\r\n")
            h.wfile.flush()
        except (BrokenPipeError, ConnectionResetError, OSError):
            pass
        finally:
            with self.lock:
                self.running -= 1
            if finish == "abort":
                self.abort_all.clear()

    @staticmethod
    def _emit(chunk, outs, batch, finish):
        for i, o in enumerate(outs):
            msg = {"text": "", "index": i,
                   "meta_info": {"output_token_logprobs": [[-0.25, t, None] for t in o], "completion_tokens": len(o),
                                 "finish_reason": {"type": finish} if finish else None}}
            chunk("data: " + json.dumps(msg) + "\n\n")
            if not batch:
                break
