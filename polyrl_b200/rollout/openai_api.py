"""OpenAI-compatible text routes of the engine server: ``POST /v1/completions`` and ``POST /v1/chat/completions`` (blocking and
``stream: true`` server-sent events), on top of the token-id request path the rollout manager uses.

Why it exists: the reference's rollout servers (SGLang / vLLM) speak this API natively and parts of the stack depend on it —
``verl/trainer/main_generation_server.py`` (offline generation through stand-alone servers, one ``/v1/chat/completions`` call per
sample), generative reward models and evaluation harnesses that point at a served policy.  Supported request fields: ``messages`` /
``prompt`` (text, list of texts, or token ids), ``max_tokens`` / ``max_completion_tokens``, ``temperature``, ``top_p``, ``top_k``,
``n``, ``seed``, ``stop`` (strings, applied to the decoded text), ``stop_token_ids``, ``ignore_eos``, ``logprobs``, ``tools``
(rendered by the chat template; calls in the output are parsed into ``tool_calls`` when the server has a tool parser), ``stream``.
"""
from __future__ import annotations

import json
import queue
import time
import uuid

from .engine import SamplingParams


def _sampling(body: dict) -> SamplingParams:
    mx = body.get("max_completion_tokens") or body.get("max_tokens") or 128
    temp = body.get("temperature")
    return SamplingParams(max_new_tokens=int(mx), temperature=1.0 if temp is None else float(temp),
                          top_p=float(body.get("top_p") if body.get("top_p") is not None else 1.0),
                          top_k=int(body.get("top_k") if body.get("top_k") is not None else -1),
                          stop_token_ids=tuple(int(t) for t in (body.get("stop_token_ids") or ())),
                          ignore_eos=bool(body.get("ignore_eos", False)), seed=body.get("seed"))


def _cut_at_stop(text: str, stops: list[str]) -> tuple[str, bool]:
    cut = min((i for i in (text.find(s) for s in stops if s) if i >= 0), default=-1)
    return (text, False) if cut < 0 else (text[:cut], True)


def _prompts(srv, body: dict, chat: bool) -> list[list[int]]:
    tok = srv.tokenizer
    if chat:
        msgs = body.get("messages")
        if not isinstance(msgs, list) or not msgs:
            raise ValueError("messages must be a non-empty list")
        kw = {"tools": body["tools"]} if body.get("tools") else {}
        kw.update(body.get("chat_template_kwargs") or {})
        ids = tok.apply_chat_template(msgs, add_generation_prompt=True, tokenize=True, **kw)
        if isinstance(ids, dict) or hasattr(ids, "input_ids"):       # transformers returns a BatchEncoding in some versions
            ids = ids["input_ids"]
        return [list(map(int, ids))]
    p = body.get("prompt")
    if isinstance(p, str):
        return [tok.encode(p, add_special_tokens=False)]
    if isinstance(p, list) and p and all(isinstance(t, int) for t in p):
        return [p]
    if isinstance(p, list) and p and all(isinstance(t, str) for t in p):
        return [tok.encode(t, add_special_tokens=False) for t in p]
    if isinstance(p, list) and p and all(isinstance(t, list) for t in p):
        return [list(map(int, t)) for t in p]
    raise ValueError("prompt must be a string, a list of strings, token ids or a list of token-id lists")


def _finish_reason(payload: dict, stopped_on_text: bool) -> str:
    fr = (payload["meta_info"].get("finish_reason") or {}).get("type")
    return "stop" if stopped_on_text or fr == "stop" else ("length" if fr == "length" else (fr or "stop"))


def _choice(srv, payload: dict, index: int, chat: bool, stops: list[str], want_lp: bool, tools: bool) -> dict:
    tok = srv.tokenizer
    ids = payload["output_ids"]
    text, cut = _cut_at_stop(tok.decode(ids, skip_special_tokens=True), stops)
    fin = _finish_reason(payload, cut)
    lp = None
    if want_lp:
        items = [{"token": tok.decode([t], skip_special_tokens=False), "logprob": float(l), "token_id": int(t)}
                 for l, t, _ in payload["meta_info"]["output_token_logprobs"]]
        lp = {"content": items} if chat else {"tokens": [i["token"] for i in items], "token_logprobs": [i["logprob"] for i in items]}
    if not chat:
        return {"index": index, "text": text, "finish_reason": fin, "logprobs": lp}
    msg: dict = {"role": "assistant", "content": text}
    if tools and srv.tool_parser:
        from ..agent.tool_parser import get_tool_parser

        content, calls = get_tool_parser(srv.tool_parser, tok).extract_tool_calls(ids)
        if calls:
            msg["content"] = _cut_at_stop(content, stops)[0] or None
            msg["tool_calls"] = [{"id": f"call_{uuid.uuid4().hex[:12]}", "type": "function",
                                  "function": {"name": c.name, "arguments": c.arguments}} for c in calls]
            fin = "tool_calls"
    return {"index": index, "message": msg, "finish_reason": fin, "logprobs": lp}


def handle_openai(srv, h, body: dict, chat: bool) -> None:
    """Serve one request on handler ``h`` of EngineServer ``srv``."""
    if srv.tokenizer is None:
        h._json(400, {"error": {"message": "this server was started without a tokenizer: use the token-id /generate route",
                                "type": "invalid_request_error"}})
        return
    try:
        prompts = _prompts(srv, body, chat)
        sp = _sampling(body)
        n = max(1, int(body.get("n") or 1))
    except (ValueError, TypeError, KeyError) as e:
        h._json(400, {"error": {"message": str(e), "type": "invalid_request_error"}})
        return
    room = srv.engine.max_model_len - 1
    if any(len(p) > room for p in prompts):
        h._json(400, {"error": {"message": f"prompt longer than the model context ({room} tokens)", "type": "invalid_request_error"}})
        return
    stops = [body["stop"]] if isinstance(body.get("stop"), str) else list(body.get("stop") or [])
    want_lp, tools = bool(body.get("logprobs")), bool(body.get("tools"))
    flat = [p for p in prompts for _ in range(n)]                 # choice index = prompt index * n + sample index
    rid = ("chatcmpl-" if chat else "cmpl-") + uuid.uuid4().hex[:24]
    obj = "chat.completion" if chat else "text_completion"
    base = {"id": rid, "created": int(time.time()), "model": body.get("model") or srv.model_name}
    keys_qs = [srv.submit([p] * n, sp) for p in prompts]          # one engine group per prompt: its n samples share the prefill
    stream = bool(body.get("stream", False))
    if stream:
        h.send_response(200)
        h.send_header("Content-Type", "text/event-stream")
        h.send_header("Transfer-Encoding", "chunked")
        h.end_headers()

    def write(o) -> None:
        d = ("data: " + (o if isinstance(o, str) else json.dumps(o)) + "\n\n").encode()
        h.wfile.write(f"{len(d):x}\r\n".encode() + d + b"\r\n")
        h.wfile.flush()

    final: list = [None] * len(flat)
    sent_text = [""] * len(flat)
    try:
        pending = {pi: n for pi in range(len(prompts))}
        while pending:
            for pi in list(pending):
                try:
                    idx, payload, fin = keys_qs[pi][1].get(timeout=0.01 if len(pending) > 1 else 3600)
                except queue.Empty:
                    continue
                ci = pi * n + idx
                if stream:
                    text, cut = _cut_at_stop(srv.tokenizer.decode(payload["output_ids"], skip_special_tokens=True), stops)
                    delta, sent_text[ci] = text[len(sent_text[ci]):], text
                    if delta and final[ci] is None:
                        ch = {"index": ci, "delta": {"content": delta}, "finish_reason": None} if chat else {"index": ci, "text": delta, "finish_reason": None}
                        write({**base, "object": obj + ".chunk" if chat else obj, "choices": [ch]})
                if fin and final[ci] is None:
                    final[ci] = payload
                    pending[pi] -= 1
                    if pending[pi] == 0:
                        del pending[pi]
                    if stream:
                        c = _choice(srv, payload, ci, chat, stops, False, tools)
                        ch = {"index": ci, "delta": ({"tool_calls": c["message"]["tool_calls"]} if chat and c["message"].get("tool_calls") else {}),
                              "finish_reason": c["finish_reason"]} if chat else {"index": ci, "text": "", "finish_reason": c["finish_reason"]}
                        write({**base, "object": obj + ".chunk" if chat else obj, "choices": [ch]})
        usage = {"prompt_tokens": sum(len(p) for p in prompts), "completion_tokens": sum(len(f["output_ids"]) for f in final)}
        usage["total_tokens"] = usage["prompt_tokens"] + usage["completion_tokens"]
        if stream:
            write("[DONE]")
            h.wfile.write(b"0\r\n\r\n")
            h.wfile.flush()
            return
        h._json(200, {**base, "object": obj, "usage": usage,
                      "choices": [_choice(srv, f, i, chat, stops, want_lp, tools) for i, f in enumerate(final)]})
    except (BrokenPipeError, ConnectionResetError, OSError):
        for key, _ in keys_qs:
            srv._command("abort_key", key)
