"""Trainer-side client of the rollout manager: submits a prompt batch to ``/batch_generate_requests`` and
turns the NDJSON result stream into ``DataProto`` micro-batches of ``stream_size`` samples while generation
is still in flight.

Reference behaviour: payload builder — one request per prompt with ``input_ids`` repeated ``n`` times and
``return_logprob`` (rlboost/.../sglang_rollout_remote.py:198-225); ``_launch_generate_remote`` — first line
is the notifier, responses accumulate, a batch is yielded every ``stream_size`` samples, the remainder is
flushed at the end (:393-482); ``postprocess_samples`` (:318-391); ``StreamingBatchIterator`` groups lines
that arrive within a short drain window (rlboost/.../stream_batch_iter.py:19-83).
"""
from __future__ import annotations

import http.client
import json
import queue
import threading
import time
from typing import Iterator

import numpy as np
import torch

from ..protocol import DataProto


class StreamingBatchIterator:
    """Iterate an NDJSON HTTP response; lines that arrive within ``drain_timeout`` are returned together."""

    def __init__(self, host: str, port: int, path: str, payload, drain_timeout: float = 0.01, timeout: float = 3600.0):
        self.q: queue.Queue = queue.Queue()
        self.drain_timeout = drain_timeout
        self.error: Exception | None = None
        self._t = threading.Thread(target=self._run, args=(host, port, path, payload, timeout), daemon=True)
        self._t.start()

    def _run(self, host, port, path, payload, timeout):
        try:
            conn = http.client.HTTPConnection(host, port, timeout=timeout)
            conn.request("POST", path, body=json.dumps(payload), headers={"Content-Type": "application/json"})
            resp = conn.getresponse()
            if resp.status != 200:
                raise RuntimeError(f"manager returned HTTP {resp.status}: {resp.read()[:300]!r}")
            for raw in resp:
                line = raw.strip()
                if line:
                    self.q.put(json.loads(line))
            conn.close()
        except Exception as e:  # surfaced on the consumer side
            self.error = e
        finally:
            self.q.put(None)

    def __iter__(self) -> Iterator[list[dict]]:
        done = False
        while not done:
            first = self.q.get()
            if first is None:
                break
            group = [first]
            deadline = time.time() + self.drain_timeout
            while True:
                try:
                    nxt = self.q.get(timeout=max(0.0, deadline - time.time()))
                except queue.Empty:
                    break
                if nxt is None:
                    done = True
                    break
                group.append(nxt)
            yield group
        if self.error is not None:
            raise self.error


def build_generate_payload(prompts: DataProto, n: int, sampling_params: dict) -> list:
    """[[id, {input_ids: [prompt]*n, sampling_params, return_logprob}], ...] with 1-based ids (0 = notifier)."""
    raw = prompts.non_tensor_batch.get("raw_prompt_ids")
    if raw is None:
        ids, mask = prompts.batch["input_ids"], prompts.batch["attention_mask"].bool()
        raw = [ids[i][mask[i]].tolist() for i in range(len(prompts))]
    sp = dict(sampling_params)
    sp.pop("n", None)
    out = [[i + 1, {"input_ids": [list(map(int, raw[i]))] * n, "sampling_params": sp, "return_logprob": True}]
           for i in range(len(prompts))]
    mm = prompts.non_tensor_batch.get("multi_modal_inputs")
    if mm is not None:                        # image prompts: the n samples of a request share one ``image_data`` entry
        from ..models.vision import encode_mm_wire

        for i, m in enumerate(mm):
            if m is not None:
                out[i][1]["image_data"] = encode_mm_wire(m)
    return out


def postprocess_samples(prompts: DataProto, results: list[tuple[int, list[dict]]], n: int, response_length: int,
                        pad_token_id: int, eos_token_id: int) -> DataProto:
    """results: list of (prompt_index, n sub-responses).  Samples of one prompt stay adjacent and share uid."""
    dev = prompts.batch["input_ids"].device
    idx = torch.tensor([i for i, _ in results], dtype=torch.long)
    sel = prompts.select_idxs(idx).repeat(n, interleave=True)
    b = len(sel)
    resp = torch.full((b, response_length), pad_token_id, dtype=torch.long)
    lps = torch.zeros(b, response_length)
    lens = torch.zeros(b, dtype=torch.long)
    aborted = np.zeros(b, dtype=bool)
    k = 0
    for _, subs in results:
        for s in subs:
            trip = s.get("meta_info", {}).get("output_token_logprobs") or []
            toks = [int(t[1]) for t in trip][:response_length]
            m = len(toks)
            if m:
                resp[k, :m] = torch.tensor(toks)
                lps[k, :m] = torch.tensor([float(t[0]) for t in trip][:m])
            lens[k] = m
            fr = (s.get("meta_info", {}).get("finish_reason") or {}).get("type")
            aborted[k] = fr == "abort"
            k += 1
    resp, lps, lens = resp.to(dev), lps.to(dev), lens.to(dev)
    ar = torch.arange(response_length, device=dev)[None]
    response_mask = (ar < lens[:, None]).long()
    p_ids, p_mask = sel.batch["input_ids"], sel.batch["attention_mask"]
    p_pos = sel.batch.get("position_ids")
    if p_pos is None:
        p_pos = (p_mask.cumsum(-1) - 1).clamp_min(0)
    sel.batch["prompts"] = p_ids
    sel.batch["responses"] = resp
    sel.batch["input_ids"] = torch.cat([p_ids, resp], dim=1)
    sel.batch["attention_mask"] = torch.cat([p_mask, response_mask], dim=1)
    delta = torch.arange(1, response_length + 1, device=dev)
    if p_pos.dim() == 3:                      # M-RoPE [b, 3, S]: responses continue every axis after the largest prompt position
        nxt = p_pos.amax(dim=(1, 2)).view(-1, 1, 1) + delta.view(1, 1, -1)
        sel.batch["position_ids"] = torch.cat([p_pos, nxt.expand(-1, 3, -1)], dim=2)
    else:
        sel.batch["position_ids"] = torch.cat([p_pos, p_pos[:, -1:] + delta[None]], dim=1)
    sel.batch["response_mask"] = response_mask
    sel.batch["rollout_log_probs"] = lps
    sel.non_tensor_batch["aborted"] = aborted
    return sel


class RemoteRolloutClient:
    """generate_sequences_remote(prompts) -> iterator of DataProto micro-batches."""

    def __init__(self, manager_endpoint: str, n: int, response_length: int, pad_token_id: int, eos_token_id: int,
                 stream_size: int = 16, sampling_params: dict | None = None):
        self.host, port = manager_endpoint.rsplit(":", 1)
        self.port = int(port)
        self.n, self.response_length = n, response_length
        self.pad, self.eos = pad_token_id, eos_token_id
        if stream_size % n != 0:
            raise ValueError("stream_size (min_stream_batch_size) must be a multiple of rollout.n")
        self.stream_size = stream_size
        self.sampling_params = sampling_params or {}
        self.timing: dict[str, float] = {}

    def _post(self, path: str, body=None, method: str = "POST") -> dict:
        conn = http.client.HTTPConnection(self.host, self.port, timeout=600)
        conn.request(method, path, body=None if body is None else json.dumps(body), headers={"Content-Type": "application/json"})
        r = conn.getresponse()
        data = r.read().decode()
        conn.close()
        try:
            return {**(json.loads(data) if data else {}), "http_status": r.status}
        except Exception:
            return {"http_status": r.status, "raw": data}

    def wait_healthy(self, timeout: float = 120.0) -> None:
        t0 = time.time()
        while time.time() - t0 < timeout:
            try:
                if self._post("/health", method="GET").get("http_status") == 200:
                    return
            except Exception:
                pass
            time.sleep(0.2)
        raise TimeoutError(f"rollout manager {self.host}:{self.port} is not healthy")

    def update_weight_version(self) -> int:
        return int(self._post("/update_weight_version")["new_weight_version"])

    def update_metrics(self, step_time_s: float, trainer_bubble_time_s: float, step_throughput: float) -> dict:
        return self._post("/update_metrics", {"step_time_s": int(step_time_s), "trainer_bubble_time_s": int(trainer_bubble_time_s),
                                              "step_throughput": float(step_throughput)})

    def instances_status(self) -> dict:
        return self._post("/get_instances_status", method="GET")

    def submit(self, prompts: DataProto, sampling_params: dict | None = None):
        """Start the generation NOW (the HTTP request is issued by a background thread) and return a handle for ``collect``."""
        sp = {**self.sampling_params, **(sampling_params or {}), "max_new_tokens": self.response_length}
        payload = build_generate_payload(prompts, self.n, sp)
        return prompts, time.time(), StreamingBatchIterator(self.host, self.port, "/batch_generate_requests", payload)

    def generate_once(self, prompts: DataProto, n: int = 1, sampling_params: dict | None = None) -> DataProto:
        """Blocking one-shot generation (validation): every prompt, ``n`` samples each, in prompt order."""
        sp = {**self.sampling_params, **(sampling_params or {}), "max_new_tokens": self.response_length}
        payload = build_generate_payload(prompts, n, sp)
        results: list[tuple[int, list[dict]]] = []
        for group in StreamingBatchIterator(self.host, self.port, "/batch_generate_requests", payload):
            for line in group:
                if line.get("id") == 0:
                    continue
                if not line.get("success"):
                    raise RuntimeError(f"generation failed for request {line.get('id')}: {line.get('error')}")
                data = line["data"]
                results.append((int(line["id"]) - 1, data if isinstance(data, list) else [data]))
        results.sort(key=lambda t: t[0])
        return postprocess_samples(prompts, results, n, self.response_length, self.pad, self.eos)

    def generate_sequences_remote(self, prompts: DataProto, sampling_params: dict | None = None) -> Iterator[DataProto]:
        return self.collect(self.submit(prompts, sampling_params))

    def collect(self, handle) -> Iterator[DataProto]:
        prompts, t0, sbi = handle
        it = iter(sbi)
        pending: list[tuple[int, list[dict]]] = []
        per = self.stream_size // self.n
        got_notifier = False
        for group in it:
            for line in group:
                if not got_notifier:
                    if line.get("id") != 0:
                        raise RuntimeError(f"expected the notifier line first, got {line}")
                    got_notifier = True
                    self.timing["local_gen"] = time.time() - t0
                    continue
                if not line.get("success"):
                    raise RuntimeError(f"generation failed for request {line.get('id')}: {line.get('error')}")
                data = line["data"]
                pending.append((int(line["id"]) - 1, data if isinstance(data, list) else [data]))
                while len(pending) >= per:
                    chunk, pending = pending[:per], pending[per:]
                    yield postprocess_samples(prompts, chunk, self.n, self.response_length, self.pad, self.eos)
        if pending:
            yield postprocess_samples(prompts, pending, self.n, self.response_length, self.pad, self.eos)
        self.timing["gen"] = time.time() - t0
