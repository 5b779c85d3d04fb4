"""Trainer-side weight sender: the manager-facing half of the weight-version protocol that the reference runs
in a separate "sender agent" process (rlboost/weight_transfer/sender_agent.py:324-340,528-647): after the
version bump it asks the manager which instances are stale (``/get_receive_instances`` — the manager CAS-marks
them ``updating_weight``), transfers, then ``/update_weights`` so the manager fans out
``/update_weights_from_agent`` and re-activates instances whose version is still the latest.  Newly joined
(elastic) instances are picked up by ``poll_and_update`` on later calls.

Two transports, chosen per instance: rollout ranks of this job's own process group (``sync_endpoints``) receive the peer-memory /
broadcast push of ``WeightSync``; every other instance — an elastic engine on another host started with
``python -m polyrl_b200.rollout.launch_server --enable-weight-transfer-agent`` — gets the full bf16 state over the parallel TCP
streams of ``tcp_engine.py`` at the handshake address it registered (reference path: FSDP full_tensor -> pinned shm -> TCP sendfile ->
receiver buffer, rlboost/weight_transfer/fsdp_interface.py:186-234 + transfer_engine.py:154-268)."""
from __future__ import annotations

import http.client
import json
import threading

import torch.distributed as dist


class TrainerWeightSender:
    def __init__(self, manager_endpoint: str, sync, sender_endpoint: str = "127.0.0.1:18861", group=None,
                 sync_endpoints: set[str] | None = None, tcp_streams: int = 1):
        self.host, port = manager_endpoint.rsplit(":", 1)
        self.port = int(port)
        self.sync = sync
        # endpoints served by ``sync`` (None: every listed instance is; with ``sync=None`` none is)
        self.sync_endpoints = sync_endpoints
        self.tcp_streams = tcp_streams
        self._staging = None
        self._tcp = None
        self.failed_targets: list[str] = []
        self.sender_endpoint = sender_endpoint
        self.group = group
        self.version = 0

    def _post(self, path: str, body: dict | None = None, method: str = "POST", timeout: float = 3600):
        c = http.client.HTTPConnection(self.host, self.port, timeout=timeout)
        c.request(method, path, body=None if body is None else json.dumps(body), headers={"Content-Type": "application/json"})
        r = c.getresponse()
        data = r.read().decode()
        c.close()
        return r.status, (json.loads(data) if data else {})

    def register_as_sender(self, ips: list[str] | None = None) -> None:
        host = self.sender_endpoint.rsplit(":", 1)[0]
        self._post("/update_weight_senders", {"weight_sender_ips": [ips or [host]]}, method="PUT")

    def update(self, engine, rank: int, new_version: int | None = None) -> int:
        """Collective over the trainer ranks.  rank 0 talks to the manager; everybody pushes."""
        n_targets = 0
        tcp_targets: list[str] = []
        http_thread = None
        if rank == 0:
            if new_version is None:
                _, out = self._post("/update_weight_version")
                new_version = int(out["new_weight_version"])
            _, out = self._post("/get_receive_instances", {"weight_sender_endpoint": self.sender_endpoint,
                                                           "sender_weight_version": new_version})
            insts = [i["instance"] for i in out.get("instances", [])]
            eps = [i["endpoint"] for i in insts]
            in_sync = (lambda ep: self.sync is not None and (self.sync_endpoints is None or ep in self.sync_endpoints))
            tcp_targets = [i.get("mooncake_handshake_addr") for i in insts if not in_sync(i["endpoint"]) and i.get("mooncake_handshake_addr")]
            n_targets = sum(1 for ep in eps if in_sync(ep))
            if eps:
                body = {"instance_endpoints": eps, "weight_version": new_version, "tensors_meta": [], "load_format": "flat",
                        "flush_cache": True, "bootstrap": False}
                http_thread = threading.Thread(target=self._post, args=("/update_weights", body), daemon=True)
                http_thread.start()     # engines now block in sync.receive() until we push
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            box = [(new_version, n_targets, tcp_targets)]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            new_version, n_targets, tcp_targets = box[0]
        if n_targets > 0:
            self.sync.push(engine, new_version)
        if tcp_targets:
            self._tcp_push(engine, rank, tcp_targets)
        if http_thread is not None:
            http_thread.join(timeout=3600)
        self.version = new_version
        return n_targets + len(tcp_targets)

    def _tcp_push(self, engine, rank: int, targets: list[str]) -> None:
        """Collective over the trainer ranks (every unit is gathered); rank 0 stages the full state in the memfd buffer, in unit order
        (text stack, then the vision tower), and streams it to every target's receiver."""
        import torch

        from .tcp_engine import StagingBuffer, TCPWeightSender

        if getattr(engine, "lora", None) is not None:
            raise NotImplementedError("LoRA adapters are merged on the colocated path only; export + reload for remote instances")
        units = engine.base_units + ([engine.vision] if getattr(engine, "vision", None) is not None else [])
        es = torch.empty(0, dtype=engine.param_dtype).element_size()
        nbytes = sum(sum(u.numels) for u in units) * es
        if rank == 0 and (self._staging is None or self._staging.nbytes != nbytes):
            self._staging = StagingBuffer(nbytes)
            self._tcp = TCPWeightSender(num_streams=self.tcp_streams)
        off = 0
        for u in units:
            full = u.acquire()
            if rank == 0:
                for t in u.views(full).values():
                    n = t.numel() * es
                    self._staging.tensor[off:off + n].copy_(t.detach().contiguous().view(torch.uint8).reshape(-1), non_blocking=True)
                    off += n
            u.release()
        if rank == 0:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self.failed_targets = []
            for addr in targets:
                host, port = addr.rsplit(":", 1)
                try:                          # a spot instance may vanish at any moment: never let that take the trainer down; the
                    self._tcp.send(self._staging, host, int(port), nbytes)      # manager drops it at the next failed health check
                except (OSError, ConnectionError) as e:
                    self.failed_targets.append(addr)
                    print(f"[weight-sender] transfer to {addr} failed ({e}); skipping this instance", flush=True)
