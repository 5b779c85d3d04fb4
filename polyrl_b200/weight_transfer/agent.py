"""Trainer-side weight sender: the manager-facing half of the weight-version protocol that the reference runs
in a separate "sender agent" process (rlboost/weight_transfer/sender_agent.py:324-340,528-647): after the
version bump it asks the manager which instances are stale (``/get_receive_instances`` — the manager CAS-marks
them ``updating_weight``), transfers, then ``/update_weights`` so the manager fans out
``/update_weights_from_agent`` and re-activates instances whose version is still the latest.  Newly joined
(elastic) instances are picked up by ``poll_and_update`` on later calls."""
from __future__ import annotations

import http.client
import json
import threading

import torch.distributed as dist


class TrainerWeightSender:
    def __init__(self, manager_endpoint: str, sync, sender_endpoint: str = "127.0.0.1:18861", group=None):
        self.host, port = manager_endpoint.rsplit(":", 1)
        self.port = int(port)
        self.sync = sync
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
        http_thread = None
        if rank == 0:
            if new_version is None:
                _, out = self._post("/update_weight_version")
                new_version = int(out["new_weight_version"])
            _, out = self._post("/get_receive_instances", {"weight_sender_endpoint": self.sender_endpoint,
                                                           "sender_weight_version": new_version})
            eps = [i["instance"]["endpoint"] for i in out.get("instances", [])]
            n_targets = len(eps)
            if eps:
                body = {"instance_endpoints": eps, "weight_version": new_version, "tensors_meta": [], "load_format": "flat",
                        "flush_cache": True, "bootstrap": False}
                http_thread = threading.Thread(target=self._post, args=("/update_weights", body), daemon=True)
                http_thread.start()     # engines now block in sync.receive() until we push
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            box = [(new_version, n_targets)]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            new_version, n_targets = box[0]
        if n_targets > 0:
            self.sync.push(engine, new_version)
        if http_thread is not None:
            http_thread.join(timeout=3600)
        self.version = new_version
        return n_targets
