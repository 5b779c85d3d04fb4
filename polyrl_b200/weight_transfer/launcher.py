"""Manager launcher and sender registration helpers (reference: rlboost/weight_transfer/launcher.py:14-106 spawns the
Rust manager with ``cargo run --release``; here it is the in-tree C++ binary) plus NIC helpers
(rlboost/weight_transfer/utils.py:12-53)."""
from __future__ import annotations

import ipaddress
import json
import subprocess
import urllib.request

from ..manager.build import build_manager


def spawn_rollout_manager(port: int = 5000, bind_host: str = "0.0.0.0", config_path: str | None = None,
                          initial_local_gen_s: int | None = None, extra_args: list[str] | None = None, log_level: str = "info"):
    """Start the C++ manager; returns (Popen, bound_port).  ``port=0`` picks a free port."""
    binary = build_manager()
    if binary is None:
        raise RuntimeError("rollout manager sources not found")
    cmd = [str(binary), "--bind-addr", f"{bind_host}:{port}", "--log-level", log_level]
    if config_path:
        cmd += ["--config-file", config_path]
    if initial_local_gen_s is not None:
        cmd += ["--initial-local-gen-s", str(int(initial_local_gen_s))]
    cmd += extra_args or []
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True)
    line = p.stdout.readline()
    if not line.startswith("ROLLOUT_MANAGER_PORT="):
        p.terminate()
        raise RuntimeError(f"rollout manager failed to start: {line!r}")
    return p, int(line.strip().split("=")[1])


def get_node_ips() -> list[str]:
    """IPv4 addresses of all up, non-loopback interfaces (loopback last as a fallback)."""
    ips: list[str] = []
    try:
        import psutil

        for name, addrs in psutil.net_if_addrs().items():
            for a in addrs:
                if a.family.name == "AF_INET" and not a.address.startswith("127."):
                    ips.append(a.address)
    except Exception:
        pass
    return ips + ["127.0.0.1"]


def filter_ips_by_config(ips: list[str], allowed: str) -> list[str]:
    nets = [ipaddress.ip_network(x.strip(), strict=False) for x in allowed.split(",") if x.strip()]
    return [ip for ip in ips if any(ipaddress.ip_address(ip) in n for n in nets)]


def register_weight_senders(manager_endpoint: str, node_ip_lists: list[list[str]], groups_per_sender: int | None = None,
                            engines_per_group: int | None = None) -> dict:
    body: dict = {"weight_sender_ips": node_ip_lists}
    if groups_per_sender:
        body["num_mooncake_groups_per_sender"] = groups_per_sender
    if engines_per_group:
        body["num_mooncake_engines_per_group"] = engines_per_group
    req = urllib.request.Request(f"http://{manager_endpoint}/update_weight_senders", data=json.dumps(body).encode(), method="PUT",
                                 headers={"Content-Type": "application/json"})
    with urllib.request.urlopen(req, timeout=30) as r:
        return json.loads(r.read().decode())
