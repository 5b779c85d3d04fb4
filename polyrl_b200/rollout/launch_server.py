"""Stand-alone rollout engine server (elastic remote instance):
    python -m polyrl_b200.rollout.launch_server --model-path llama3-8b --port 30000 --rollout-manager-address H:P
CLI mirrors the flags the reference adds to SGLang's launcher (rlboost/sglang/patches.py:35-75):
--enable-weight-transfer-agent, --transfer-agent-handshake-port, --weight-sender-rpyc-endpoint,
--weight-receiver-zmq-bind-host, --rollout-manager-address, --transfer-agent-sender-group-idx,
--num-transfer-engines-per-group.  Weights arrive through the TCP transfer engine (other hosts) or are
random/dummy initialised (``--load-format dummy``) until the first update."""
from __future__ import annotations

import argparse
import os

import torch
from pathlib import Path

from ..models import TransformerWeights, get_model_config
from .engine import RolloutEngine
from .server import EngineServer


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=30000)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--max-num-seqs", type=int, default=256)
    ap.add_argument("--max-model-len", type=int, default=2048)
    ap.add_argument("--mem-fraction-static", type=float, default=0.6)
    ap.add_argument("--load-format", default="auto", help="auto: load --model-path weights when it is a checkpoint directory; dummy: random init")
    ap.add_argument("--stream-interval", type=int, default=20)
    ap.add_argument("--rollout-manager-address", default=None)
    ap.add_argument("--enable-weight-transfer-agent", action="store_true")
    ap.add_argument("--transfer-agent-handshake-port", type=int, default=21000)
    ap.add_argument("--weight-sender-rpyc-endpoint", default=None)
    ap.add_argument("--weight-receiver-zmq-bind-host", default="0.0.0.0")
    ap.add_argument("--transfer-agent-sender-group-idx", type=int, default=0)
    ap.add_argument("--num-transfer-engines-per-group", type=int, default=1)
    ap.add_argument("--enable-prefix-cache", action="store_true", help="cross-request KV page reuse (radix-cache equivalent)")
    ap.add_argument("--tokenizer-path", default=None, help="HF tokenizer directory (or 'byte'): enables the OpenAI-compatible /v1/completions "
                                                           "and /v1/chat/completions routes; default: --model-path when it holds a tokenizer")
    ap.add_argument("--served-model-name", default=None)
    ap.add_argument("--tool-call-parser", default=None, help="hermes | gpt-oss | json_block: parse tool calls of chat completions into tool_calls")
    ap.add_argument("--tp-size", type=int, default=1, help="tensor-parallel degree; launch with torchrun --nproc-per-node TP")
    a = ap.parse_args(argv)
    cfg = get_model_config(a.model_path)
    dtype = torch.bfloat16 if a.device.startswith("cuda") else torch.float32
    tp_ctx = cmd_group = None
    tp_rank = 0
    if a.tp_size > 1:
        # one process per TP rank (torchrun); rank 0 owns the HTTP front-end, the others mirror its commands
        import torch.distributed as dist

        from ..parallel.symm import SymmetricHeap, make_heap
        from .tp import TPContext

        tp_rank = int(os.environ["RANK"])
        heap = None
        if a.device.startswith("cuda"):
            a.device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
            torch.cuda.set_device(a.device)
            dist.init_process_group("nccl", device_id=torch.device(a.device))
            heap = make_heap(group=None, device=torch.device(a.device))
        else:                                  # CPU plumbing path (library all-reduce)
            os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
            dist.init_process_group("gloo")
        cmd_group = dist.group.WORLD          # scheduler commands ride on the TP group itself (no extra rendezvous)
        tp_ctx = TPContext(cfg, heap, tp_rank, a.tp_size, max_batch=a.max_num_seqs, group=dist.group.WORLD)
    w = TransformerWeights(cfg, device=a.device, dtype=dtype, init=True)          # same seed on every TP rank
    vw = None
    if cfg.vision is not None:                  # vision-language model: the engine runs the ViT tower at prefill
        from ..models.vision import VisionWeights

        vw = VisionWeights(cfg, device=a.device, dtype=dtype, init=True)
    from pathlib import Path as _P

    mp = _P(a.model_path)
    if a.load_format != "dummy" and mp.is_dir() and (list(mp.glob("*.safetensors")) or list(mp.glob("pytorch_model*.bin"))):
        from ..models.hf_io import load_state_dict as _load_hf

        sd = {k: v.to(a.device, dtype) for k, v in _load_hf(cfg, mp).items()}
        w.load_state_dict(sd)
        if vw is not None:
            vw.load_state_dict(sd)
        print(f"[engine] loaded weights from {mp}", flush=True)
    eng = RolloutEngine(cfg, a.device, max_num_seqs=a.max_num_seqs, max_model_len=a.max_model_len,
                        gpu_memory_utilization=a.mem_fraction_static, tp=tp_ctx, enable_prefix_cache=a.enable_prefix_cache)
    eng.set_weights(w.embed, w.layers, w.head)
    if vw is not None:
        eng.set_vision_weights(vw)
    if tp_rank > 0:
        from .server import run_tp_follower

        run_tp_follower(eng, cmd_group)
        return 0
    receiver = None
    if a.enable_weight_transfer_agent:
        from ..weight_transfer.tcp_engine import TCPWeightReceiver

        named = list(w.named_tensors()) + (list(vw.named_tensors()) if vw is not None else [])     # trainer unit order
        receiver = TCPWeightReceiver(named, bind_host=a.weight_receiver_zmq_bind_host, port=a.transfer_agent_handshake_port,
                                     num_streams=a.num_transfer_engines_per_group)

    def on_update(body):
        if receiver is not None and not body.get("bootstrap"):
            try:
                receiver.wait_and_apply(timeout=float(os.environ.get("POLYRL_WEIGHT_RECV_TIMEOUT_S", "900")))
            except TimeoutError as e:         # keep serving the old policy; the manager sees the failure and keeps the instance stale
                return False, str(e)
        return True, "ok"

    tokenizer = None
    tok_path = a.tokenizer_path or (a.model_path if any((Path(a.model_path) / f).exists() for f in ("tokenizer.json", "tokenizer_config.json")) else None)
    if tok_path:
        from ..data.rl_dataset import load_tokenizer

        tokenizer = load_tokenizer(tok_path)
    srv = EngineServer(eng, host=a.host, port=a.port, stream_interval=a.stream_interval, weight_update_fn=on_update,
                       tp_cmd_group=cmd_group, tokenizer=tokenizer, model_name=a.served_model_name or a.model_path,
                       tool_parser=a.tool_call_parser).start()
    if a.rollout_manager_address:
        import time as _t

        info, t0 = None, _t.time()
        while info is None:                   # elastic join: the manager may not be up yet (or restarts); keep trying for a while
            try:
                info = srv.register_with_manager(a.rollout_manager_address, local=False, handshake_port=a.transfer_agent_handshake_port)
            except OSError as e:
                if _t.time() - t0 > 600:
                    raise
                print(f"[engine] manager {a.rollout_manager_address} not reachable yet ({e}); retrying", flush=True)
                _t.sleep(1.0)
        print(f"[engine] registered with manager: {info}", flush=True)
    print(f"[engine] serving {cfg.name} on {srv.endpoint}", flush=True)
    srv.serve_until_shutdown()
    print(f"[engine] exiting: weight_version={eng.weight_version} decode_tokens={eng.stats['decode_tokens']} "
          f"prefill_tokens={eng.stats['prefill_tokens']}", flush=True)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
