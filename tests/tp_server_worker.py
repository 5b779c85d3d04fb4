"""Lock-step command channel of the TP engine server (torchrun, 2 ranks, gloo): rank 0 serves HTTP, rank 1 mirrors.
Both ranks hold full replicas here (the channel, not the TP math, is under test — tests/tp_worker.py covers that on GPUs)."""
import json
import os
import sys
import urllib.request

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def post(ep, path, body, timeout=120):
    req = urllib.request.Request(f"http://{ep}{path}", data=json.dumps(body).encode(), headers={"Content-Type": "application/json"})
    with urllib.request.urlopen(req, timeout=timeout) as r:
        return json.loads(r.read().decode())


def main():
    rank = int(os.environ["RANK"])
    dist.init_process_group("gloo")
    from polyrl_b200.models import TransformerWeights, get_model_config
    from polyrl_b200.rollout.engine import RolloutEngine
    from polyrl_b200.rollout.server import EngineServer, run_tp_follower

    cfg = get_model_config("tiny-llama")
    w = TransformerWeights(cfg, "cpu", torch.float32, seed=4)
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=8, max_model_len=96, use_cuda_graph=False, seed=9)
    eng.set_weights(w.embed, w.layers, w.head)
    seen = {}
    orig_step = eng.step

    def step():
        fin = orig_step()
        for r in fin:
            seen[r.rid] = (list(r.output_ids), r.finish_reason)
        return fin

    eng.step = step
    versions = []
    grp = dist.group.WORLD
    if rank == 0:
        srv = EngineServer(eng, port=0, weight_update_fn=lambda b: (versions.append(b.get("weight_version")) or True, "ok"),
                           tp_cmd_group=grp).start()
        ep = srv.endpoint
        sp = {"max_new_tokens": 12, "temperature": 1.0, "ignore_eos": True}
        out1 = post(ep, "/generate", {"input_ids": [[5, 6, 7, 8]] * 3, "sampling_params": sp})
        post(ep, "/update_weights_from_agent", {"weight_version": 3})
        out2 = post(ep, "/generate", {"input_ids": [9, 10, 11], "sampling_params": {**sp, "max_new_tokens": 5}})
        post(ep, "/abort_request", {"abort_all": True})
        srv.stop()
        srv._engine_thread.join(timeout=30)
        mine = {k: v for k, v in seen.items()}
        http = {o["meta_info"]["id"]: o["output_ids"] for o in out1} | {out2["meta_info"]["id"]: out2["output_ids"]}
    else:
        run_tp_follower(eng, grp, weight_update_fn=lambda b: versions.append(b.get("weight_version")))
        mine, http = {k: v for k, v in seen.items()}, None
    box = [None, None]
    dist.all_gather_object(box, (mine, versions, eng.weight_version))
    if rank == 0:
        (m0, v0, wv0), (m1, v1, wv1) = box
        same = m0 == m1 and len(m0) == 4
        http_ok = all(m0[r][0] == ids for r, ids in http.items())
        print(f"TP_SERVER_RESULTS same={int(same)} http_ok={int(http_ok)} n={len(m0)} versions={v0 == v1 == [3]} wv={wv0 == wv1 == 3}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
