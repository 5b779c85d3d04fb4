#!/bin/bash
# Verification debt from round 1 (features finished after the GPU budget ran out; see ROADMAP.md).  Each block prints PASS/FAIL and the
# script keeps going.  Needs 4 GPUs for everything, 1 GPU for the first block.  usage: gpurun --gpus 4 -- 'bash scripts/gpu_round2_checks.sh'
set -o pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { local name=$1; shift; echo "=== $name"; if timeout ${T:-600} "$@" > gpurun_out/r2_$name.log 2>&1; then echo "PASS $name"; else echo "FAIL $name (gpurun_out/r2_$name.log)"; tail -5 gpurun_out/r2_$name.log; fi; }
NG=$(python -c "import torch; print(torch.cuda.device_count())")

# 1 GPU: vision-language path on the CUDA kernels, MoE decode, activation offload
POLYRL_TEST_VLM_GPU=1 run vlm_gpu python -m pytest tests/test_gpu_e2e.py -k vlm -m gpu -x -q
run moe_gpu python scripts/gpu_moe_check.py
POLYRL_MOE_STATIC_DECODE=1 run moe_static_graph_gpu python scripts/gpu_moe_check.py
run act_offload python - <<'PY'
import torch
from polyrl_b200.models import get_model_config, pack_padded
from polyrl_b200.parallel.zero3 import Zero3Engine
cfg = get_model_config("tiny-llama-128")
ids = torch.randint(3, cfg.vocab_size, (4, 512), device="cuda")
inp, _ = pack_padded(ids, torch.ones_like(ids))
labels = torch.randint(0, cfg.vocab_size, (inp.input_ids.numel(),), device="cuda")
res = {}
for off in (False, True):
    eng = Zero3Engine(cfg, "cuda", lr=1e-3, seed=1)
    eng.act_ckpt, eng.act_offload = "none", off
    lp, ent = eng.forward_logprob_entropy(inp, labels, 1.0)
    (-lp.mean()).backward()
    res[off] = (eng.step()["grad_norm"], dict(eng.offload_stats))
print(res)
assert res[True][1]["bytes"] > 0 and abs(res[True][0] - res[False][0]) < 1e-3 * max(1.0, res[False][0])
PY

if [ "$NG" -ge 4 ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1"
  EP_TRAIN_DEVICE=cuda run ep_nccl $TR --master-port 29741 tests/ep_train_worker.py
  EP_TRAIN_DEVICE=cuda EP_PEER=1 run ep_peer $TR --master-port 29742 tests/ep_train_worker.py
  T=900 NGPU=4 NROLL=2 run disagg_2trainers_2rollout bash scripts/gpu_disagg_check.sh
  T=900 NGPU=4 NROLL=2 run disagg_tp2 bash scripts/gpu_disagg_check.sh actor_rollout_ref.rollout.tensor_model_parallel_size=2
  POLYRL_TEST_MULTIGPU_WORLD=4 T=1500 run multigpu_suite python -m pytest tests/test_multigpu.py -m gpu -x -q
fi
