"""Launch tests/multigpu_worker.py under torchrun on all visible GPUs (>= 2) and check its report."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
from _spawn import run_torchrun  # noqa: E402
import json
import os
import sys

from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def test_peer_memory_paths_match_nccl():
    n = min(torch.cuda.device_count(), int(os.environ.get("POLYRL_TEST_MULTIGPU_WORLD", "2")))   # 2 = the measured configuration
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29655", str(ROOT / "tests" / "multigpu_worker.py")], capture_output=True, text=True,
                       timeout=600, env={**os.environ, "PYTHONPATH": str(ROOT)})
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTIGPU_RESULTS ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads(line[0].split(" ", 1)[1])
    assert res["allgather_pull_exact"] and res["ag_gemm_gather_exact"] and res["weight_push_exact"]
    assert res["reduce_scatter_accum_err"] < 1e-5 and res["reduce_scatter_sumsq_err"] < 1e-3
    assert res["ag_gemm_err"] < 5e-3
    assert res["zero3_p2p_weight_err"] < 2e-2 and res["zero3_nccl_weight_err"] < 2e-2
    st = res["zero3_p2p_grad_staging"]                       # the big 2-D weights' gradients are produced in the slot, not copied
    assert st["in_place"] > 3 * st["copied"], st


def test_tensor_parallel_rollout_engine():
    n = 2
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29656", str(ROOT / "tests" / "tp_worker.py")], capture_output=True, text=True,
                       timeout=600, env={**os.environ, "PYTHONPATH": str(ROOT), "TP_TIMING": "0"})   # timing: run the worker directly
    line = [l for l in r.stdout.splitlines() if l.startswith("TP_RESULTS ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads(line[0].split(" ", 1)[1])
    assert res["lockstep_graph0"] and res["lockstep_graph1"]
    # greedy decoding on random weights diverges for good after the first near-tie flip (bf16, different reduction order):
    # measured 0.92 / 0.80 on B200 — the hard requirement is lock-step between the TP ranks, asserted above
    assert res["first8_agreement_graph0"] > 0.75 and res["first8_agreement_graph1"] > 0.75
    assert res["token_agreement_vs_single_gpu_graph0"] > 0.5 and res["token_agreement_vs_single_gpu_graph1"] > 0.5


def _torchrun(worker: str, n: int, port: int, tag: str, timeout: int = 600) -> dict:
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(ROOT / "tests" / worker)], capture_output=True, text=True,
                       timeout=timeout, env={**os.environ, "PYTHONPATH": str(ROOT)})
    line = [l for l in r.stdout.splitlines() if l.startswith(tag)]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-3000:]
    return dict(kv.split("=") for kv in line[0].split()[1:])


def test_ulysses_sp_nccl_and_peer_memory():
    """SP=2 == no SP (bf16 tolerances), and the NVLink peer-memory all-to-all == the NCCL all-to-all."""
    v = _torchrun("ulysses_worker.py", 2, 29657, "ULYSSES_RESULTS")
    assert float(v["lp_err"]) < 5e-2 and float(v["w_err"]) < 0.15      # measured 0 / 1.9e-2 (bf16 + AdamW at lr 1e-2)
    # forward log-probs are bit-identical; the updated weights go through the flash-attn backward, whose dq accumulation uses
    # atomics (run-to-run 1.5e-5 on 2x B200), so the weight comparison is tight but not exact
    assert float(v["peer_lp_err"]) < 1e-6 and float(v["peer_w_err"]) < 1e-4


def test_vocab_parallel_linear_ce_kernels():
    v = _torchrun("ce_tp_worker.py", 2, 29658, "CE_TP_RESULTS")
    assert float(v["lp"]) < 3e-2 and float(v["ent"]) < 3e-2 and float(v["dh"]) < 3e-2 and float(v["dw"]) < 3e-2


def test_hsdp_and_resident_replica():
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    v = _torchrun("hsdp_worker.py", 4, 29659, "HSDP_RESULTS")
    # bf16: saved-activation vs recompute gradients differ in rounding and AdamW (lr 1e-2, two steps) amplifies that on near-zero
    # gradients — measured on 4x B200: w_err 2.1e-3, resident_err 6.6e-2 (the fp32 gloo test pins exact equality)
    assert float(v["w_err"]) < 2e-2 and float(v["resident_err"]) < 0.2


def test_tensor_parallel_training_kernels():
    """TP=2 training on the bf16 kernel path (tcgen05 GEMMs, fused SwiGLU, vocab-parallel fused CE) vs the fp32 oracle."""
    v = _torchrun("tp_train_worker.py", 2, 29660, "TP_TRAIN_RESULTS")
    # bf16 kernels vs fp32 oracle after two AdamW steps at lr 1e-2 — measured on B200 (dp2 x tp2): lp 0.14, grad-norm 1.2e-2, w 7.9e-2
    assert float(v["lp_err"]) < 0.5 and float(v["gn_err"]) < 0.05 and float(v["w_err"]) < 0.25
