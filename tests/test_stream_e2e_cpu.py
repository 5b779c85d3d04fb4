"""BASELINE config #1 (plumbing, no GPU): GPT-2-family PPO with 1 rollout worker + 1 train worker on CPU/gloo
world_size=2 — streaming trainer, C++ manager, native engine over HTTP, broadcast weight sync every step,
mini/micro-batch boundaries, checkpoint save + resume.  Also the colocated single-process GRPO loop."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
from _spawn import run_torchrun  # noqa: E402
import os

import pytest
import sys

from pathlib import Path


ROOT = Path(__file__).resolve().parent.parent


def _run(nproc: int, overrides: list[str], timeout: int = 600):
    env = {**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "2", "POLYRL_ENGINE_PORT_BASE": "31000"}
    if nproc == 1:
        cmd = [sys.executable, "-m", "polyrl_b200.trainer.main_stream", *overrides]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
               "--master-port", "29641", "-m", "polyrl_b200.trainer.main_stream", *overrides]
    return run_torchrun(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


COMMON = ["trainer.device=cpu", "data.synthetic=True", "data.synthetic_num_samples=64", "data.max_prompt_length=24",
          "data.synthetic_prompt_len_range=[6,16]", "data.max_response_length=8", "data.train_batch_size=4",
          "actor_rollout_ref.rollout.n=2", "actor_rollout_ref.rollout.min_stream_batch_size=4",
          "actor_rollout_ref.actor.ppo_mini_batch_size=2", "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=2",
          "actor_rollout_ref.rollout.log_prob_micro_batch_size_per_gpu=4", "actor_rollout_ref.rollout.max_num_seqs=8",
          "actor_rollout_ref.actor.optim.lr=1e-3", "trainer.logger=[console]", "trainer.total_epochs=1"]


def test_gpt2_ppo_disaggregated_gloo_ws2(tmp_path):
    ov = COMMON + ["actor_rollout_ref.model.path=tiny-gpt2", "critic.model.path=tiny-gpt2", "algorithm.adv_estimator=gae",
                   "critic.ppo_micro_batch_size_per_gpu=2", "critic.ppo_mini_batch_size=2", "trainer.n_rollout_gpus=1",
                   "trainer.n_gpus_per_node=2", "trainer.total_training_steps=2", "trainer.save_freq=2",
                   f"trainer.default_local_dir={tmp_path}/ckpt", "actor_rollout_ref.rollout.rollout_manager.port=0",
                   "actor_rollout_ref.rollout.rollout_manager.max_local_gen_s=30"]
    r = _run(2, ov)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:2" in r.stdout and "actor/pg_loss" in r.stdout and "critic/vf_loss" in r.stdout
    assert (tmp_path / "ckpt" / "latest_checkpointed_iteration.txt").read_text().strip() == "2"
    assert (tmp_path / "ckpt" / "global_step_2" / "actor" / "model_world_size_1_rank_0.pt").exists()
    # resume: one more step starting from step 2
    ov2 = [o for o in ov if not o.startswith("trainer.total_training_steps")] + ["trainer.total_training_steps=3"]
    r2 = _run(2, ov2)
    assert r2.returncode == 0, r2.stdout[-3000:] + r2.stderr[-3000:]
    assert "resumed from" in r2.stdout and "step:3" in r2.stdout and "step:1 " not in r2.stdout


def test_grpo_colocated_single_process(tmp_path):
    ov = COMMON + ["actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "trainer.n_gpus_per_node=1",
                   "trainer.total_training_steps=2", "actor_rollout_ref.actor.use_kl_loss=True", "trainer.test_freq=2",
                   "data.val_batch_size=8", f"trainer.rollout_data_dir={tmp_path}/gen", f"trainer.validation_data_dir={tmp_path}/val",
                   "algorithm.filter_groups.enable=True"]
    r = _run(1, ov)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:2" in r.stdout and "actor/kl_loss" in r.stdout and "val-core/" in r.stdout and "dapo/kept_group_ratio" in r.stdout
    import json

    dumps = sorted((tmp_path / "gen").glob("train_*.jsonl"))
    assert dumps and list((tmp_path / "val").glob("val_*.jsonl"))
    row = json.loads(dumps[0].read_text().splitlines()[0])
    assert "score" in row and ("output_ids" in row or "output" in row)


def test_ulysses_sp_equivalence_gloo():
    env = {**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "2", "ULYSSES_DEVICE": "cpu"}
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29677", str(ROOT / "tests" / "ulysses_worker.py")], capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("ULYSSES_RESULTS")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    vals = dict(kv.split("=") for kv in line[0].split()[1:])
    assert float(vals["lp_err"]) < 1e-4 and float(vals["w_err"]) < 1e-3
    for fam in ("tiny-qwen2-vl", "tiny-qwen3-vl", "tiny-glm4v"):          # image rows (+ DeepStack rows) / 3-axis positions sliced with the tokens
        assert float(vals[f"vlm_lp_err[{fam}]"]) < 1e-4 and float(vals[f"vlm_w_err[{fam}]"]) < 1e-3


def test_hsdp_equivalence_gloo():
    """fsdp_size=2 x 2 replicas must match plain ZeRO-3 over 4 ranks (reference HSDP mesh: fsdp_workers.py:99-118)."""
    env = {**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "1", "HSDP_DEVICE": "cpu"}
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29688", str(ROOT / "tests" / "hsdp_worker.py")], capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("HSDP_RESULTS")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    vals = dict(kv.split("=") for kv in line[0].split()[1:])
    assert float(vals["w_err"]) < 1e-4 and float(vals["gn_err"]) < 1e-4
    assert float(vals["resident_err"]) < 1e-4          # resident replica + saved activations == ZeRO-3 + recompute
    assert float(vals["lora_resident_err"]) < 1e-4     # LoRA: frozen base gathered once, adapters refreshed every step


def test_vocab_parallel_linear_ce_gloo():
    """TP (vocab-parallel) fused linear-CE == unsharded op, values and gradients (reference TP mode: kernels.py:642-694)."""
    env = {**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "2", "CE_TP_DEVICE": "cpu"}
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29691", str(ROOT / "tests" / "ce_tp_worker.py")], capture_output=True, text=True, timeout=300, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("CE_TP_RESULTS")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    vals = dict(kv.split("=") for kv in line[0].split()[1:])
    assert all(float(v) < 1e-4 for v in vals.values()), vals


def test_disaggregated_tp2_rollout_gloo_ws3(tmp_path):
    """BASELINE config 3 in miniature: 1 trainer rank + a TP-2 rollout engine (2 ranks: HTTP front-end + lock-step follower),
    weights broadcast to both rollout ranks every step."""
    ov = COMMON + ["actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "trainer.n_rollout_gpus=2",
                   "trainer.n_gpus_per_node=3", "trainer.total_training_steps=2", "actor_rollout_ref.rollout.tensor_model_parallel_size=2",
                   "actor_rollout_ref.rollout.rollout_manager.port=0", "actor_rollout_ref.rollout.rollout_manager.max_local_gen_s=30",
                   "actor_rollout_ref.rollout.cuda_graph=False"]
    r = _run(3, ov)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:2" in r.stdout and "actor/pg_loss" in r.stdout and "training/num_rollout_instances:1" in r.stdout


def test_hybrid_local_plus_remote_instances_gloo_ws2(tmp_path):
    """PolyRL's hybrid mode: the trainer rank also serves a LOCAL engine (aliasing its weights) to the manager; generation is
    shared between the local and the remote instance, the local one is released before training."""
    ov = COMMON + ["actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "trainer.n_rollout_gpus=1",
                   "trainer.n_gpus_per_node=2", "trainer.total_training_steps=2", "trainer.local_rollout=True",
                   "actor_rollout_ref.rollout.rollout_manager.port=0", "actor_rollout_ref.rollout.rollout_manager.max_local_gen_s=30",
                   "actor_rollout_ref.rollout.cuda_graph=False", "data.train_batch_size=8"]
    r = _run(2, ov)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:2" in r.stdout and "actor/pg_loss" in r.stdout
    assert "registered local instance" in (r.stdout + r.stderr) or "local" in r.stdout


def test_tensor_parallel_training_dp2_tp2_gloo():
    """(dp=2 x tp=2) TPTrainEngine (column/row-parallel blocks, vocab-parallel fused CE, DP-averaged grads, global-norm clip,
    AdamW on local slices) tracks a single-process oracle over two optimizer steps."""
    env = {**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "1", "TP_TRAIN_DEVICE": "cpu"}
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29752", str(ROOT / "tests" / "tp_train_worker.py")], capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("TP_TRAIN_RESULTS")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    vals = dict(kv.split("=") for kv in line[0].split()[1:])
    assert vals["dp"] == "2" and vals["tp"] == "2"
    assert float(vals["lp_err"]) < 1e-3 and float(vals["gn_err"]) < 1e-4 and float(vals["w_err"]) < 2e-3
    # Megatron sequence parallelism inside the TP group (token-sharded norms / residuals, all-gather + reduce-scatter around the GEMMs)
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29753", str(ROOT / "tests" / "tp_train_worker.py")], capture_output=True, text=True, timeout=600,
                       env={**env, "TP_TRAIN_SP": "1"})
    line = [l for l in r.stdout.splitlines() if l.startswith("TP_TRAIN_RESULTS")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    vals = dict(kv.split("=") for kv in line[0].split()[1:])
    assert float(vals["lp_err"]) < 1e-3 and float(vals["gn_err"]) < 1e-4 and float(vals["w_err"]) < 2e-3
    # GLM-4 blocks (RMSNorm on the summed projection outputs, partial rotary, qkv bias) under sequence parallelism
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29754", str(ROOT / "tests" / "tp_train_worker.py")], capture_output=True, text=True, timeout=600,
                       env={**env, "TP_TRAIN_SP": "1", "TP_TRAIN_MODEL": "tiny-glm4"})
    line = [l for l in r.stdout.splitlines() if l.startswith("TP_TRAIN_RESULTS")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    vals = dict(kv.split("=") for kv in line[0].split()[1:])
    assert float(vals["lp_err"]) < 1e-3 and float(vals["gn_err"]) < 1e-4 and float(vals["w_err"]) < 2e-3


def test_one_step_off_policy_gloo_ws2(tmp_path):
    """trainer.one_step_off: batch k+1 is generated (with the weights of step k) while the trainer updates on batch k; the
    weights are pushed only after batch k+1 is complete (reference recipe: verl/recipe/one_step_off_policy)."""
    ov = COMMON + ["actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "trainer.n_rollout_gpus=1",
                   "trainer.n_gpus_per_node=2", "trainer.total_training_steps=3", "trainer.one_step_off=True",
                   "actor_rollout_ref.rollout.rollout_manager.port=0", "actor_rollout_ref.rollout.rollout_manager.max_local_gen_s=30",
                   "actor_rollout_ref.rollout.cuda_graph=False"]
    r = _run(2, ov)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:3" in r.stdout and "timing_s/wait_next_rollout" in r.stdout and "actor/pg_loss" in r.stdout


def test_colocated_ws2_bench_topology_gloo():
    env = {**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "2"}
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29771", str(ROOT / "tests" / "colocated_ws2_worker.py")], capture_output=True, text=True, timeout=600, env=env)
    lines = [l for l in r.stdout.splitlines() if l.startswith("COLO2 step")]
    assert r.returncode == 0 and len(lines) == 2, r.stdout[-2000:] + r.stderr[-2000:]
    assert "resident True" in lines[1] and "saved 2" in lines[1] and "n 16" in lines[1]


@pytest.mark.parametrize("model", ["tiny-qwen2-vl", "tiny-glm4v"])
def test_vlm_grpo_disaggregated_gloo_ws2(tmp_path, model):
    """Vision-language GRPO through the whole streaming stack on CPU: image dataset (jsonl + PNG files, byte tokenizer) ->
    manager -> remote engine that runs the ViT tower at prefill -> actor update (tower + text) with KL to the reference ->
    broadcast weight sync including the tower -> validation."""
    import json

    import numpy as np
    from PIL import Image

    rng = np.random.default_rng(0)
    rows = []
    for i in range(16):
        row = {"data_source": "hiyouga/geometry3k", "reward_model": {"ground_truth": str(i % 3)}, "extra_info": {"index": i}}
        if i % 4 != 3:
            p = tmp_path / f"im{i}.png"
            Image.fromarray(rng.integers(0, 255, size=(12 + 4 * (i % 3), 20, 3), dtype=np.uint8)).save(p)
            row["images"] = [{"path": str(p)}]
            row["prompt"] = [{"role": "user", "content": f"<image>angle {i}?"}]
        else:
            row["prompt"] = [{"role": "user", "content": f"no picture {i}"}]
        rows.append(row)
    f = tmp_path / "geo.jsonl"
    f.write_text("\n".join(json.dumps(r) for r in rows))
    ov = [o for o in COMMON if not o.startswith(("data.synthetic", "data.max_prompt_length"))] + [
        f"actor_rollout_ref.model.path={model}", "actor_rollout_ref.model.tokenizer_path=byte", f"data.train_files={f}",
        f"data.val_files={f}", "data.max_prompt_length=48", "data.image_min_pixels=64", "data.image_max_pixels=1024",
        "data.val_batch_size=8", "trainer.test_freq=2", "algorithm.adv_estimator=grpo", "actor_rollout_ref.actor.use_kl_loss=True",
        "trainer.n_rollout_gpus=1", "trainer.n_gpus_per_node=2", "trainer.total_training_steps=2",
        "actor_rollout_ref.rollout.rollout_manager.port=0", "actor_rollout_ref.rollout.rollout_manager.max_local_gen_s=30"]
    r = _run(2, ov)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:2" in r.stdout and "actor/pg_loss" in r.stdout and "actor/kl_loss" in r.stdout and "val-core/" in r.stdout


def test_pipeline_parallel_training_1f1b_gloo_ws4():
    """PipelineEngine (1F1B): pp=2 x dp=2, pp=4 with tied embeddings and pp=4 with fewer micro-batches than stages all track a
    single-process oracle (losses, grad norm, weights after two steps); in-flight micro-batches stay within pp - stage."""
    env = {**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "1"}
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29758", str(ROOT / "tests" / "pp_train_worker.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "PP_TRAIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    # the same schedules over DeepSeek-V3 blocks (latent attention, dense first block, sigmoid group-limited experts + shared expert)
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29759", str(ROOT / "tests" / "pp_train_worker.py")], capture_output=True, text=True, timeout=600,
                       env={**env, "PP_TRAIN_MODEL": "tiny-deepseek-v3"})
    assert r.returncode == 0 and "PP_TRAIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_expert_parallel_training_gloo_ws4():
    """EPTrainEngine: 8 experts over 4 ranks, variable-size dispatch / combine all-to-all (differentiable), dense part data-parallel —
    log-probs, grad norm and weights after two steps equal a single-process oracle."""
    env = {**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "1"}
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29761", str(ROOT / "tests" / "ep_train_worker.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "EP_TRAIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    # GLM-4.5-style stack: first block dense, sigmoid group-limited router with a selection bias, replicated shared experts.  The first
    # step is exact (1e-6); AdamW's sign-like first update then turns summation-order noise on near-zero gradient elements into +-lr
    # weight differences (one flipped embedding element = 4e-3 relative), hence the wider bounds on the second step.
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29762", str(ROOT / "tests" / "ep_train_worker.py")], capture_output=True, text=True, timeout=600,
                       env={**env, "EP_TRAIN_MODEL": "tiny-glm4-moe", "EP_TRAIN_TOL": "2e-3,1e-4,2e-2"})
    assert r.returncode == 0 and "EP_TRAIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    # DeepSeek-V3 architecture: the same routing on top of multi-head latent attention (models/mla.py)
    r = run_torchrun([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
                        "--master-port", "29763", str(ROOT / "tests" / "ep_train_worker.py")], capture_output=True, text=True, timeout=600,
                       env={**env, "EP_TRAIN_MODEL": "tiny-deepseek-v3", "EP_TRAIN_TOL": "2e-3,1e-4,2e-2"})
    assert r.returncode == 0 and "EP_TRAIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_elastic_external_instance_joins_a_running_job(tmp_path):
    """PolyRL's elastic deployment on CPU: the job has NO in-job rollout ranks (``trainer.elastic_rollout``): the trainer serves a
    hybrid local engine behind the manager; a stand-alone ``launch_server`` process started separately keeps retrying until the
    manager exists, registers as a remote instance, receives every new policy over TCP and takes part in generation."""
    import subprocess
    import time

    mport = 32611
    env = {**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "2", "POLYRL_ENGINE_PORT_BASE": "31400"}
    ext = subprocess.Popen([sys.executable, "-m", "polyrl_b200.rollout.launch_server", "--model-path", "tiny-llama", "--device", "cpu", "--port", "32621",
                            "--max-num-seqs", "8", "--max-model-len", "64", "--load-format", "dummy", "--enable-weight-transfer-agent",
                            "--transfer-agent-handshake-port", "32631", "--weight-receiver-zmq-bind-host", "127.0.0.1",
                            "--rollout-manager-address", f"127.0.0.1:{mport}"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=ROOT)
    try:
        time.sleep(3.0)                                   # the engine is up and retrying before the job (and its manager) exists
        ov = COMMON + ["actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "trainer.n_gpus_per_node=2",
                       "trainer.elastic_rollout=True", "trainer.total_training_steps=6", f"actor_rollout_ref.rollout.rollout_manager.port={mport}",
                       "actor_rollout_ref.rollout.rollout_manager.max_local_gen_s=1", "data.train_batch_size=8",
                       "actor_rollout_ref.rollout.min_stream_batch_size=4", "actor_rollout_ref.actor.ppo_mini_batch_size=4"]
        r = _run(2, ov)                                     # two trainer ranks: rank 0 owns the manager connection, both serve engines
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert "step:6" in r.stdout and "actor/pg_loss" in r.stdout
        # the job is gone (and its manager with it): stop the external engine and read what it saw
        import urllib.request

        urllib.request.urlopen(urllib.request.Request("http://127.0.0.1:32621/shutdown", method="POST", data=b"{}",
                                                      headers={"Content-Type": "application/json"}), timeout=30).read()
        out, _ = ext.communicate(timeout=60)
        assert "registered with manager" in out, out[-2000:]
        tail = [l for l in out.splitlines() if l.startswith("[engine] exiting")]
        assert tail, out[-2000:]
        stats = dict(kv.split("=") for kv in tail[0].split()[2:])
        assert int(stats["weight_version"]) >= 1, out[-2000:]          # at least one policy arrived over TCP
    finally:
        if ext.poll() is None:
            ext.terminate()


def test_disaggregated_two_trainers_one_rollout_gloo_ws3():
    """More than one trainer rank in a disaggregated job: rank 0 streams from the manager and broadcasts, the other trainer ranks
    must follow the remote path too (they hold no manager connection)."""
    ov = COMMON + ["actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "trainer.n_rollout_gpus=1",
                   "trainer.n_gpus_per_node=3", "trainer.total_training_steps=2", "actor_rollout_ref.rollout.rollout_manager.port=0",
                   "actor_rollout_ref.rollout.rollout_manager.max_local_gen_s=30", "data.train_batch_size=8",
                   "actor_rollout_ref.actor.ppo_mini_batch_size=4"]
    r = _run(3, ov)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:2" in r.stdout and "actor/pg_loss" in r.stdout
    r = _run(3, ov + ["trainer.one_step_off=True", "trainer.total_training_steps=3"])      # the prefetched batch reaches both trainers
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:3" in r.stdout


def test_model_based_reward_and_val_generation_log_gloo_ws2(tmp_path):
    """reward_model.enable in the streaming trainer: every trainer rank scores its slice with the value-head reward model, the slices
    are all-gathered and drive GRPO; trainer.log_val_generations writes a few validation samples per validation pass."""
    ov = COMMON + ["actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "trainer.n_gpus_per_node=2",
                   "trainer.total_training_steps=1", "reward_model.enable=True", "reward_model.model.path=tiny-llama",
                   "reward_model.micro_batch_size_per_gpu=4", "trainer.test_freq=1", "data.val_batch_size=8", "trainer.log_val_generations=3",
                   "trainer.project_name=t", f"trainer.experiment_name=rm{os.getpid()}"]
    r = _run(2, ov)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "step:1" in r.stdout and "critic/score/mean" in r.stdout and "val-core/" in r.stdout
    logs = list((ROOT / "outputs" / "val_generations" / "t" / f"rm{os.getpid()}").glob("*.jsonl")) if (ROOT / "outputs").exists() else []
    logs += list(__import__("pathlib").Path("outputs/val_generations/t").glob(f"rm{os.getpid()}/*.jsonl"))
    assert logs and len(logs[0].read_text().splitlines()) == 3
    import shutil

    shutil.rmtree(logs[0].parent, ignore_errors=True)
