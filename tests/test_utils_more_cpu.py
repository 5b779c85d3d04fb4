"""Direct unit tests for utility modules that were only covered end to end: FLOPs counter, checkpoint merger (incl. HF export),
tracking back-ends, profiler timers, TCP weight-transfer engine (loopback), sync trainer entry point."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_flops_counter_matches_6nd_rule():
    from polyrl_b200.models import get_model_config
    from polyrl_b200.utils.flops_counter import DEVICE_PEAK_BF16, FlopsCounter

    cfg = get_model_config("llama3-8b")
    fc = FlopsCounter(cfg)
    n_dense = fc._dense_params()
    assert 7.4e9 < n_dense < 7.6e9                              # 8.03 B minus the 0.53 B input embedding
    tf_s, peak = fc.estimate_flops([512] * 32, delta_time=1.0)
    tokens = 512 * 32
    assert abs(tf_s * 1e12 - (6 * n_dense * tokens + 12 * cfg.num_layers * cfg.q_size * 32 * 512 * 512)) / (tf_s * 1e12) < 1e-9
    fwd, _ = fc.estimate_flops([512] * 32, 1.0, forward_only=True)
    assert abs(fwd * 3 - tf_s) / tf_s < 1e-9
    assert DEVICE_PEAK_BF16["B200"] == 2.25e15
    # vision-language: the tower's FLOPs are added per image (patch count before the merge); the text part is unchanged
    vl = get_model_config("qwen2-vl-7b") if "qwen2-vl-7b" in __import__("polyrl_b200.models.config", fromlist=["PRESETS"]).PRESETS else None
    if vl is not None:
        fv = FlopsCounter(vl)
        base, _ = fv.estimate_flops([1024], 1.0)
        with_img, _ = fv.estimate_flops([1024], 1.0, images_seqlens=[1024])
        vit = (with_img - base) * 1e12
        assert 6 * 0.6e9 * 1024 < vit < 6 * 0.75e9 * 1024 * 1.3          # ~0.67 B-parameter tower + its attention
    # sliding-window layers attend to at most `window` keys
    import dataclasses

    sw = dataclasses.replace(cfg) if dataclasses.is_dataclass(cfg) else cfg
    object.__setattr__(sw, "sliding_window", 128)
    object.__setattr__(sw, "max_window_layers", 0)
    short, _ = FlopsCounter(sw).estimate_flops([512] * 32, 1.0)
    assert short < tf_s and abs((tf_s - short) * 1e12 - 12 * cfg.num_layers * cfg.q_size * 32 * (512 * 512 - 512 * 128)) / (tf_s * 1e12) < 1e-9


def test_flops_counter_latent_attention_and_plain_mlp_families():
    """FLOPs-relevant parameters = all parameters minus the input embedding, the norm gains / activation scalars and the idle experts:
    holds for the multi-head-latent-attention MoE stacks (DeepSeek-V3, V2-Lite, Moonlight) and the non-gated xIELU MLP (Apertus);
    MLA attention FLOPs count nope + rope query / key features and v_head_dim value features."""
    from polyrl_b200.models import get_model_config
    from polyrl_b200.utils.flops_counter import FlopsCounter

    for name in ("deepseek-v3", "deepseek-v2-lite", "moonlight-16b-a3b", "apertus-8b"):
        c = get_model_config(name)
        fc = FlopsCounter(c)
        small = c.num_layers * 2 * c.hidden_size + c.hidden_size                     # block norms + final norm
        if c.attn_type == "mla":
            small += c.num_layers * (c.kv_lora_rank + c.q_lora_rank)                    # latent norm gains
        if c.mlp_act == "xielu":
            small += 2 * c.num_layers
        if c.qk_norm:
            small += 2 * c.head_dim * c.num_layers                                      # per-head q / k norm gains
        if c.moe_router == "sigmoid_group":
            small += c.num_sparse_layers() * c.num_experts                              # selection-only score bias
        want = c.num_active_params() - c.vocab_size * c.hidden_size - small
        assert abs(fc._dense_params() - want) / want < 1e-9, name
        if c.attn_type == "mla":
            a, _ = fc.estimate_flops([1024], 1.0)
            b, _ = fc.estimate_flops([2048], 1.0)
            quad = ((b - 2 * a) * 1e12) / (12.0 * c.num_layers * (2048 ** 2 - 2 * 1024 ** 2))
            assert abs(quad - c.num_heads * 0.5 * (c.head_dim + c.v_head_dim)) / quad < 1e-6
    v3 = get_model_config("deepseek-v3")
    assert 6.6e11 < v3.num_params() < 6.9e11 and 3.6e10 < v3.num_active_params() < 3.9e10          # 671 B total / 37 B active
    assert get_model_config("deepseek-v2-lite").kv_cache_layout() == (1, 576, False)


def test_model_merger_roundtrip_and_hf_names(tmp_path):
    from polyrl_b200.models import get_model_config
    from polyrl_b200.models import transformer as tf
    from polyrl_b200.parallel.zero3 import Zero3Engine
    from polyrl_b200.utils.checkpoint import ShardedCheckpointManager
    from polyrl_b200.utils.model_merger import merge

    cfg = get_model_config("tiny-llama")
    eng = Zero3Engine(cfg, "cpu", seed=9)
    ShardedCheckpointManager(eng, cfg).save(str(tmp_path / "actor"), global_step=1)
    mcfg, sd = merge(str(tmp_path / "actor"))
    want = dict(eng.named_full_tensors())
    assert mcfg.hidden_size == cfg.hidden_size and set(sd) == set(want)
    assert all(torch.equal(sd[k].float(), want[k].float()) for k in sd)
    hf = tf.export_hf_state_dict(cfg, sd)
    assert "model.layers.0.self_attn.q_proj.weight" in hf and "model.embed_tokens.weight" in hf
    back = tf.convert_hf_state_dict(cfg, hf)                        # HF -> ours restores the fused / interleaved layout
    assert all(torch.equal(back[k].float(), sd[k].float()) for k in sd)
    r = subprocess.run([sys.executable, "-m", "polyrl_b200.utils.model_merger", "--ckpt", str(tmp_path / "actor"), "--out",
                        str(tmp_path / "hf"), "--hf"], capture_output=True, text=True, env={**os.environ, "PYTHONPATH": str(ROOT)})
    assert r.returncode == 0 and (tmp_path / "hf" / "pytorch_model.bin").exists(), r.stderr[-1000:]


def test_tracking_backends_and_validation_logger(tmp_path, monkeypatch):
    from polyrl_b200.utils.tracking import Tracking, ValidationGenerationsLogger

    monkeypatch.chdir(tmp_path)
    t = Tracking("proj", "exp", ["console", "file"], config={"a": 1})
    t.log({"actor/pg_loss": torch.tensor(0.5), "n": 3, "s": "x"}, step=1)
    t.log({"actor/pg_loss": 0.25}, step=2)
    t.finish()
    files = list(tmp_path.rglob("*.jsonl"))
    assert files, list(tmp_path.rglob("*"))
    rows = [json.loads(l) for l in files[0].read_text().splitlines()]
    assert rows[0]["step"] == 1 and abs(rows[0]["actor/pg_loss"] - 0.5) < 1e-9 and rows[0]["s"] == "x" and rows[1]["step"] == 2
    ValidationGenerationsLogger("proj", "exp", root=str(tmp_path / "val")).log(["file"], [("q", "a", 1.0)], step=3)
    assert list((tmp_path / "val").rglob("*"))
    try:
        Tracking("proj", "exp", ["definitely_not_a_backend"])
        assert False
    except (ValueError, AssertionError, ImportError):
        pass


def test_tracking_external_backends_through_fake_clients(tmp_path, monkeypatch):
    """The wandb / trackio / mlflow / swanlab / clearml adapters and the per-backend validation-generation tables, driven against
    recording stand-ins for the client libraries (none is installed here): init arguments, metric routing, mlflow key sanitisation
    and flattened params, clearml title / series split, the wandb table that grows by one row per step; a backend whose library is
    missing is skipped without stopping the run."""
    import sys
    import types

    from polyrl_b200.utils import tracking as T

    calls: list = []

    def rec(name):
        return lambda *a, **k: calls.append((name, a, k))

    wandb = types.ModuleType("wandb")
    wandb.init, wandb.log, wandb.finish = rec("wandb.init"), rec("wandb.log"), rec("wandb.finish")
    wandb.Settings = lambda **k: ("settings", k)
    wandb.Table = lambda columns, data: {"columns": columns, "data": data}
    trackio = types.ModuleType("trackio")
    trackio.init, trackio.log, trackio.finish = rec("trackio.init"), rec("trackio.log"), rec("trackio.finish")
    mlflow = types.ModuleType("mlflow")
    for f in ("set_tracking_uri", "start_run", "log_params", "log_metrics", "end_run", "log_text"):
        setattr(mlflow, f, rec("mlflow." + f))
    mlflow.set_experiment = lambda name: types.SimpleNamespace(experiment_id="17")
    swanlab = types.ModuleType("swanlab")
    for f in ("login", "init", "log", "finish"):
        setattr(swanlab, f, rec("swanlab." + f))
    swanlab.Text = lambda text, caption=None: ("text", caption, text)

    class FakeLogger:
        report_scalar, report_text, report_table = (staticmethod(rec("clearml.scalar")), staticmethod(rec("clearml.text")),
                                                    staticmethod(rec("clearml.table")))

    class FakeTask:
        _cur = None

        @classmethod
        def init(cls, **k):
            calls.append(("clearml.init", (), k))
            cls._cur = cls()
            return cls._cur

        @classmethod
        def current_task(cls):
            return cls._cur

        def connect_configuration(self, cfg, name=None):
            calls.append(("clearml.connect", (cfg,), {"name": name}))

        def get_logger(self):
            return FakeLogger()

        def close(self):
            calls.append(("clearml.close", (), {}))

    clearml = types.ModuleType("clearml")
    clearml.Task = FakeTask
    for name, mod in (("wandb", wandb), ("trackio", trackio), ("mlflow", mlflow), ("swanlab", swanlab), ("clearml", clearml)):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.setenv("WANDB_ENTITY", "team")
    monkeypatch.setenv("SWANLAB_API_KEY", "k")
    monkeypatch.setenv("SWANLAB_MODE", "local")
    monkeypatch.delenv("MLFLOW_TRACKING_URI", raising=False)
    monkeypatch.chdir(tmp_path)
    cfg = {"trainer": {"wandb_proxy": "http://p:1", "n_gpus": 8}, "data": {"files": ["a", "b"]}}
    t = T.Tracking("proj", "exp", ["tracking", "trackio", "mlflow", "swanlab", "clearml", "vemlp_wandb"], config=cfg)
    assert set(t.loggers) == {"wandb", "trackio", "mlflow", "swanlab", "clearml"}          # volcengine client missing: skipped
    by = lambda n: [c for c in calls if c[0] == n]                                            # noqa: E731
    assert by("wandb.init")[0][2] == {"project": "proj", "name": "exp", "config": cfg, "entity": "team", "settings": ("settings", {"https_proxy": "http://p:1"})}
    assert by("mlflow.set_tracking_uri")[0][1] == ("sqlite:////tmp/mlruns.db",) and by("mlflow.start_run")[0][2] == {"experiment_id": "17", "run_name": "exp"}
    assert by("mlflow.log_params")[0][1][0] == {"trainer/wandb_proxy": "http://p:1", "trainer/n_gpus": 8, "data/files/list_len": 2, "data/files/0": "a", "data/files/1": "b"}
    assert by("swanlab.login")[0][1] == ("k",) and by("swanlab.init")[0][2]["mode"] == "local" and by("swanlab.init")[0][2]["config"]["FRAMEWORK"] == "polyrl_b200"
    assert by("clearml.init")[0][2]["task_name"] == "exp" and by("clearml.connect")[0][1][0] == cfg
    t.log({"val-core/gsm8k/acc/mean@4": 0.5, "actor/pg_loss": torch.tensor(0.25), "note": "hi"}, step=3)
    assert by("wandb.log")[0][2] == {"step": 3} and by("trackio.log")[0][1][0]["note"] == "hi"
    assert by("mlflow.log_metrics")[0][2] == {"metrics": {"val-core/gsm8k/acc/mean_at_4": 0.5, "actor/pg_loss": 0.25}, "step": 3}
    sc = {(c[2]["title"], c[2]["series"]): c[2]["value"] for c in by("clearml.scalar")}
    assert sc == {("val-core", "gsm8k/acc/mean@4"): 0.5, ("actor", "pg_loss"): 0.25} and "note = hi" in by("clearml.text")[0][1][0]
    t.log({"x": 1.0}, step=4, backend=["mlflow"])
    assert len(by("mlflow.log_metrics")) == 2 and len(by("wandb.log")) == 1
    # validation generations in every backend's own form
    vg = T.ValidationGenerationsLogger("proj", "exp", root=str(tmp_path / "val"))
    T.ValidationGenerationsLogger._wandb_rows.clear()
    vg.log(["wandb", "swanlab", "mlflow", "clearml", "tensorboard", "vemlp_wandb"], [("q1", "a1", 1.0), ("q2", "a2", 0.0)], step=5)
    vg.log(["wandb"], [("q3", "a3", 0.5)], step=6)
    tables = [c[1][0]["val/generations"] for c in by("wandb.log")[1:]]
    assert tables[0]["columns"] == ["step", "input_1", "output_1", "score_1", "input_2", "output_2", "score_2"] and tables[0]["data"] == [[5, "q1", "a1", 1.0, "q2", "a2", 0.0]]
    assert tables[1]["data"] == [[5, "q1", "a1", 1.0, "q2", "a2", 0.0], [6, "q3", "a3", 0.5, None, None, None]]       # one row per step
    assert [c[1] for c in by("swanlab.log")[-1][1][0]["val/generations"]] == ["sample 1", "sample 2"]
    assert json.loads(by("mlflow.log_text")[0][1][0])[1] == {"input": "q2", "output": "a2", "score": 0.0} and by("mlflow.log_text")[0][1][1] == "val_step5.json"
    assert list(by("clearml.table")[0][2]["table_plot"]["output"]) == ["a1", "a2"]
    assert (tmp_path / "val" / "proj" / "exp" / "5.jsonl").exists() and list(tmp_path.rglob("events.out.tfevents*"))
    t.finish()
    assert by("wandb.finish") and by("mlflow.end_run") and by("swanlab.finish") and by("clearml.close") and not t.loggers
    assert T.sanitize_mlflow_key("a b:c/d-e.f(g)@2") == "a b:c/d-e.f_g__at_2"


def test_profiler_timers_and_annotate():
    from polyrl_b200.utils import profiler as P

    timing: dict = {}
    with P.marked_timer("gen", timing):
        time.sleep(0.01)
    with P.marked_timer("gen", timing):
        time.sleep(0.01)
    assert timing["gen"] >= 0.02                                   # accumulates across uses

    @P.annotate("unit")
    def f(x):
        return x + 1

    assert f(1) == 2
    assert isinstance(P.log_gpu_memory_usage("here"), str)


def test_tcp_transfer_engine_loopback():
    from polyrl_b200.weight_transfer.tcp_engine import StagingBuffer, TCPWeightReceiver, TCPWeightSender, flat_state_nbytes

    torch.manual_seed(0)
    src = [("a", torch.randn(1000, 33)), ("b", torch.randn(7).to(torch.bfloat16)), ("c", torch.arange(5000, dtype=torch.int32))]
    dst = [(n, torch.zeros_like(t)) for n, t in src]
    n = flat_state_nbytes(src)
    port = 24871
    rx = TCPWeightReceiver(dst, bind_host="127.0.0.1", port=port, num_streams=3)
    tx = TCPWeightSender(num_streams=3)
    st = StagingBuffer(n, name=f"polyrl_test_{os.getpid()}")
    try:
        layout = tx.stage(src, st)
        assert sum(l[2] for l in layout) == n
        bid = tx.send(st, "127.0.0.1", port, nbytes=n)              # returns the batch id; status table records the outcome
        assert tx.status[bid] == "success" and rx.wait_and_apply(timeout=30)
        assert all(torch.equal(a, b) for (_, a), (_, b) in zip(src, dst))
        # second transfer re-uses the receiver
        src[0][1].add_(1.0)
        tx.stage(src, st)
        tx.send(st, "127.0.0.1", port, nbytes=n)
        assert rx.wait_and_apply(timeout=30) and torch.equal(src[0][1], dst[0][1])
        # a partly delivered transfer (one of three streams) must not be completed by the streams of the next one,
        # and a stream with the wrong token is refused
        before = [t.clone() for _, t in dst]
        per = ((n + 2) // 3 + 4095) // 4096 * 4096
        tx._send_range("127.0.0.1", port, st.fd, 0, min(per, n), False, st.mm, tid=12345, total=n)
        src[0][1].add_(1.0)
        tx.stage(src, st)
        tx.send(st, "127.0.0.1", port, nbytes=n)
        assert rx.wait_and_apply(timeout=30)
        assert all(torch.equal(a, b) for (_, a), (_, b) in zip(src, dst)) and not torch.equal(before[0], dst[0][1])
        bad = TCPWeightSender(num_streams=3, token="not-the-secret")
        with pytest.raises(ConnectionError):
            bad.send(st, "127.0.0.1", port, nbytes=n)
        with pytest.raises(TimeoutError):
            rx.wait_and_apply(timeout=0.5)
    finally:
        st.close()
        rx.close()


def test_sync_trainer_entry_main_ppo(tmp_path):
    ov = ["trainer.device=cpu", "data.synthetic=True", "data.synthetic_num_samples=32", "data.max_prompt_length=16",
          "data.synthetic_prompt_len_range=[4,12]", "data.max_response_length=6", "data.train_batch_size=4",
          "actor_rollout_ref.rollout.n=2", "actor_rollout_ref.actor.ppo_mini_batch_size=4", "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=4",
          "actor_rollout_ref.rollout.log_prob_micro_batch_size_per_gpu=8", "actor_rollout_ref.rollout.max_num_seqs=8",
          "actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "trainer.n_gpus_per_node=1",
          "trainer.total_training_steps=2", "trainer.logger=[console]", "trainer.total_epochs=1", "actor_rollout_ref.rollout.cuda_graph=False"]
    r = subprocess.run([sys.executable, "-m", "polyrl_b200.trainer.main_ppo", *ov], capture_output=True, text=True, timeout=600,
                       env={**os.environ, "PYTHONPATH": str(ROOT), "OMP_NUM_THREADS": "2"}, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "step:2" in r.stdout and "actor/pg_loss" in r.stdout


def test_sft_trainer_loss_matches_reference_and_decreases(tmp_path):
    from polyrl_b200.data.rl_dataset import ByteTokenizer
    from polyrl_b200.models import TransformerWeights, forward_hidden, lm_head_weight
    from polyrl_b200.trainer.sft import SFTConfig, SFTDataset, SFTTrainer

    rows = [{"prompt": f"q{i}: what is {i}+{i}?", "response": f" it is {2 * i}"} for i in range(16)]
    f = tmp_path / "sft.jsonl"
    f.write_text("\n".join(json.dumps(r) for r in rows))
    cfg = SFTConfig()
    cfg.trainer.device, cfg.trainer.total_steps, cfg.trainer.total_epochs = "cpu", 6, 3
    cfg.data.train_files, cfg.data.max_length, cfg.data.train_batch_size, cfg.data.micro_batch_size_per_gpu = str(f), 48, 8, 4
    cfg.optim.lr, cfg.optim.weight_decay = 5e-3, 0.0
    tok = ByteTokenizer()
    ds = SFTDataset(str(f), tok, cfg.data)
    s0 = ds[0]
    n_resp = len(tok.encode(rows[0]["response"])) + 1
    assert int(s0["loss_mask"].sum()) == n_resp and int(s0["attention_mask"].sum()) == len(tok.encode(rows[0]["prompt"])) + n_resp
    tr = SFTTrainer(cfg, device="cpu")
    # loss parity with a plain PyTorch forward on the same weights (padded batch, shifted labels, response-only mask)
    batch = {k: torch.stack([ds[i][k] for i in range(4)]) for k in s0}
    w = TransformerWeights(tr.model_cfg, "cpu", torch.float32, init=False)
    w.load_state_dict(dict(tr.engine.named_full_tensors()))
    with torch.no_grad():
        mine = float(tr._loss_on(batch, 1.0))
        from polyrl_b200.models import pack_padded, unpack_to_padded

        inp, idx = pack_padded(batch["input_ids"], batch["attention_mask"])
        h = forward_hidden(tr.model_cfg, w, inp)
        logits = unpack_to_padded(h @ lm_head_weight(tr.model_cfg, w.head, w.embed).t(), idx, 4, 48)
        lp = torch.log_softmax(logits[:, :-1].float(), -1).gather(-1, batch["input_ids"][:, 1:, None]).squeeze(-1)
        want = float(-(lp * batch["loss_mask"][:, 1:]).sum())
    assert abs(mine - want) / abs(want) < 1e-4, (mine, want)
    before = tr.evaluate(ds)
    out = tr.fit(ds)
    after = tr.evaluate(ds)
    assert tr.step_idx == 6 and after < before - 0.05 and out["train/grad_norm"] > 0


def test_reward_loop_async_scoring_and_generative_reward_model():
    """Async reward loop: per-sample scoring with bounded concurrency (sync scorers in the executor, async scorers awaited), DAPO
    over-long shaping, and a generative reward model judged by an engine server over HTTP."""
    import threading
    import time

    import numpy as np

    from polyrl_b200.data.reward_loop import GenerativeRewardModel, RewardLoopManager, get_reward_loop_cls
    from polyrl_b200.data.rl_dataset import ByteTokenizer
    from polyrl_b200.protocol import DataProto

    tok = ByteTokenizer()
    texts = ["the answer is #### 4", "no idea", "#### 4", "x" * 30]
    R = 32
    resp = torch.zeros(4, R, dtype=torch.long)
    mask = torch.zeros(4, R, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = tok.encode(t)
        resp[i, : len(ids)] = torch.tensor(ids)
        mask[i, : len(ids)] = 1
    dp = DataProto.from_dict({"responses": resp, "response_mask": mask},
                             non_tensors={"data_source": np.array(["openai/gsm8k"] * 4, dtype=object),
                                          "reward_model": np.array([{"ground_truth": "4"}] * 4, dtype=object),
                                          "extra_info": np.array([{"question": f"q{i}"} for i in range(4)], dtype=object)})
    active, peak = [0], [0]
    lock = threading.Lock()

    def slow_score(data_source, solution_str, ground_truth, extra_info=None):
        with lock:
            active[0] += 1
            peak[0] = max(peak[0], active[0])
        time.sleep(0.05)
        with lock:
            active[0] -= 1
        return {"score": float("#### 4" in solution_str), "acc": "#### 4" in solution_str}

    loop = RewardLoopManager("naive", tok, compute_score=slow_score, max_concurrency=3)
    out = loop(dp, return_dict=True)
    assert out["reward_tensor"].sum(-1).tolist() == [1.0, 0.0, 1.0, 0.0] and out["reward_extra_info"]["acc"] == [True, False, True, False]
    assert out["reward_tensor"][0, len(tok.encode(texts[0])) - 1] == 1.0
    assert 2 <= peak[0] <= 3                                                  # overlapped, bounded by max_concurrency
    dapo = RewardLoopManager("dapo", tok, compute_score=slow_score, max_resp_len=32, overlong_buffer_len=8, penalty_factor=1.0)
    o2 = dapo(dp, return_dict=True)
    assert o2["reward_tensor"][3].sum().item() == pytest.approx(-(30 - 24) / 8) and o2["reward_extra_info"]["overlong"] == [False, False, False, True]
    with pytest.raises(ValueError):
        get_reward_loop_cls("nope")

    # generative reward model over the engine's HTTP front-end: the judge text is whatever the tiny random model emits; the parse
    # function sees exactly the text the engine generated
    from polyrl_b200.models import TransformerWeights, get_model_config
    from polyrl_b200.rollout.engine import RolloutEngine
    from polyrl_b200.rollout.server import EngineServer

    cfg = get_model_config("tiny-llama")
    w = TransformerWeights(cfg, seed=3)
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=8, max_model_len=256, kv_cache_tokens=8 * 256)
    eng.set_weights(w.embed, w.layers, w.head)
    srv = EngineServer(eng, port=0).start()
    try:
        seen = []
        rm = GenerativeRewardModel([srv.endpoint], tok, parse=lambda text: (seen.append(text), float(len(text) % 2))[1], max_new_tokens=6)
        loop = RewardLoopManager("naive", tok, compute_score=rm.compute_score, max_concurrency=4)
        o3 = loop(dp, return_dict=True)
        assert len(seen) == 4 and len(o3["reward_extra_info"]["judge_text"]) == 4
        assert set(o3["reward_tensor"].sum(-1).tolist()) <= {0.0, 1.0}
    finally:
        srv.stop()


def test_dynamic_dataset_and_curriculum_sampler(tmp_path):
    import json

    import numpy as np

    from polyrl_b200.data.dynamic import DynamicGenDataset, MockDataGenerator, SuccessRateCurriculumSampler, load_extern_type
    from polyrl_b200.data.rl_dataset import ByteTokenizer
    from polyrl_b200.protocol import DataProto

    f = tmp_path / "d.jsonl"
    f.write_text("\n".join(json.dumps({"prompt": f"p{i}", "reward_model": {"ground_truth": str(i)}, "extra_info": {"index": i}}) for i in range(6)))
    ds = DynamicGenDataset(f, ByteTokenizer(), max_prompt_length=16, generator=MockDataGenerator(), max_rows=8)
    for _ in range(4):
        ds.on_batch_end(None)
    assert len(ds) == 8 and ds.generated == 2 and ds[7]["raw_prompt_ids"] == ds[0]["raw_prompt_ids"]
    assert load_extern_type(None, "MockDataGenerator") is MockDataGenerator
    gen_py = tmp_path / "gen.py"
    gen_py.write_text("from polyrl_b200.data.dynamic import AbstractDataGenerator\n"
                      "class Hard(AbstractDataGenerator):\n"
                      "    def generate(self, dataset, batch):\n"
                      "        return [{'prompt': 'harder ' + str(len(dataset)), 'reward_model': {'ground_truth': '0'}}]\n")
    ds2 = DynamicGenDataset(f, ByteTokenizer(), max_prompt_length=16, generator=load_extern_type(str(gen_py), "Hard")({}))
    ds2.on_batch_end(None)
    assert len(ds2) == 7 and ByteTokenizer().decode(ds2[6]["raw_prompt_ids"]) == "harder 6"

    smp = SuccessRateCurriculumSampler(range(6), seed=3)
    # prompts 0/1 always solved, 2/3 never, 4/5 half the time
    scores = torch.zeros(12, 4)
    idx = np.repeat(np.arange(6), 2)
    for k, i in enumerate(idx):
        scores[k, 0] = 1.0 if i < 2 or (i >= 4 and k % 2 == 0) else 0.0
    fb = DataProto.from_dict({"token_level_scores": scores}, non_tensors={"index": np.array(list(idx), dtype=object)})
    for _ in range(12):
        smp.update(fb)
    w = smp._weights()
    assert w[4] > 3 * w[0] and w[5] > 3 * w[2] and abs(w.sum() - 1) < 1e-9
    draws = [i for _ in range(40) for i in smp]
    assert sum(d >= 4 for d in draws) > 0.5 * len(draws)
    sd = smp.state_dict()
    s2 = SuccessRateCurriculumSampler(range(6))
    s2.load_state_dict(sd)
    assert np.allclose(s2._weights(), w) and s2.epoch == smp.epoch


def test_multiturn_sft_and_pairwise_rm_datasets(tmp_path):
    """Chat transcripts: only assistant turns carry loss (not the tokens that announce them); pairwise RM rows: [2, L] chosen/rejected."""
    from polyrl_b200.data.rl_dataset import ByteTokenizer
    from polyrl_b200.trainer.sft import MultiTurnSFTDataset, PairwiseRMDataset, SFTDataConfig, main

    tok = ByteTokenizer()
    msgs = [{"role": "system", "content": "be brief"}, {"role": "user", "content": "2+2?"}, {"role": "assistant", "content": "4"},
            {"role": "user", "content": "and 3+3?"}, {"role": "assistant", "content": "6!"}]
    f = tmp_path / "chat.jsonl"
    f.write_text("\n".join(json.dumps({"messages": msgs}) for _ in range(8)))
    ds = MultiTurnSFTDataset(f, tok, SFTDataConfig(max_length=96))
    it = ds[0]
    ids, mask, attn = it["input_ids"], it["loss_mask"].bool(), it["attention_mask"].bool()
    assert tok.decode(ids[attn].tolist()) == tok.apply_chat_template(msgs, add_generation_prompt=False)
    assert tok.decode(ids[mask].tolist()) == "4\n6!\n"                        # assistant contents + their end-of-turn, nothing else
    assert not (mask & ~attn).any()
    with pytest.raises(ValueError):
        MultiTurnSFTDataset(f, tok, SFTDataConfig(max_length=8, truncation="error"))[0]
    short = MultiTurnSFTDataset(f, tok, SFTDataConfig(max_length=40))[0]
    assert int(short["attention_mask"].sum()) == 40
    g = tmp_path / "pairs.jsonl"
    g.write_text(json.dumps({"prompt": [{"role": "user", "content": "hi"}], "chosen": "hello there", "rejected": "go away"}))
    rm = PairwiseRMDataset(g, tok, max_length=48)[0]
    assert rm["input_ids"].shape == (2, 48)
    c, r = (tok.decode(rm["input_ids"][j][rm["attention_mask"][j].bool()].tolist()) for j in (0, 1))
    assert c.endswith("hello there") and r.endswith("go away") and c[: c.index("hello")] == r[: r.index("go away")]
    # the SFT entry point trains on the transcripts
    assert main(["trainer.device=cpu", f"data.train_files={f}", "data.multiturn_messages_key=messages", "data.max_length=96", "model.path=tiny-llama",
                 "data.train_batch_size=4", "data.micro_batch_size_per_gpu=2", "trainer.total_steps=2", "optim.lr=1e-3"]) == 0


def test_tcp_receiver_recovers_from_an_incomplete_transfer():
    """A transfer that stops half way times out, and the next complete transfer is applied correctly."""
    import socket

    from polyrl_b200.weight_transfer.tcp_engine import HEADER, StagingBuffer, TCPWeightReceiver, TCPWeightSender, token_digest

    dst = [("a", torch.zeros(1000)), ("b", torch.zeros(24))]
    port = 32701
    rx = TCPWeightReceiver(dst, bind_host="127.0.0.1", port=port, num_streams=1)
    try:
        s = socket.create_connection(("127.0.0.1", port))
        s.sendall(HEADER.pack(token_digest(None), 7, rx.nbytes, 0, 4096) + b"\\x01" * 100)   # announces 4096 bytes, sends 100, dies
        s.close()
        with pytest.raises(TimeoutError):
            rx.wait_and_apply(timeout=0.5)
        src = [("a", torch.arange(1000, dtype=torch.float32)), ("b", torch.ones(24))]
        st = StagingBuffer(rx.nbytes)
        tx = TCPWeightSender(num_streams=1)
        tx.stage(src, st)
        tx.send(st, "127.0.0.1", port, rx.nbytes)
        assert rx.wait_and_apply(timeout=10)
        assert torch.equal(dst[0][1], src[0][1]) and torch.equal(dst[1][1], src[1][1])
        st.close()
    finally:
        rx.close()


def test_dist_profiler_rank_filter_discrete_ranges_and_torch_window(tmp_path, monkeypatch):
    """Per-role rank filter, discrete nsys capture ranges per annotated phase, torch traces limited to a micro-step window."""
    from types import SimpleNamespace as NS

    from polyrl_b200.utils.profiler import DistProfiler

    g = NS(tool="nsys", save_path=str(tmp_path), global_tool_config={"nsys": {"discrete": True}})
    role = NS(enable=True, all_ranks=False, ranks=[1], tool_config={})
    assert not DistProfiler(0, g, role).this_rank and DistProfiler(1, g, role).this_rank
    assert DistProfiler(3, g, NS(enable=True, all_ranks=True, ranks=[], tool_config={})).this_rank

    class W:
        def __init__(self, prof):
            self.profiler = prof

        @DistProfiler.annotate("gen")
        def gen(self, x):
            return x + 1

    p = DistProfiler(1, g, role, role="actor")
    w = W(p)
    assert w.gen(1) == 2 and p.ranges_opened == 0          # outside a profiled step nothing is captured
    p.start()
    w.gen(1), w.gen(2)
    p.stop()
    assert p.ranges_opened == 2                              # one capture range per annotated call
    muted = W(DistProfiler(0, g, role))
    muted.profiler.start()
    muted.gen(1)
    assert muted.profiler.ranges_opened == 0                 # filtered rank
    # torch tool with a step window: only micro-steps [1, 3) are recorded -> exactly one trace file
    g2 = NS(tool="torch", save_path=str(tmp_path / "tr"), global_tool_config={})
    r2 = NS(enable=True, all_ranks=True, ranks=[], tool_config={"torch": {"step_start": 1, "step_end": 3}})
    t = DistProfiler(0, g2, r2, role="actor")
    t.start()
    for _ in range(5):
        torch.ones(4).sum()
        t.step()
    t.stop()
    files = list((tmp_path / "tr").glob("trace_actor_rank0_*.json"))
    assert len(files) == 1 and t.ranges_opened == 1


def test_multi_node_jobs_shard_inside_the_node():
    from polyrl_b200.workers.hybrid_worker import auto_fsdp_size

    assert auto_fsdp_size(-1, 16, "p2p", rank=1, local_world=8) == 8          # 2 nodes x 8 GPUs: HSDP 8 x 2
    assert auto_fsdp_size(-1, 8, "p2p", rank=1, local_world=8) == -1          # one node: plain ZeRO over all ranks
    assert auto_fsdp_size(-1, 16, "nccl", rank=1, local_world=8) == -1        # library collectives span nodes themselves
    assert auto_fsdp_size(4, 16, "p2p", rank=1, local_world=8) == 4           # explicit setting wins


def test_offline_generation_and_evaluation_entry_points(tmp_path):
    """main_generation (engine-backed batch sampling to parquet / jsonl) + main_eval (per-data-source scores of a generations file);
    reference: verl/trainer/main_generation.py, main_eval.py."""
    import json

    from polyrl_b200.trainer import main_eval, main_generation

    rows = [{"prompt": [5, 6, 7, 8 + i], "data_source": "openai/gsm8k", "reward_model": {"ground_truth": str(i)}} for i in range(5)]
    src = tmp_path / "prompts.jsonl"
    src.write_text("\n".join(json.dumps(r) for r in rows))
    out = tmp_path / "gen" / "out.parquet"
    assert main_generation.main(["--model", "tiny-llama", "--data", str(src), "--output", str(out), "--n-samples", "3",
                                 "--response-length", "6", "--batch-size", "2"]) == 0
    import pyarrow.parquet as pq

    got = pq.read_table(str(out)).to_pylist()
    assert len(got) == 5 and all(len(g["response_ids"]) == 3 and all(len(x) <= 6 for x in g["response_ids"]) for g in got)
    assert got[2]["prompt"] == rows[2]["prompt"]
    with pytest.raises(ValueError):
        main_generation.generate("tiny-llama", rows, n_samples=2, temperature=0.0)
    # evaluation on text responses: gsm8k scorer looks for "#### <answer>"
    ev = [{"data_source": "openai/gsm8k", "reward_model": {"ground_truth": "7"}, "responses": ["so #### 7", "maybe #### 8"]},
          {"data_source": "openai/gsm8k", "reward_model": {"ground_truth": "3"}, "responses": ["#### 4", "#### 5"]},
          {"data_source": "other", "reward_model": {"ground_truth": "x"}, "responses": ["x"]}]
    custom = tmp_path / "rw.py"
    custom.write_text("def compute_score(data_source, solution_str, ground_truth, extra_info=None):\n    return float(solution_str.strip() == ground_truth)\n")
    m = main_eval.evaluate(ev[:2], workers=1)
    assert m["test_score/openai/gsm8k"] == 0.25 and m["test_score/openai/gsm8k/best@2"] == 0.5 and m["num_prompts/openai/gsm8k"] == 2.0
    m2 = main_eval.evaluate(ev[2:], custom_reward=f"{custom}:compute_score", workers=1)
    assert m2["test_score/other"] == 1.0


def test_command_line_tools(tmp_path, capsys):
    """init_random_model (HF checkpoint that HF itself loads), print_cfg, diagnose, rollout_viewer."""
    import json

    from polyrl_b200.tools import diagnose, init_random_model, print_cfg, rollout_viewer

    out = tmp_path / "rand"
    assert init_random_model.main(["--model", "tiny-llama", "--out", str(out), "--override", "num_layers=1", "--dtype", "float32"]) == 0
    import transformers

    m = transformers.AutoModelForCausalLM.from_pretrained(str(out))
    assert m.config.num_hidden_layers == 1 and sum(p.numel() for p in m.parameters()) > 0
    capsys.readouterr()
    assert print_cfg.main(["--json", "actor_rollout_ref.rollout.n=4", "actor_rollout_ref.actor.use_dynamic_bsz=True"]) == 0
    cfg = json.loads(capsys.readouterr().out)
    assert cfg["actor_rollout_ref"]["rollout"]["n"] == 4
    info = diagnose.collect()
    assert info["python"] and "extension" in info and "optional_modules" in info and info["optional_modules"]["transformers"]
    d = tmp_path / "dump"
    d.mkdir()
    (d / "train_1.jsonl").write_text("\n".join(json.dumps({"step": 1, "score": s, "input": f"q{s}", "output": f"a #### {s}"}) for s in (0.0, 1.0, 0.5)))
    (d / "train_2.jsonl").write_text(json.dumps({"step": 2, "score": 0.25, "input": "q", "output": "zzz"}))
    rows = rollout_viewer.load(str(d))
    assert len(rows) == 4 and rollout_viewer.stats(rows)[1]["mean"] == 0.5
    assert [r["score"] for r in rollout_viewer.select(rows, step=1, min_score=0.4)] == [1.0, 0.5]
    assert len(rollout_viewer.select(rows, grep="zzz")) == 1
    assert rollout_viewer.main([str(d), "--stats"]) == 0 and rollout_viewer.main([str(d), "--top", "2"]) == 0


def test_fs_staging_and_checkpoint_mirror(tmp_path, monkeypatch):
    """utils/fs.py: local paths pass through, hdfs:// paths go through the `hdfs dfs` CLI into an md5-named cache (fetched once),
    checkpoints are mirrored to trainer.default_hdfs_dir (plain directory or hdfs://)."""
    import stat

    from polyrl_b200.utils import fs

    # a fake `hdfs` executable that maps hdfs://ns/<p> onto <tmp>/remote/<p>
    remote = tmp_path / "remote"
    remote.mkdir()
    bindir = tmp_path / "bin"
    bindir.mkdir()
    script = bindir / "hdfs"
    script.write_text(f"""#!/bin/bash
# usage: hdfs dfs -get|-put|-cp|-test|-mkdir ...
shift
map() {{ echo "$1" | sed "s#^hdfs://ns#{remote}#"; }}
case "$1" in
  -get) cp -r "$(map "$3")" "$4" ;;
  -put) mkdir -p "$(dirname "$(map "$4")")"; cp -r "$3" "$(map "$4")" ;;
  -cp) cp -r "$(map "$3")" "$(map "$4")" ;;
  -test) test -e "$(map "$3")" ;;
  -mkdir) mkdir -p "$(map "$3")" ;;
  *) exit 2 ;;
esac
""")
    script.chmod(script.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{bindir}:{__import__('os').environ['PATH']}")
    monkeypatch.setenv("POLYRL_CACHE_DIR", str(tmp_path / "cache"))
    (remote / "models" / "m1").mkdir(parents=True)
    (remote / "models" / "m1" / "config.json").write_text("{}")
    assert fs.is_non_local("hdfs://ns/models/m1") and not fs.is_non_local("/data/x") and not fs.is_non_local("file:///data/x")
    assert fs.exists("hdfs://ns/models/m1") and not fs.exists("hdfs://ns/models/none")
    loc = fs.copy_to_local("hdfs://ns/models/m1")
    assert (tmp_path / "cache") in __import__("pathlib").Path(loc).parents and (__import__("pathlib").Path(loc) / "config.json").exists()
    (remote / "models" / "m1" / "config.json").write_text('{"changed": 1}')
    assert fs.copy_to_local("hdfs://ns/models/m1") == loc and (__import__("pathlib").Path(loc) / "config.json").read_text() == "{}"   # cached
    assert fs.stage_path(["hdfs://ns/models/m1", "/plain/path"]) == [loc, "/plain/path"]
    assert fs.copy_to_local("file:///plain/path") == "/plain/path" and fs.copy_to_local(None) is None
    # checkpoint mirror: to a directory and to hdfs://
    ck = tmp_path / "ckpt" / "global_step_3"
    (ck / "actor").mkdir(parents=True)
    (ck / "actor" / "w.pt").write_text("x")
    (ck.parent / "latest_checkpointed_iteration.txt").write_text("3")
    dst = fs.mirror_checkpoint(ck, tmp_path / "mirror")
    assert (tmp_path / "mirror" / "global_step_3" / "actor" / "w.pt").exists() and (tmp_path / "mirror" / "latest_checkpointed_iteration.txt").read_text() == "3"
    assert dst.endswith("global_step_3")
    fs.mirror_checkpoint(ck, "hdfs://ns/ckpts")
    assert (remote / "ckpts" / "global_step_3" / "actor" / "w.pt").exists()
    assert fs.mirror_checkpoint(ck, None) is None
    monkeypatch.setenv("PATH", "/nonexistent")
    with pytest.raises(RuntimeError, match="hdfs"):
        fs.exists("hdfs://ns/x")


def test_reward_model_worker_scores_and_switches_tokenizers():
    """RewardModelWorker: a value-head model scores (prompt, response) on the last response token; with
    reward_model.model.input_tokenizer the pair is decoded and re-encoded for the reward model first (same score position in the
    original response)."""
    from polyrl_b200.config.schema import load_config
    from polyrl_b200.protocol import DataProto
    from polyrl_b200.workers.hybrid_worker import RewardModelWorker

    def batch():
        ids = torch.randint(40, 120, (3, 20))
        mask = torch.ones(3, 20, dtype=torch.long)
        mask[0, :3] = 0                     # left-padded prompt
        mask[1, 17:] = 0                    # short response
        return DataProto.from_dict({"input_ids": ids, "attention_mask": mask, "position_ids": (mask.cumsum(-1) - 1).clamp_min(0),
                                    "responses": ids[:, -8:], "response_mask": mask[:, -8:]})

    torch.manual_seed(0)
    cfg = load_config(overrides=["trainer.device=cpu", "reward_model.enable=True", "reward_model.model.path=tiny-llama",
                                 "reward_model.micro_batch_size_per_gpu=2"])
    w = RewardModelWorker(cfg, device="cpu")
    w.init_model()
    b = batch()
    s = w.compute_rm_score(b).batch["rm_scores"]
    last = b.batch["response_mask"].sum(-1) - 1
    assert s.shape == (3, 8) and all(float(s[i].abs().sum()) == float(s[i, last[i]].abs()) for i in range(3)) and s.abs().sum() > 0
    cfg2 = load_config(overrides=["trainer.device=cpu", "reward_model.enable=True", "reward_model.model.path=tiny-llama",
                                  "reward_model.micro_batch_size_per_gpu=2", "reward_model.model.input_tokenizer=byte",
                                  "reward_model.model.tokenizer_path=byte"])
    w2 = RewardModelWorker(cfg2, device="cpu")
    w2.init_model()
    s2 = w2.compute_rm_score(b).batch["rm_scores"]
    assert s2.shape == (3, 8) and all(float(s2[i].abs().sum()) == float(s2[i, last[i]].abs()) for i in range(3))
    sw = w2._switch_tokenizer(b)
    assert sw.batch["input_ids"].shape == b.batch["input_ids"].shape and sw.batch["attention_mask"][:, 12:].sum() > 0


def _slow_square(x, delay=0.0):
    import time as _t

    _t.sleep(delay)
    return x * x


def _raises(x):
    raise KeyError(x)


def test_py_functional_groupwise_import_utils_and_dtypes(tmp_path, monkeypatch):
    """The small utility surface users of the reference stack rely on (its tests: test_timeout_decorator_cpu, test_temp_env_on_cpu,
    test_groupwise, test_import_utils_on_cpu): wall-clock limits in a child process and through SIGALRM, scoped environment
    overrides, conflict-checked dict union, nested namespaces, run-time extensible enums, group statistics equal to a per-group
    loop, external-library / extern-type loading, precision names."""
    import os
    import warnings

    from polyrl_b200.utils import groupwise, import_utils, py_functional as P
    from polyrl_b200.utils.torch_dtypes import PrecisionType

    # ---- timeout_limit: child-process mode (kills a call that never returns), exceptions travel back, signal mode
    assert P.timeout_limit(5.0)(_slow_square)(7) == 49
    with pytest.raises(TimeoutError):
        P.timeout_limit(0.3)(_slow_square)(3, delay=5.0)
    with pytest.raises(KeyError):
        P.timeout_limit(5.0)(_raises)("k")
    with pytest.raises(TimeoutError):
        P.timeout_limit(0.2, use_signals=True)(_slow_square)(3, delay=5.0)
    assert P.timeout_limit(2.0, use_signals=True)(_slow_square)(4) == 16
    # ---- temp_env_var restores (also on error, also for previously unset keys)
    monkeypatch.setenv("PRL_T", "old")
    monkeypatch.delenv("PRL_U", raising=False)
    with P.temp_env_var("PRL_T", "new"), P.temp_env_var("PRL_U", "1"):
        assert os.environ["PRL_T"] == "new" and os.environ["PRL_U"] == "1"
        with P.temp_env_var("PRL_T", None):
            assert "PRL_T" not in os.environ
        assert os.environ["PRL_T"] == "new"
    assert os.environ["PRL_T"] == "old" and "PRL_U" not in os.environ
    with pytest.raises(RuntimeError):
        with P.temp_env_var("PRL_U", "x"):
            raise RuntimeError
    assert "PRL_U" not in os.environ
    # ---- dict helpers, namespaces, dynamic enums
    assert P.union_two_dict({"a": 1, "t": torch.ones(2)}, {"b": 2, "a": 1, "t": torch.ones(2)})["b"] == 2
    with pytest.raises(ValueError):
        P.union_two_dict({"a": 1}, {"a": 2})
    acc = P.append_to_dict(P.append_to_dict({}, {"loss": 1.0}, "actor/"), {"loss": 2.0}, "actor/")
    assert acc == {"actor/loss": [1.0, 2.0]}
    ns = P.NestedNamespace({"actor": {"optim": {"lr": 1e-6}, "stages": [{"n": 1}, {"n": 2}]}})
    assert ns.actor.optim.lr == 1e-6 and ns.actor.stages[1].n == 2 and ns.to_dict() == {"actor": {"optim": {"lr": 1e-6}, "stages": [{"n": 1}, {"n": 2}]}}

    class Estimator(P.DynamicEnum):
        pass

    class Other(P.DynamicEnum):
        pass

    g, r = Estimator.register("grpo"), Estimator.register("rloo")
    assert Estimator.GRPO is g and r.value == 1 and Estimator["RLOO"] is r and "GRPO" in Estimator and len(Estimator) == 2 and list(Estimator) == [g, r]
    assert len(Other) == 0 and Estimator.from_name("nope") is None
    with pytest.raises(ValueError):
        Estimator.register("GRPO")
    Estimator.remove("grpo")
    assert not hasattr(Estimator, "GRPO") and len(Estimator) == 1
    # ---- group statistics
    gen = torch.Generator().manual_seed(0)
    scores = torch.randn(23, generator=gen)
    uids = [f"u{int(i)}" for i in torch.randint(0, 6, (22,), generator=gen)] + ["lonely"]
    gi = groupwise.as_torch_index(uids)
    assert gi[-1].item() == gi.max().item() and gi[0].item() == 0
    mean, std, cnt = groupwise.group_mean_std(scores, gi)
    for k in range(int(gi.max()) + 1):
        sel = scores[gi == k]
        if sel.numel() == 1:
            assert mean[k] == 0 and std[k] == 1 and cnt[k] == 1
        else:
            assert abs(mean[k] - sel.mean()) < 1e-5 and abs(std[k] - sel.std(unbiased=True).clamp_min(1e-3)) < 1e-4 and cnt[k] == sel.numel()
    assert torch.equal(groupwise.as_torch_index(torch.tensor([5, 5, 9, 5, 2])), torch.tensor([0, 0, 1, 0, 2]))
    assert all(t.numel() == 0 for t in groupwise.group_mean_std(torch.empty(0), torch.empty(0, dtype=torch.long)))
    # ---- import utils
    assert import_utils.is_available("torch") and not import_utils.is_available("definitely_not_installed_xyz") and not import_utils.is_sglang_available()
    (tmp_path / "prl_ext_lib.py").write_text("REGISTERED = []\nREGISTERED.append('hello')\n\nclass MyReward:\n    tag = 'mine'\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    mods = import_utils.import_external_libs("prl_ext_lib, json")
    assert mods[0].REGISTERED == ["hello"] and import_utils.import_external_libs(None) == []
    cls = import_utils.load_extern_type(str(tmp_path / "prl_ext_lib.py"), "MyReward")
    assert cls.tag == "mine" and import_utils.load_extern_type(str(tmp_path / "prl_ext_lib.py"), "MyReward") is cls          # cached module
    assert import_utils.load_extern_type("pkg://json", "dumps") is json.dumps
    with pytest.raises(AttributeError):
        import_utils.load_extern_type(str(tmp_path / "prl_ext_lib.py"), "Missing")
    with pytest.raises(FileNotFoundError):
        import_utils.load_extern_type(str(tmp_path / "nope.py"), "X")

    @import_utils.deprecated("new_fn")
    def old_fn(x):
        return x + 1

    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert old_fn(1) == 2
    assert rec and "new_fn" in str(rec[0].message) and issubclass(rec[0].category, FutureWarning)
    # ---- precision names
    assert PrecisionType.to_dtype("bf16") is torch.bfloat16 and PrecisionType.to_dtype("16-mixed") is torch.float16 and PrecisionType.to_dtype(torch.float32) is torch.float32
    assert PrecisionType.is_bf16("torch.bfloat16") and PrecisionType.is_fp32("float") and PrecisionType.to_str("half") == "fp16" and not PrecisionType.supported("int8")


def test_numa_affinity_uses_the_gpus_cpu_mask(monkeypatch):
    """``set_numa_affinity``: opt-in, maps the local rank through CUDA_VISIBLE_DEVICES to the NVML index, applies the intersection of
    the GPU's CPU mask with the CPUs this process may use, and survives NVML errors (driven against a stand-in pynvml)."""
    import os
    import sys
    import types

    from polyrl_b200.utils import affinity

    if not hasattr(os, "sched_setaffinity"):
        pytest.skip("no sched_setaffinity")
    allowed = sorted(os.sched_getaffinity(0))
    asked: list = []
    nv = types.ModuleType("pynvml")
    nv.nvmlInit = lambda: None
    nv.nvmlShutdown = lambda: None
    nv.nvmlDeviceGetHandleByIndex = lambda i: asked.append(i) or i
    keep = allowed[: max(1, len(allowed) // 2)]
    mask_bits = sum(1 << c for c in keep) | (1 << 1000)                        # plus a CPU this container does not have
    nv.nvmlDeviceGetCpuAffinity = lambda h, words: [(mask_bits >> (64 * w)) & ((1 << 64) - 1) for w in range(max(words, 16))]
    monkeypatch.setitem(sys.modules, "pynvml", nv)
    monkeypatch.delenv("POLYRL_NUMA_AFFINITY", raising=False)
    try:
        assert affinity.set_numa_affinity(0) is None and not asked                # off by default
        monkeypatch.setenv("POLYRL_NUMA_AFFINITY", "1")
        monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "4,6")
        assert affinity.set_numa_affinity(1) == keep and asked == [6] and sorted(os.sched_getaffinity(0)) == keep
        nv.nvmlDeviceGetHandleByIndex = lambda i: (_ for _ in ()).throw(RuntimeError("nvml down"))
        assert affinity.set_numa_affinity(0) is None                               # reported, not raised
    finally:
        os.sched_setaffinity(0, allowed)
