"""Multi-turn tool / interaction agent loop (reference test model: verl tests/experimental/agent_loop/ with a fake
model server that scripts tool calls)."""
import itertools
import json

import pytest

import torch

from polyrl_b200.agent import (CalculatorTool, Gsm8kInteraction, Gsm8kTool, HermesToolParser, MultiTurnRollout, ToolSchema,
                               load_tools)
from polyrl_b200.agent.agent_loop import outputs_to_dataproto
from polyrl_b200.agent.tools import safe_arith
from polyrl_b200.data.rl_dataset import ByteTokenizer
from polyrl_b200.rollout.engine import Request, SamplingParams


class ScriptedEngine:
    """Engine stand-in: the reply is a function of the decoded context."""

    def __init__(self, tok, policy):
        self.tok, self.policy = tok, policy
        self.waiting, self._rid = [], itertools.count()

    def add_request(self, prompt_ids, params=None, rid=None, group=None):
        params = params if isinstance(params, SamplingParams) else SamplingParams.from_dict(params)
        r = Request(rid or f"s{next(self._rid)}", list(prompt_ids), params)
        self.waiting.append(r)
        return r

    def has_work(self):
        return bool(self.waiting)

    def step(self):
        r = self.waiting.pop(0)
        text = self.policy(self.tok.decode(r.prompt_ids))
        ids = self.tok.encode(text)[: r.params.max_new_tokens]
        r.output_ids, r.logprobs = ids, [-0.5] * len(ids)
        r.finish_reason = "length" if len(ids) >= r.params.max_new_tokens and len(self.tok.encode(text)) > len(ids) else "stop"
        return [r]


def test_safe_arith_and_schema():
    assert safe_arith("(3+4)*5") == 35 and abs(safe_arith("2**10/4") - 256) < 1e-9
    for bad in ("__import__('os')", "a+1", "9**9**9"):
        try:
            safe_arith(bad)
            assert False, bad
        except Exception:
            pass
    s = ToolSchema.from_openai(CalculatorTool.default_schema().to_openai())
    assert s.validate_arguments('{"expression": "1+1"}')[1] is None
    assert "missing" in s.validate_arguments("{}")[1]
    assert "JSON" in s.validate_arguments("{nope")[1]


def test_hermes_parser():
    tok = ByteTokenizer()
    txt = 'let me compute <tool_call>{"name": "calculator", "arguments": {"expression": "6*7"}}</tool_call> ok <tool_call>broken</tool_call>'
    content, calls = HermesToolParser(tok).extract_tool_calls(tok.encode(txt))
    assert len(calls) == 1 and calls[0].name == "calculator" and json.loads(calls[0].arguments) == {"expression": "6*7"}
    assert "<tool_call>" not in content


def test_gpt_oss_harmony_parser():
    """gpt-oss (harmony) tool calls: calls are the segments addressed to=functions.NAME and closed by <|call|>; tool-call markup inside
    the analysis channel (chain of thought) is not a call, also when the prompt already holds the leading <|start|>assistant;
    padding is dropped; a plain final answer has no calls (reference: tests/experimental/agent_loop/test_gpt_oss_tool_parser.py)."""
    from polyrl_b200.agent import get_tool_parser

    class Tok(ByteTokenizer):
        pad_token = "<|pad|>"

    tok = Tok()
    p = get_tool_parser("gpt-oss", tok)
    txt = ('<|channel|>analysis<|message|>maybe to=functions.fake <|constrain|>json<|message|>{"x":1}<|call|> no<|end|>'
           '<|start|>assistant<|channel|>commentary to=functions.get_weather <|constrain|>json<|message|>{"city": "Paris"}<|call|><|pad|><|pad|>')
    content, calls = p.extract_tool_calls(tok.encode(txt))
    assert [(c.name, json.loads(c.arguments)) for c in calls] == [("get_weather", {"city": "Paris"})] and content == ""
    two = ('<|start|>assistant<|channel|>analysis<|message|>think<|end|><|start|>assistant to=functions.calc<|channel|>commentary json'
           '<|message|>{"a": 2}<|call|><|start|>assistant<|channel|>commentary to=functions.b <|constrain|>json<|message|>{}<|call|>')
    _, calls = p.extract_tool_calls(tok.encode(two))
    assert [c.name for c in calls] == ["calc", "b"] and calls[0].arguments == '{"a": 2}'
    final = "<|start|>assistant<|channel|>final<|message|>The answer is 4<|return|>"
    assert p.extract_tool_calls(tok.encode(final)) == (final, [])


def test_tool_loop_masks_and_rewards():
    tok = ByteTokenizer()

    def policy(ctx: str) -> str:
        if "<tool>" not in ctx:                    # first turn: call the calculator
            return 'I will use a tool. <tool_call>{"name": "calculator", "arguments": {"expression": "6*7"}}</tool_call>'
        if "reward=" not in ctx:                   # second turn: submit the answer
            ans = ctx.split("<tool>")[1].split("\n")[0]
            return f'<tool_call>{{"name": "calc_gsm8k_reward", "arguments": {{"answer": "{ans}"}}}}</tool_call>'
        return "The answer is 42. #### 42"

    tools = load_tools({"tools": [{"class_name": "calculator"}, {"class_name": "calc_gsm8k_reward"}]})
    roll = MultiTurnRollout(ScriptedEngine(tok, policy), tok, tools=tools, response_length=600, max_assistant_turns=5)
    sample = {"messages": [{"role": "user", "content": "what is 6*7?"}],
              "tools_kwargs": {"calc_gsm8k_reward": {"create_kwargs": {"ground_truth": "42"}}}}
    outs = roll.generate([sample, dict(sample)], {"temperature": 1.0})
    assert len(outs) == 2
    o = outs[0]
    assert o.metrics["tool_calls"] == 2 and o.finish_reason == "stop"
    assert len(o.response_ids) == len(o.response_mask) == len(o.response_logprobs)
    text_all = tok.decode(o.response_ids)
    assert "<tool>42" in text_all and "reward=1.0" in text_all and text_all.endswith("#### 42")
    # assistant tokens carry loss + logprobs, tool tokens do not
    loss_text = tok.decode([t for t, m in zip(o.response_ids, o.response_mask) if m])
    ctx_text = tok.decode([t for t, m in zip(o.response_ids, o.response_mask) if not m])
    assert "<tool>" in ctx_text and "<tool>" not in loss_text and "#### 42" in loss_text
    assert all((lp == 0.0) == (m == 0) for lp, m in zip(o.response_logprobs, o.response_mask))
    assert o.tool_rewards[-1] == 1.0 or 1.0 in o.tool_rewards          # calc_reward at release
    dp = outputs_to_dataproto(outs, prompt_length=64, response_length=600, pad_id=tok.pad_token_id)
    b = dp.batch
    assert b["input_ids"].shape == (2, 664) and b["response_mask"].sum() == 2 * sum(o.response_mask)
    assert (b["attention_mask"][:, 64:].sum(-1) == len(o.response_ids)).all()
    assert (b["response_mask"] <= b["attention_mask"][:, 64:]).all()
    assert int(dp.non_tensor_batch["__num_turns__"][0]) == o.num_turns


def test_budget_exhaustion_and_unknown_tool():
    tok = ByteTokenizer()

    def policy(ctx):
        return '<tool_call>{"name": "nope", "arguments": {}}</tool_call>' + "x" * 40

    roll = MultiTurnRollout(ScriptedEngine(tok, policy), tok, tools=[CalculatorTool()], response_length=200, max_assistant_turns=8)
    o = roll.generate([{"messages": [{"role": "user", "content": "hi"}]}])[0]
    assert len(o.response_ids) <= 200 and "unknown tool" in tok.decode(o.response_ids)
    assert o.finish_reason == "length"


def test_interaction_loop():
    tok = ByteTokenizer()
    n = {"k": 0}

    def policy(ctx):
        n["k"] += 1
        return "#### 41" if "incorrect" not in ctx else "#### 42"

    roll = MultiTurnRollout(ScriptedEngine(tok, policy), tok, interactions={"gsm8k": Gsm8kInteraction({"name": "gsm8k"})},
                            response_length=400, max_assistant_turns=4)
    o = roll.generate([{"messages": [{"role": "user", "content": "6*7?"}],
                        "interaction_kwargs": {"name": "gsm8k", "ground_truth": "42"}}])[0]
    txt = tok.decode(o.response_ids)
    assert "incorrect" in txt and txt.rstrip().endswith("#### 42")
    assert o.turn_scores == [0.0, 1.0]


def _scripted_user(messages, **kw):
    """Simulated user for the worker-level CollabLLM test: one follow-up question, then done."""
    mine = sum(m["role"] == "assistant" for m in messages)
    return "and then?" if mine < 2 else "ok [[TERMINATE CHAT]]"


def test_real_engine_smoke():
    """Random tiny model: whatever it emits, the loop must terminate with consistent bookkeeping."""
    from polyrl_b200.models import TransformerWeights, get_model_config
    from polyrl_b200.rollout.engine import RolloutEngine

    cfg = get_model_config("tiny-llama")
    w = TransformerWeights(cfg, "cpu", torch.float32, seed=3)
    eng = RolloutEngine(cfg, "cpu", max_num_seqs=8, max_model_len=160, use_cuda_graph=False, seed=1)
    eng.set_weights(w.embed, w.layers, w.head)
    tok = ByteTokenizer()
    roll = MultiTurnRollout(eng, tok, tools=[CalculatorTool()], response_length=48, max_assistant_turns=3)
    outs = roll.generate([{"messages": [{"role": "user", "content": f"q{i}"}]} for i in range(3)], {"temperature": 1.0})
    for o in outs:
        assert 0 < len(o.response_ids) <= 48 and len(o.response_ids) == len(o.response_mask)


import pytest


@pytest.mark.parametrize("agent", [None, "react_agent", "collabllm_agent"])
def test_worker_multi_turn_generate_and_update(tmp_path, agent):
    """Colocated worker with multi_turn.enable: trajectories with tool config flow through generate -> log-prob -> update; the agent
    loop class is selected by ``rollout.agent.default_agent_loop`` (graph-structured ReAct agent, CollabLLM forward sampling)."""
    import numpy as np
    import yaml

    from polyrl_b200.config.schema import load_config
    from polyrl_b200.protocol import DataProto
    from polyrl_b200.workers.hybrid_worker import ActorRolloutRefWorker

    tools_yaml = tmp_path / "tools.yaml"
    tools_yaml.write_text(yaml.safe_dump({"tools": [{"class_name": "calculator"}]}))
    cfg = load_config(None, ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-llama", "data.max_prompt_length=16",
                             "data.max_response_length=24", "actor_rollout_ref.rollout.n=2", "actor_rollout_ref.rollout.max_num_seqs=8",
                             "actor_rollout_ref.rollout.cuda_graph=False", "actor_rollout_ref.rollout.multi_turn.enable=True",
                             f"actor_rollout_ref.rollout.multi_turn.tool_config_path={tools_yaml}",
                             "actor_rollout_ref.actor.ppo_mini_batch_size=2", "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=2",
                             "actor_rollout_ref.rollout.log_prob_micro_batch_size_per_gpu=4"]
                      + ([f"actor_rollout_ref.rollout.agent.default_agent_loop={agent}"] if agent else []))
    nt = {"uid": np.array(["a", "b"], dtype=object)}
    if agent == "collabllm_agent":
        inter_yaml = tmp_path / "inter.yaml"
        inter_yaml.write_text(yaml.safe_dump({"interaction": [{"name": "collabllm", "class_name": "collabllm",
                                                               "config": {"user_model": "tests.test_agent_loop_cpu._scripted_user"}}]}))
        cfg.actor_rollout_ref.rollout.multi_turn.interaction_config_path = str(inter_yaml)
        cfg.actor_rollout_ref.rollout.agent["agent_loop_kwargs"] = {"num_repeat_rollouts": 2, "metric_weights": {"token_amount": -0.001, "interactivity": 1.0}}
        ik = {"name": "collabllm", "single_turn_prompt": "find x", "ground_truth": "4"}
        nt["interaction_kwargs"] = np.array([ik, ik], dtype=object)
        nt["raw_prompt"] = np.array([[{"role": "user", "content": "hi"}], [{"role": "user", "content": "yo"}]], dtype=object)
    w = ActorRolloutRefWorker(cfg, role="actor_rollout", device="cpu")
    w.init_model()
    ids = torch.randint(3, 200, (2, 16))
    mask = torch.ones_like(ids)
    mask[0, :5] = 0
    prompts = DataProto.from_dict({"input_ids": ids, "attention_mask": mask, "position_ids": (mask.cumsum(-1) - 1).clamp_min(0)},
                                  non_tensors=nt)
    out = w.generate_sequences(prompts)
    assert out.batch["responses"].shape == (4, 24) and "__num_turns__" in out.non_tensor_batch
    if agent:
        assert type(w._multi_turn).__name__ == {"react_agent": "ReactAgentLoop", "collabllm_agent": "CollabLLMAgentLoop"}[agent]
    if agent == "collabllm_agent":
        assert "rm_scores" in out.batch and (out.non_tensor_batch["collabllm/num_futures"] == 2).all()
    assert (out.batch["response_mask"] <= out.batch["attention_mask"][:, -24:]).all()
    old = w.compute_log_prob(out)
    out = out.union(old)
    out.batch["advantages"] = torch.randn(4, 24) * out.batch["response_mask"]
    out.meta_info.update({"is_opt_step": True})
    m = w.update_actor_stream(out)
    assert "actor/pg_loss" in m.meta_info["metrics"]


def test_rollout_trace_file_backend(tmp_path):
    import json as _json

    from polyrl_b200.utils.rollout_trace import RolloutTraceConfig, rollout_trace_attr, rollout_trace_op

    RolloutTraceConfig.reset()

    class T:
        @rollout_trace_op
        def work(self, x, k=1):
            return x * k

    t = T()
    assert t.work(3, k=2) == 6                                   # tracing off: plain call
    cfg = RolloutTraceConfig.init("proj", "exp", "file", out_dir=str(tmp_path))
    with rollout_trace_attr(step=5, sample_index=2, rollout_n=1):
        assert t.work(4, k=3) == 12
    rows = [_json.loads(l) for l in cfg.path.read_text().splitlines()]
    assert len(rows) == 1 and rows[0]["step"] == 5 and rows[0]["output"] == 12 and rows[0]["inputs"] == {"x": 4, "k": 3}
    RolloutTraceConfig.reset()


def test_rollout_trace_weave_mlflow_async_and_token2text(tmp_path, monkeypatch):
    """The weave and mlflow trace back-ends against recording stand-ins for the client libraries (neither is installed): one call /
    span per traced op with named inputs inside the per-sample attribute scope, exceptions reported, async methods traced, and
    ``token2text`` adding the decoded prompt / response next to the ids."""
    import asyncio
    import contextlib
    import sys
    import types

    from polyrl_b200.utils.rollout_trace import RolloutTraceConfig, rollout_trace_attr, rollout_trace_op

    log: list = []

    class Out:
        def __init__(self):
            self.prompt_ids, self.response_ids, self.reward = [3 + ord("h"), 3 + ord("i")], [3 + ord("o"), 3 + ord("k")], 1.0

    class Agent:
        tokenizer = ByteTokenizer()

        @rollout_trace_op
        async def run(self, prompt, n=1):
            await asyncio.sleep(0)
            return Out()

        @rollout_trace_op
        def boom(self, x):
            raise ValueError("bad " + str(x))

    # ---- weave
    class Client:
        def create_call(self, op, inputs, attributes):
            log.append(("create", op, inputs, attributes))
            return len(log)

        def finish_call(self, call, output=None, exception=None):
            log.append(("finish", call, output, repr(exception) if exception else None))

    scope: list = [{}]
    weave = types.ModuleType("weave")
    weave.init = lambda project: Client()

    @contextlib.contextmanager
    def attributes(d):
        scope.append(d)
        try:
            yield
        finally:
            scope.pop()

    weave.attributes = attributes
    ctx_mod = types.ModuleType("weave.trace.context")
    ctx_mod.call_context = types.SimpleNamespace(call_attributes=types.SimpleNamespace(get=lambda: scope[-1]))
    for name, mod in (("weave", weave), ("weave.trace", types.ModuleType("weave.trace")), ("weave.trace.context", ctx_mod)):
        monkeypatch.setitem(sys.modules, name, mod)
    RolloutTraceConfig.reset()
    RolloutTraceConfig.init("proj", "exp", "weave", token2text=True)
    a = Agent()
    with rollout_trace_attr(step=2, sample_index=7, rollout_n=1, validate=True):
        out = asyncio.run(a.run("hello", n=4))
        with pytest.raises(ValueError):
            a.boom(5)
    assert isinstance(out, Out)
    assert log[0][:3] == ("create", "test_rollout_trace_weave_mlflow_async_and_token2text.<locals>.Agent.run", {"prompt": "hello", "n": 4})
    assert log[0][3] == {"step": 2, "sample_index": 7, "rollout_n": 1, "validate": True, "experiment_name": "exp"}
    assert log[1][2]["prompt_text"] == "hi" and log[1][2]["response_text"] == "ok" and log[1][2]["reward"] == 1.0 and log[1][3] is None
    assert log[2][2] == {"x": 5} and "bad 5" in log[3][3]
    # ---- mlflow
    log.clear()

    class Span:
        trace_id = "t1"

        def __init__(self, name):
            self.name = name

        def set_inputs(self, i):
            log.append(("inputs", self.name, i))

        def set_outputs(self, o):
            log.append(("outputs", self.name, o))

    @contextlib.contextmanager
    def start_span(name):
        log.append(("open", name))
        try:
            yield Span(name)
        finally:
            log.append(("close", name))

    mlflow = types.ModuleType("mlflow")
    mlflow.start_span = start_span
    mlflow.set_trace_tag = lambda tid, k, v: log.append(("tag", tid, k, v))
    mlflow.set_tracking_uri = lambda uri: log.append(("uri", uri))
    mlflow.set_experiment = lambda name: log.append(("experiment", name))
    mlflow.config = types.SimpleNamespace(enable_async_logging=lambda: log.append(("async",)))
    monkeypatch.setitem(sys.modules, "mlflow", mlflow)
    monkeypatch.delenv("MLFLOW_TRACKING_URI", raising=False)
    RolloutTraceConfig.reset()
    RolloutTraceConfig.init("proj", "exp", "mlflow")
    assert log[:3] == [("async",), ("uri", "sqlite:////tmp/mlruns.db"), ("experiment", "proj")]
    with rollout_trace_attr(step=3, sample_index=0, name="agent_loop"):
        asyncio.run(a.run("q"))
        with pytest.raises(ValueError):
            a.boom(1)
    names = [e[:2] for e in log[3:] if e[0] in ("open", "close")]
    run, boom = "Agent.run", "Agent.boom"
    assert [n[0] for n in names] == ["open", "open", "close", "open", "close", "close"] and names[0][1] == "agent_loop" and names[1][1].endswith(run) and names[3][1].endswith(boom)
    assert ("tag", "t1", "step", "3") in log and ("tag", "t1", "experiment_name", "exp") in log
    outs = [e for e in log if e[0] == "outputs"]
    assert len(outs) == 1 and isinstance(outs[0][2], Out)              # token2text off: the object itself; the failed call has no output
    RolloutTraceConfig.reset()


def test_sandbox_code_reward_and_interpreter_tool():
    from polyrl_b200.agent.tools import CodeInterpreterTool
    from polyrl_b200.data.reward_score import code_exec, default_compute_score
    from polyrl_b200.utils.sandbox import run_python

    assert run_python("print(sum(range(10)))").stdout.strip() == "45"
    assert run_python("while True: pass", timeout_s=1.0).status == "timeout"
    assert run_python("raise ValueError('x')").status == "runtime_error"
    assert run_python("x = bytearray(3 * 1024**3)", mem_mb=256).status in ("runtime_error", "limit")      # address-space limit
    sol = "Here is my solution:\n```python\nn = int(input())\nprint(n * 2)\n```\n"
    s, info = code_exec.compute_score(sol, {"inputs": ["3", "10"], "outputs": ["6", "20"]})
    assert s == 1.0 and info == {"passed": 2, "total": 2}
    s, _ = code_exec.compute_score(sol, {"inputs": ["3", "10"], "outputs": ["6", "21"]})
    assert s == 0.5 and code_exec.compute_score(sol, {"inputs": ["3", "10"], "outputs": ["6", "21"]}, continuous=False)[0] == 0.0
    fn = "```python\ndef add(a, b):\n    return a + b\n```"
    assert code_exec.compute_score(fn, {"fn_name": "add", "inputs": [[1, 2], [5, 5]], "outputs": [3, 10]})[0] == 1.0
    assert code_exec.compute_score(fn, {"assert_case": ["assert add(2, 2) == 4", "assert add(1, 1) == 3"]})[0] == 0.5
    assert code_exec.compute_score("no code", {"inputs": [], "outputs": []})[0] == 0.0
    assert default_compute_score("code_exec", sol, {"inputs": ["4"], "outputs": ["8"]}) == 1.0
    tool = CodeInterpreterTool()
    iid, _ = tool.create()
    resp, _, m = tool.execute(iid, {"code": "print(6*7)"})
    assert resp.text == "42" and m["sandbox_status"] == "ok"
    resp, _, m = tool.execute(iid, {"code": "1/0"})
    assert "ZeroDivisionError" in resp.text and m["sandbox_status"] == "runtime_error"


def test_search_tool_bm25_and_qa_reward():
    from polyrl_b200.agent.tools import SearchTool
    from polyrl_b200.data.reward_score import qa_em

    corpus = [{"title": "Eiffel Tower", "contents": "The Eiffel Tower is a wrought-iron lattice tower in Paris, France."},
              {"title": "Big Ben", "contents": "Big Ben is the nickname for the Great Bell of the clock in London."},
              {"title": "Colosseum", "contents": "The Colosseum is an ancient amphitheatre in the centre of Rome, Italy."}]
    tool = SearchTool({"corpus": corpus, "topk": 2})
    iid, _ = tool.create()
    resp, _, m = tool.execute(iid, {"query_list": ["where is the eiffel tower", "amphitheatre rome"]})
    first, second = resp.text.split("\n---\n")
    assert first.startswith("Doc 1 (Title: Eiffel Tower)") and "Colosseum" in second.splitlines()[0] and m["queries"] == 2
    assert tool.execute(iid, {"query_list": ["zzzz qqqq"]})[0].text == "no results"
    assert qa_em.compute_score("<answer>Paris</answer>", {"target": ["paris"]}) == 1.0


def test_remote_and_image_tools_against_local_fake_services():
    """MCP (JSON-RPC over HTTP), sandbox-fusion code execution and the image zoom tool, exercised against tiny in-process HTTP
    servers; the code-exec scorer can run its test cases through the remote sandbox."""
    import json
    import threading
    from http.server import BaseHTTPRequestHandler, HTTPServer

    import numpy as np
    from PIL import Image

    from polyrl_b200.agent.remote_tools import ImageZoomInTool, MCPSearchTool, MCPTool, SandboxFusionTool, TokenBucket
    from polyrl_b200.agent.tools import load_tools
    from polyrl_b200.data.reward_score import code_exec
    from polyrl_b200.utils.sandbox import run_python

    seen = []

    class H(BaseHTTPRequestHandler):
        def log_message(self, *a):
            pass

        def do_POST(self):
            body = json.loads(self.rfile.read(int(self.headers["Content-Length"])))
            if self.path == "/mcp":
                seen.append(body["method"])
                if body["method"] == "initialize":
                    res = {"protocolVersion": "2025-03-26", "capabilities": {"tools": {}}}
                elif body["method"] == "tools/list":
                    res = {"tools": [{"name": "web_search", "description": "search", "inputSchema": {"type": "object", "required": ["query"], "properties": {"query": {"type": "string"}}}},
                                     {"name": "echo", "inputSchema": {"type": "object", "properties": {}}}]}
                elif body["params"]["name"] == "web_search":
                    res = {"content": [{"type": "text", "text": json.dumps({"results": [{"title": "T", "url": "u", "content": "about " + body["params"]["arguments"]["query"]}]})}]}
                elif body["params"]["name"] == "echo":
                    res = {"content": [{"type": "text", "text": json.dumps(body["params"]["arguments"])}], "isError": False}
                else:
                    out = {"jsonrpc": "2.0", "id": body["id"], "error": {"code": -32601, "message": "unknown tool"}}
                    return self._send(out)
                return self._send({"jsonrpc": "2.0", "id": body["id"], "result": res})
            if self.path == "/run_code":                # sandbox-fusion: really run it, in the local sandbox
                r = run_python(body["code"], stdin=body.get("stdin", ""), timeout_s=body.get("run_timeout", 5))
                return self._send({"status": "Success" if r.status == "ok" else "Failed",
                                   "run_result": {"stdout": r.stdout, "stderr": r.stderr, "return_code": 0 if r.status == "ok" else 1}})
            self.send_response(404)
            self.end_headers()

        def _send(self, obj):
            d = json.dumps(obj).encode()
            self.send_response(200)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(d)))
            self.end_headers()
            self.wfile.write(d)

    srv = HTTPServer(("127.0.0.1", 0), H)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    base = f"http://127.0.0.1:{srv.server_port}"
    try:
        tools = MCPTool.discover(base + "/mcp")
        assert [t.name for t in tools] == ["web_search", "echo"] and tools[0].schema.parameters["required"] == ["query"]
        assert seen.count("initialize") == 1                                  # one handshake, shared client
        resp, _, m = tools[1].execute("i", {"a": 1})
        assert json.loads(resp.text) == {"a": 1} and m == {"mcp_error": 0}
        st = load_tools({"tools": [{"class_name": "mcp_search", "config": {"url": base + "/mcp", "tool_name": "web_search", "rate_limit": 100},
                                    "tool_schema": {"type": "function", "function": {"name": "search", "parameters": {"type": "object", "properties": {}}}}}]})[0]
        assert isinstance(st, MCPSearchTool) and st.name == "search"
        resp, _, _ = st.execute("i", {"query": "blackwell"})
        assert resp.text == "[1] T (u)\nabout blackwell"
        bad = MCPTool({"url": base + "/mcp", "tool_name": "nope"})
        assert "tool call failed" in bad.execute("i", {})[0].text
        # remote sandbox
        sb = SandboxFusionTool({"sandbox_fusion_url": base + "/run_code"})
        assert sb.execute("i", {"code": "print(6*7)"})[0].text.strip() == "42"
        r, _, m = sb.execute("i", {"code": "raise ValueError('boom')"})
        assert "boom" in r.text and m == {"sandbox_ok": 0}
        sol = "```python\nn=int(input())\nprint(n*2)\n```"
        score, info = code_exec.compute_score(sol, {"inputs": ["2", "5"], "outputs": ["4", "10"]}, sandbox_fusion_url=base + "/run_code")
        assert score == 1.0 and info == {"passed": 2, "total": 2}
    finally:
        srv.shutdown()
    # token bucket: 3 immediate, then paced
    import time

    tb = TokenBucket(50, burst=3)
    t0 = time.monotonic()
    for _ in range(5):
        tb.acquire()
    assert 0.03 <= time.monotonic() - t0 < 0.5
    # image zoom
    img = Image.fromarray(np.arange(100 * 80 * 3, dtype=np.uint32).reshape(100, 80, 3).astype(np.uint8))
    z = ImageZoomInTool({"min_size": 28})
    iid, _ = z.create(image=img)
    resp, pen, m = z.execute(iid, {"bbox_2d": [10, 20, 50, 90], "label": "corner"})
    assert resp.image.size == (40, 70) and "corner" in resp.text and m == {"zoom_ok": 1}
    resp, _, _ = z.execute(iid, {"bbox_2d": [0, 0, 10, 10]})
    assert min(resp.image.size) >= 28                                       # tiny crops are enlarged
    assert z.execute(iid, {"bbox_2d": [200, 200, 300, 300]})[1] < 0 and z.execute(iid, {"bbox_2d": [1, 2]})[1] < 0
    g = load_tools({"tools": [{"class_name": "calc_geo3k_reward"}]})[0]
    gid, _ = g.create(ground_truth="45")
    assert g.execute(gid, {"answer": "45"})[0].text.endswith("reward=1.0") and g.calc_reward(gid) == 1.0


def test_image_tool_responses_extend_the_vision_context(tmp_path):
    """Vision-language multi-turn: the zoom tool's crop enters the context as an image-token run, the trajectory's patches grow, the
    batch carries 3-axis positions over prompt + response and the trainer can score it (tower rows == placeholders)."""
    import numpy as np
    from PIL import Image

    from polyrl_b200.config.schema import load_config
    from polyrl_b200.models import vision as V
    from polyrl_b200.models.config import get_model_config
    from polyrl_b200.workers.hybrid_worker import ActorRolloutRefWorker

    cfg = get_model_config("tiny-qwen2-vl")
    tok = ByteTokenizer()
    img = Image.fromarray(np.random.default_rng(0).integers(0, 255, size=(32, 48, 3), dtype=np.uint8))
    pv, g = V.preprocess_image(cfg, img, 64, 1024)
    n0 = V.num_image_tokens(cfg, [g])[0]
    prompt = tok.encode("<user>") + [cfg.vision_start_token_id] + [cfg.image_token_id] * n0 + [cfg.vision_end_token_id] + tok.encode("what is in the corner?\n<assistant>")

    seen_mm = []

    class VLScripted(ScriptedEngine):
        def add_request(self, prompt_ids, params=None, rid=None, group=None, mm=None):
            seen_mm.append(None if mm is None else (mm["pixel_values"].shape[0], list(mm["grid_thw"])))
            return super().add_request(prompt_ids, params, rid, group)

    def policy(ctx: str) -> str:
        if "<tool>" not in ctx:
            return 'look closer <tool_call>{"name": "image_zoom_in_tool", "arguments": {"bbox_2d": [0, 0, 24, 16], "label": "corner"}}</tool_call>'
        return "a cat"

    tools = load_tools({"tools": [{"class_name": "image_zoom_in_tool", "config": {"min_size": 8}}]})
    roll = MultiTurnRollout(VLScripted(tok, policy), tok, tools=tools, response_length=200, max_assistant_turns=3, model_cfg=cfg,
                            image_min_pixels=64, image_max_pixels=1024)
    sample = {"prompt_ids": prompt, "multi_modal_inputs": {"pixel_values": pv, "image_grid_thw": torch.tensor([g])},
              "tools_kwargs": {"image_zoom_in_tool": {"create_kwargs": {"image": img}}}}
    (o,) = roll.generate([sample], {"temperature": 1.0})
    assert o.finish_reason == "stop" and o.metrics["tool_calls"] == 1
    grids = [tuple(x) for x in o.multi_modal_inputs["image_grid_thw"].tolist()]
    assert len(grids) == 2 and grids[0] == g and seen_mm[0] == (pv.shape[0], [g]) and seen_mm[1][1] == grids
    n1 = V.num_image_tokens(cfg, grids)[1]
    assert o.response_ids.count(cfg.image_token_id) == n1 and o.multi_modal_inputs["pixel_values"].shape[0] == pv.shape[0] + grids[1][0] * grids[1][1] * grids[1][2]
    img_pos = [k for k, t in enumerate(o.response_ids) if t == cfg.image_token_id]
    assert all(o.response_mask[k] == 0 for k in img_pos)                      # tool image tokens carry no loss
    assert tok.decode([t for t, m in zip(o.response_ids, o.response_mask) if m]).endswith("a cat")
    dp = outputs_to_dataproto([o], prompt_length=96, response_length=200, pad_id=0, model_cfg=cfg)
    pos3 = dp.batch["position_ids"]
    n = len(o.response_ids)
    want, _ = V.mrope_positions(cfg, prompt + o.response_ids, grids)
    assert pos3.shape == (1, 3, 296) and torch.equal(pos3[0, :, 96 - len(prompt): 96 + n], want)
    assert not torch.equal(pos3[0, 0], pos3[0, 1])                            # image rows really use separate axes
    # the trainer consumes it: log-probs over the response, tower fed with both images
    c = load_config(None, ["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-qwen2-vl", "data.max_prompt_length=96",
                           "data.max_response_length=200", "actor_rollout_ref.rollout.max_num_seqs=4", "actor_rollout_ref.rollout.cuda_graph=False",
                           "actor_rollout_ref.actor.ppo_mini_batch_size=1", "actor_rollout_ref.actor.ppo_micro_batch_size_per_gpu=1"])
    wk = ActorRolloutRefWorker(c, role="actor", device="cpu")
    wk.init_model()
    dp.meta_info.update(temperature=1.0, micro_batch_size=1)
    lp = wk.compute_log_prob(dp).batch["old_log_probs"]
    assert lp.shape == (1, 200) and torch.isfinite(lp).all() and (lp[0, :n][torch.tensor(o.response_mask, dtype=torch.bool)] < 0).all()


def test_react_graph_agent_equals_the_tool_loop_and_custom_graphs_run_batched():
    """agent/graph.py: the ReAct graph (model node <-> tools node, one thread per trajectory, every ``ctx.chat`` a request into the
    shared engine) yields token for token the trajectory of the built-in tool loop; a custom graph adds a node that injects a user
    message and scores the conversation; graph construction errors are reported at compile time; a crashing node ends only its
    own trajectory."""
    import pytest

    from polyrl_b200.agent import get_agent_loop_cls
    from polyrl_b200.agent.graph import END, GraphAgentLoop, StateGraph

    tok = ByteTokenizer()

    def policy(ctx: str) -> str:
        if "<tool>" not in ctx:
            return 'I will use a tool. <tool_call>{"name": "calculator", "arguments": {"expression": "6*7"}}</tool_call>'
        if "are you sure" in ctx and "yes, 42" not in ctx:
            return "yes, 42"
        if "reward=" not in ctx:
            ans = ctx.split("<tool>")[1].split("\n")[0]
            return f'<tool_call>{{"name": "calc_gsm8k_reward", "arguments": {{"answer": "{ans}"}}}}</tool_call>'
        return "The answer is 42. #### 42"

    spec = {"tools": [{"class_name": "calculator"}, {"class_name": "calc_gsm8k_reward"}]}
    sample = {"messages": [{"role": "user", "content": "what is 6*7?"}],
              "tools_kwargs": {"calc_gsm8k_reward": {"create_kwargs": {"ground_truth": "42"}}}}
    base = MultiTurnRollout(ScriptedEngine(tok, policy), tok, tools=load_tools(spec), response_length=600, max_assistant_turns=5)
    want = base.generate([dict(sample) for _ in range(3)], {"temperature": 1.0})
    react = get_agent_loop_cls("react_agent")(ScriptedEngine(tok, policy), tok, tools=load_tools(spec), response_length=600, max_assistant_turns=5)
    got = react.generate([dict(sample) for _ in range(3)], {"temperature": 1.0})
    for a, b in zip(got, want):
        assert a.response_ids == b.response_ids and a.response_mask == b.response_mask and a.response_logprobs == b.response_logprobs
        assert a.metrics["tool_calls"] == 2 and a.num_turns == b.num_turns and a.finish_reason == "stop"
    assert [m["role"] for m in react.last_messages[0]] == ["user", "assistant", "tool", "assistant", "tool", "assistant"]

    # a custom graph: model -> (tool calls -> tools -> model) | challenge once -> model | grade -> END
    def model(state, ctx):
        msg = ctx.chat(state["messages"])
        return {"messages": [msg]} if msg is not None else {"exhausted": True}

    def tools(state, ctx):
        return {"messages": ctx.run_tools(state["messages"][-1]["tool_calls"])}

    def challenge(state, ctx):
        return {"messages": [{"role": "user", "content": "are you sure?"}], "challenged": True}

    def grade(state, ctx):
        if ctx.sample.get("boom"):
            raise RuntimeError("grader crashed")
        return {"reward_score": 1.0 if "42" in (state["messages"][-1].get("content") or "") else 0.0}

    def route(state):
        last = state["messages"][-1]
        if state.get("exhausted"):
            return "grade"
        if last.get("tool_calls"):
            return "tools"
        return "challenge" if not state.get("challenged") else "grade"

    g = (StateGraph().add_node("model", model).add_node("tools", tools).add_node("challenge", challenge).add_node("grade", grade)
         .set_entry_point("model").add_conditional_edges("model", route).add_edge("tools", "model").add_edge("challenge", "model")
         .add_edge("grade", END))
    loop = GraphAgentLoop(ScriptedEngine(tok, policy), tok, graph_factory=g.compile, tools=load_tools(spec), response_length=900,
                          max_assistant_turns=8)
    outs = loop.generate([dict(sample), {**sample, "boom": True}], {"temperature": 1.0})
    o = outs[0]
    text = tok.decode(o.response_ids)
    assert o.reward_score == 1.0 and "are you sure?" in text and "yes, 42" in text
    ctx_text = tok.decode([t for t, m in zip(o.response_ids, o.response_mask) if not m])
    assert "are you sure?" in ctx_text and "yes, 42" not in ctx_text          # injected turn: context; model turns: trained
    assert outs[1].finish_reason.startswith("error: RuntimeError") and outs[1].reward_score is None and outs[1].response_ids
    dp = outputs_to_dataproto(outs, prompt_length=64, response_length=900, pad_id=tok.pad_token_id)
    assert float(dp.batch["rm_scores"][0].sum()) == 1.0 and float(dp.batch["rm_scores"][1].sum()) == 0.0
    # budget: one assistant turn only -> chat returns None on the second call and the graph still terminates
    tight = GraphAgentLoop(ScriptedEngine(tok, policy), tok, graph_factory=g.compile, tools=load_tools(spec), response_length=900,
                           max_assistant_turns=1)
    t = tight.generate([dict(sample)], {"temperature": 1.0})[0]
    assert t.finish_reason == "length" and t.metrics["tool_calls"] == 1
    # construction errors
    with pytest.raises(ValueError):
        StateGraph().add_node("a", model).compile()                           # no entry point
    with pytest.raises(ValueError):
        StateGraph().add_node("a", model).set_entry_point("a").compile()       # no way out of "a"
    with pytest.raises(ValueError):
        StateGraph().add_node("a", model).set_entry_point("a").add_edge("a", "nowhere").compile()
    # chat() must extend the conversation the engine has seen
    def bad(state, ctx):
        ctx.chat(state["messages"])
        return {"messages": [ctx.chat([{"role": "user", "content": "something else"}])]}

    bg = StateGraph().add_node("bad", bad).set_entry_point("bad").add_edge("bad", END)
    r = GraphAgentLoop(ScriptedEngine(tok, policy), tok, graph_factory=bg.compile, response_length=300).generate([dict(sample)])[0]
    assert r.finish_reason.startswith("error: ValueError")
