"""Tool abstraction for multi-turn rollouts.

A tool has an OpenAI-function style schema (rendered into the system prompt by the chat template) and a
per-trajectory lifecycle: ``create`` -> ``execute``* -> ``calc_reward`` -> ``release``.

Capability parity: 3rdparty/verl/verl/tools/base_tool.py:24-95 (BaseTool lifecycle), tools/schemas.py:21-80
(OpenAI function schema + parsed call), tools/gsm8k_tool.py:31-110 (answer-checking tool with step rewards),
tools/utils/tool_registry.py (config-driven instantiation).  The implementation is synchronous — the batched
``MultiTurnRollout`` scheduler runs tool calls on a thread pool while the engine keeps decoding the other
trajectories of the continuous batch.
"""
from __future__ import annotations

import ast
import importlib
import json
import operator
import uuid
from dataclasses import dataclass, field
from typing import Any


@dataclass
class ToolSchema:
    name: str
    description: str = ""
    parameters: dict = field(default_factory=lambda: {"type": "object", "properties": {}, "required": []})
    strict: bool = False

    def to_openai(self) -> dict:
        return {"type": "function", "function": {"name": self.name, "description": self.description,
                                                 "parameters": self.parameters, "strict": self.strict}}

    @classmethod
    def from_openai(cls, d: dict) -> "ToolSchema":
        f = d.get("function", d)
        return cls(name=f["name"], description=f.get("description", ""),
                   parameters=f.get("parameters") or {"type": "object", "properties": {}, "required": []},
                   strict=bool(f.get("strict", False)))

    def validate_arguments(self, args: Any) -> tuple[dict, str | None]:
        """Returns (arguments, error).  Checks JSON-object-ness and required keys."""
        if isinstance(args, str):
            try:
                args = json.loads(args)
            except json.JSONDecodeError as e:
                return {}, f"arguments are not valid JSON: {e}"
        if not isinstance(args, dict):
            return {}, "arguments must be a JSON object"
        missing = [k for k in self.parameters.get("required", []) if k not in args]
        if missing:
            return args, f"missing required argument(s): {', '.join(missing)}"
        return args, None


@dataclass
class ToolResponse:
    text: str = ""
    image: Any = None          # multimodal payloads are carried through untouched
    video: Any = None

    def is_empty(self) -> bool:
        return not self.text and self.image is None and self.video is None


class BaseTool:
    """Override ``execute`` (and optionally ``create`` / ``calc_reward`` / ``release``)."""

    def __init__(self, config: dict | None = None, schema: ToolSchema | None = None):
        self.config = config or {}
        self.schema = schema or self.default_schema()
        self.name = self.schema.name
        self._instances: dict[str, dict] = {}

    @classmethod
    def default_schema(cls) -> ToolSchema:
        return ToolSchema(name=cls.__name__.lower())

    def create(self, instance_id: str | None = None, **kwargs) -> tuple[str, ToolResponse]:
        iid = instance_id or uuid.uuid4().hex
        self._instances[iid] = dict(kwargs)
        return iid, ToolResponse()

    def execute(self, instance_id: str, parameters: dict[str, Any], **kwargs) -> tuple[ToolResponse, float, dict]:
        """Returns (response shown to the model, step reward, metrics)."""
        return ToolResponse(text="ok"), 0.0, {}

    def calc_reward(self, instance_id: str, **kwargs) -> float:
        return 0.0

    def release(self, instance_id: str, **kwargs) -> None:
        self._instances.pop(instance_id, None)


# ------------------------------------------------------------------------------------------------ built-in tools
_BIN = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
        ast.FloorDiv: operator.floordiv, ast.Mod: operator.mod, ast.Pow: operator.pow}
_UN = {ast.UAdd: operator.pos, ast.USub: operator.neg}


def safe_arith(expr: str) -> float:
    """Evaluate + - * / // % ** and parentheses over numbers — no names, calls or attribute access."""
    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return node.value
        if isinstance(node, ast.BinOp) and type(node.op) in _BIN:
            a, b = ev(node.left), ev(node.right)
            if isinstance(node.op, ast.Pow) and (abs(a) > 1e6 or abs(b) > 64):
                raise ValueError("exponent too large")
            return _BIN[type(node.op)](a, b)
        if isinstance(node, ast.UnaryOp) and type(node.op) in _UN:
            return _UN[type(node.op)](ev(node.operand))
        raise ValueError(f"unsupported expression element: {type(node).__name__}")
    return ev(ast.parse(expr.strip(), mode="eval"))


class CalculatorTool(BaseTool):
    @classmethod
    def default_schema(cls) -> ToolSchema:
        return ToolSchema(name="calculator", description="Evaluate an arithmetic expression.",
                          parameters={"type": "object", "required": ["expression"],
                                      "properties": {"expression": {"type": "string", "description": "e.g. (3+4)*5"}}})

    def execute(self, instance_id, parameters, **kwargs):
        try:
            val = safe_arith(str(parameters.get("expression", "")))
            return ToolResponse(text=f"{val:g}" if isinstance(val, float) else str(val)), 0.0, {"calc_ok": 1}
        except Exception as e:                                     # the model sees the error and may retry
            return ToolResponse(text=f"error: {e}"), 0.0, {"calc_ok": 0}


class Gsm8kTool(BaseTool):
    """`calc_gsm8k_reward(answer)`: tells the model whether its answer matches the ground truth; repeated
    non-improving submissions get a small penalty (same shaping as the reference tool)."""

    @classmethod
    def default_schema(cls) -> ToolSchema:
        return ToolSchema(name="calc_gsm8k_reward", description="Check a candidate final answer for the GSM8K problem.",
                          parameters={"type": "object", "required": ["answer"],
                                      "properties": {"answer": {"type": "string", "description": "the numeric answer"}}})

    def create(self, instance_id=None, ground_truth: str | None = None, **kwargs):
        iid, resp = super().create(instance_id, **kwargs)
        self._instances[iid] = {"ground_truth": ground_truth, "response": "", "reward": 0.0}
        return iid, resp

    def execute(self, instance_id, parameters, **kwargs):
        st = self._instances[instance_id]
        ans = str(parameters.get("answer", "")).strip()
        st["response"] = ans if ans.startswith("#### ") else "#### " + ans
        reward = self.calc_reward(instance_id)
        step = -0.05 if reward <= st["reward"] else 0.0
        st["reward"] = max(st["reward"], reward)
        return ToolResponse(text=f"Current parsed answer={ans!r} reward={reward}"), step, {}

    def calc_reward(self, instance_id, **kwargs) -> float:
        from ..data.reward_score import gsm8k

        st = self._instances[instance_id]
        return float(gsm8k.compute_score(st["response"], st["ground_truth"], method="flexible", format_score=0.0, score=1.0))


class CodeInterpreterTool(BaseTool):
    """`code_interpreter(code)`: run a Python snippet in the local sandbox and return its stdout / error (reference capability:
    verl/tools/sandbox_fusion_tools.py)."""

    @classmethod
    def default_schema(cls) -> ToolSchema:
        return ToolSchema(name="code_interpreter", description="Execute Python code and return what it prints.",
                          parameters={"type": "object", "required": ["code"],
                                      "properties": {"code": {"type": "string", "description": "Python source"}}})

    def execute(self, instance_id, parameters, **kwargs):
        from ..utils.sandbox import run_python

        r = run_python(str(parameters.get("code", "")), timeout_s=float(self.config.get("timeout_s", 5.0)),
                       mem_mb=int(self.config.get("memory_limit_mb", 1024)))
        text = r.stdout if r.status == "ok" else f"[{r.status}] {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}\n{r.stdout}"
        return ToolResponse(text=text.strip()), 0.0, {"sandbox_status": r.status, "sandbox_wall_s": r.wall_s}


class SearchTool(BaseTool):
    """`search(query_list)`: BM25 retrieval over a local passage corpus (JSONL rows with ``contents`` / ``text`` and optional
    ``title``), top-k passages per query — the offline stand-in for the retrieval service of the reference's search tool
    (verl/tools/search_tool.py, Search-R1 style multi-turn QA; reward: ``reward_score/qa_em.py``)."""

    @classmethod
    def default_schema(cls) -> ToolSchema:
        return ToolSchema(name="search", description="Search the knowledge corpus for passages relevant to the queries.",
                          parameters={"type": "object", "required": ["query_list"],
                                      "properties": {"query_list": {"type": "array", "items": {"type": "string"},
                                                                    "description": "one or more search queries"}}})

    def __init__(self, config: dict | None = None, schema: ToolSchema | None = None):
        super().__init__(config, schema)
        import math
        import re

        self._tok = lambda t: re.findall(r"[a-z0-9]+", t.lower())
        self.topk = int(self.config.get("topk", 3))
        docs = self.config.get("corpus") or []
        if isinstance(docs, str):
            with open(docs) as f:
                docs = [json.loads(l) for l in f if l.strip()]
        self.docs = [{"title": d.get("title", ""), "text": d.get("contents") or d.get("text") or ""} if isinstance(d, dict)
                     else {"title": "", "text": str(d)} for d in docs]
        self._tf = []
        df: dict[str, int] = {}
        for d in self.docs:
            toks = self._tok(d["title"] + " " + d["text"])
            tf: dict[str, int] = {}
            for t in toks:
                tf[t] = tf.get(t, 0) + 1
            self._tf.append((tf, len(toks)))
            for t in tf:
                df[t] = df.get(t, 0) + 1
        n = max(1, len(self.docs))
        self._idf = {t: math.log(1 + (n - c + 0.5) / (c + 0.5)) for t, c in df.items()}
        self._avg = sum(l for _, l in self._tf) / n if self._tf else 1.0

    def _bm25(self, query: str) -> list[tuple[float, int]]:
        k1, b = 1.5, 0.75
        q = self._tok(query)
        scores = []
        for i, (tf, ln) in enumerate(self._tf):
            s = 0.0
            for t in q:
                f = tf.get(t, 0)
                if f:
                    s += self._idf.get(t, 0.0) * f * (k1 + 1) / (f + k1 * (1 - b + b * ln / self._avg))
            if s > 0:
                scores.append((s, i))
        return sorted(scores, reverse=True)[: self.topk]

    def execute(self, instance_id, parameters, **kwargs):
        queries = parameters.get("query_list") or ([parameters["query"]] if parameters.get("query") else [])
        if isinstance(queries, str):
            queries = [queries]
        out = []
        for q in queries:
            hits = self._bm25(str(q))
            body = "\n".join(f"Doc {r + 1} (Title: {self.docs[i]['title']}) {self.docs[i]['text']}" for r, (_, i) in enumerate(hits))
            out.append(body or "no results")
        return ToolResponse(text="\n---\n".join(out)), 0.0, {"queries": len(queries)}


_TOOLS: dict[str, type] = {"search": SearchTool, "calculator": CalculatorTool, "calc_gsm8k_reward": Gsm8kTool, "gsm8k": Gsm8kTool,
                           "code_interpreter": CodeInterpreterTool}


def register_tool(name: str):
    def deco(cls):
        _TOOLS[name] = cls
        return cls
    return deco


def _resolve(class_name: str) -> type:
    if class_name in _TOOLS:
        return _TOOLS[class_name]
    mod, _, attr = class_name.rpartition(".")
    return getattr(importlib.import_module(mod), attr)


def load_tools(spec) -> list[BaseTool]:
    """``spec``: path to a YAML/JSON file or an in-memory dict/list shaped like the reference's tool config:
    ``{"tools": [{"class_name": "...", "config": {...}, "tool_schema": {...openai...}}]}``."""
    if spec is None:
        return []
    if isinstance(spec, str):
        import yaml

        with open(spec) as f:
            spec = yaml.safe_load(f)
    items = spec.get("tools", []) if isinstance(spec, dict) else list(spec)
    tools = []
    for it in items:
        if isinstance(it, BaseTool):
            tools.append(it)
            continue
        cls = _resolve(it.get("class_name") or it.get("name"))
        schema = ToolSchema.from_openai(it["tool_schema"]) if it.get("tool_schema") else None
        tools.append(cls(config=it.get("config") or {}, schema=schema))
    return tools
