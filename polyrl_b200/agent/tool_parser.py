"""Extract tool calls from generated text.

Capability parity: 3rdparty/verl/verl/experimental/agent_loop/tool_parser.py:29-160 (FunctionCall, ToolParser
registry, Hermes ``<tool_call>{json}</tool_call>`` format)."""
from __future__ import annotations

import json
import re
from dataclasses import dataclass


@dataclass
class FunctionCall:
    name: str
    arguments: str          # JSON string (validated later against the tool schema)


class ToolParser:
    _registry: dict[str, type] = {}

    def __init__(self, tokenizer):
        self.tokenizer = tokenizer

    def extract_tool_calls(self, response_ids: list[int]) -> tuple[str, list[FunctionCall]]:
        """Returns (assistant text with the tool-call markup removed, calls)."""
        raise NotImplementedError

    @classmethod
    def register(cls, name: str):
        def deco(sub):
            cls._registry[name] = sub
            return sub
        return deco


@ToolParser.register("hermes")
class HermesToolParser(ToolParser):
    OPEN, CLOSE = "<tool_call>", "</tool_call>"
    _rx = re.compile(r"<tool_call>(.*?)</tool_call>", re.DOTALL)

    def extract_tool_calls(self, response_ids):
        text = self.tokenizer.decode(response_ids, skip_special_tokens=False)
        if self.OPEN not in text or self.CLOSE not in text:
            return text, []
        calls = []
        for m in self._rx.findall(text):
            try:
                obj = json.loads(m)
                calls.append(FunctionCall(name=str(obj["name"]), arguments=json.dumps(obj.get("arguments", {}), ensure_ascii=False)))
            except Exception:
                continue                 # malformed call: ignored, like a non-call
        return self._rx.sub("", text), calls


@ToolParser.register("json_block")
class JsonBlockToolParser(ToolParser):
    """```json {"tool": name, "args": {...}} ``` fenced blocks — a simple alternative format."""
    _rx = re.compile(r"```json\s*(\{.*?\})\s*```", re.DOTALL)

    def extract_tool_calls(self, response_ids):
        text = self.tokenizer.decode(response_ids, skip_special_tokens=False)
        calls = []
        for m in self._rx.findall(text):
            try:
                obj = json.loads(m)
                calls.append(FunctionCall(name=str(obj.get("tool") or obj["name"]),
                                          arguments=json.dumps(obj.get("args", obj.get("arguments", {})))))
            except Exception:
                continue
        return self._rx.sub("", text), calls


def get_tool_parser(name: str, tokenizer) -> ToolParser:
    if name not in ToolParser._registry:
        raise ValueError(f"unknown tool parser {name!r} (have {sorted(ToolParser._registry)})")
    return ToolParser._registry[name](tokenizer)
