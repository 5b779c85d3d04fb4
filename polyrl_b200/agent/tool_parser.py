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


@ToolParser.register("gpt-oss")
class GptOssToolParser(ToolParser):
    """OpenAI harmony format (gpt-oss models): an assistant turn is a run of ``<|start|>assistant<|channel|>CH ...<|message|>BODY``
    segments closed by ``<|end|>`` / ``<|call|>`` / ``<|return|>``.  A tool call is a segment addressed ``to=functions.NAME`` whose
    body is the JSON arguments and which ends with ``<|call|>``.  Reasoning in the ``analysis`` channel may itself mention tool-call
    markup, so it is removed before calls are searched (also when the leading ``<|start|>assistant`` belongs to the prompt).
    Capability parity: 3rdparty/verl/verl/experimental/agent_loop/tool_parser.py:110-160 (GptOssToolParser); arguments are passed on
    unvalidated, the tool schema check happens at execution."""
    _cot = re.compile(r"(?:<\|start\|>assistant)?<\|channel\|>analysis<\|message\|>.*?<\|end\|>", re.DOTALL)
    # header = everything between the role and <|message|>: channel name, recipient (before or after the channel) and content type
    _call = re.compile(r"(?:<\|start\|>assistant)?(?P<head>(?:(?!<\|message\|>|<\|start\|>).)*?to=functions\.(?P<name>[^\s<]+)(?:(?!<\|message\|>|<\|start\|>).)*?)"
                       r"<\|message\|>(?P<args>.*?)<\|call\|>", re.DOTALL)

    def extract_tool_calls(self, response_ids):
        text = self.tokenizer.decode(response_ids, skip_special_tokens=False)
        pad = getattr(self.tokenizer, "pad_token", None)
        if pad:
            text = text.replace(pad, "")
        text = self._cot.sub("", text)
        calls = [FunctionCall(name=m.group("name"), arguments=m.group("args").strip()) for m in self._call.finditer(text)
                 if "<|channel|>" in m.group("head")]
        if not calls:
            return text, []
        return self._call.sub(lambda m: "" if "<|channel|>" in m.group("head") else m.group(0), text), calls


def get_tool_parser(name: str, tokenizer) -> ToolParser:
    if name not in ToolParser._registry:
        raise ValueError(f"unknown tool parser {name!r} (have {sorted(ToolParser._registry)})")
    return ToolParser._registry[name](tokenizer)
