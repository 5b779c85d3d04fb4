"""Agent loops: multi-turn trajectories (assistant turn -> tool calls / simulated user -> assistant turn ...)
executed INSIDE the rollout engine's continuous batch.

Design (B200-first): the reference runs one asyncio coroutine per trajectory, each talking HTTP to an inference
server (3rdparty/verl/verl/experimental/agent_loop/tool_agent_loop.py:81-478 state machine PENDING -> GENERATING ->
PROCESSING_TOOLS / INTERACTING -> TERMINATED; request-level path of sglang_rollout.py:807-1497).  Here one
scheduler owns every trajectory of the batch: finished turns are parsed on the host while the GPU keeps decoding
the other sequences, tool calls run on a thread pool, and a trajectory re-enters the continuous batch as soon as
its tool response is tokenised.  Tool / user tokens are part of ``responses`` with ``response_mask = 0`` (no loss,
log-prob 0), exactly the layout the trainer expects (Appendix D of SURVEY.md).
"""
from __future__ import annotations

import json
import time
from concurrent.futures import Future, ThreadPoolExecutor
from dataclasses import dataclass, field
from enum import Enum
from typing import Any

import numpy as np
import torch

from ..protocol import DataProto
from .interactions import BaseInteraction
from .tool_parser import FunctionCall, get_tool_parser
from .tools import BaseTool, ToolResponse
from ..utils.rollout_trace import rollout_trace_attr, rollout_trace_op


class AgentState(str, Enum):
    PENDING = "pending"
    GENERATING = "generating"
    PROCESSING_TOOLS = "processing_tools"
    INTERACTING = "interacting"
    TERMINATED = "terminated"


@dataclass
class AgentLoopOutput:
    prompt_ids: list[int]
    response_ids: list[int]
    response_mask: list[int]
    response_logprobs: list[float]
    num_turns: int = 1
    reward_score: float | None = None
    tool_rewards: list[float] = field(default_factory=list)
    turn_scores: list[float] = field(default_factory=list)
    metrics: dict = field(default_factory=dict)
    finish_reason: str = "stop"
    # vision-language trajectories: every image of the sequence in order — the prompt's and those returned by tools
    multi_modal_inputs: dict | None = None


@dataclass
class _Traj:
    index: int
    prompt_ids: list[int]
    messages: list[dict]
    extra: dict
    state: AgentState = AgentState.PENDING
    response_ids: list[int] = field(default_factory=list)
    response_mask: list[int] = field(default_factory=list)
    logprobs: list[float] = field(default_factory=list)
    assistant_turns: int = 0
    user_turns: int = 0
    tool_rewards: list[float] = field(default_factory=list)
    turn_scores: list[float] = field(default_factory=list)
    tool_instances: dict = field(default_factory=dict)
    interaction_id: str | None = None
    finish_reason: str = "stop"
    metrics: dict = field(default_factory=lambda: {"generate_s": 0.0, "tool_s": 0.0, "tool_calls": 0})
    t_submit: float = 0.0
    mm: dict | None = None                    # {"pixel_values": [N, patch_dim], "grid_thw": [(t, h, w), ...]} so far


def _engine_mm(mm: dict | None) -> dict | None:
    """Dataset-style ``multi_modal_inputs`` (``image_grid_thw`` tensor) -> the engine's {"pixel_values", "grid_thw": [tuples]}."""
    if mm is None:
        return None
    grid = mm["image_grid_thw"] if "image_grid_thw" in mm else mm["grid_thw"]
    grid = grid.tolist() if hasattr(grid, "tolist") else list(grid)
    return {"pixel_values": mm["pixel_values"], "grid_thw": [tuple(int(x) for x in g) for g in grid]}


def _truncate(text: str, limit: int, side: str) -> str:
    if limit <= 0 or len(text) <= limit:
        return text
    if side == "left":
        return text[:limit] + "...(truncated)"
    if side == "right":
        return "(truncated)..." + text[-limit:]
    half = limit // 2
    return text[:half] + "...(truncated)..." + text[-half:]


class MultiTurnRollout:
    """Batched tool/interaction agent loop over a ``RolloutEngine``.

    ``sample``s are dicts with ``prompt_ids`` (or ``messages``), optional ``tools_kwargs`` ({tool_name:
    {"create_kwargs": {...}}}), ``interaction_kwargs`` ({"name": ..., ...})."""

    def __init__(self, engine, tokenizer, tools: list[BaseTool] | None = None,
                 interactions: dict[str, BaseInteraction] | None = None, parser: str = "hermes",
                 response_length: int = 512, max_assistant_turns: int = 4, max_user_turns: int = 4,
                 max_parallel_calls: int = 1, max_tool_response_length: int = 256,
                 tool_response_truncate_side: str = "middle", tool_workers: int = 8, model_cfg=None,
                 image_min_pixels: int | None = None, image_max_pixels: int | None = None):
        self.engine, self.tok = engine, tokenizer
        # vision-language policies: images returned by tools become image-token runs in the context (reference: the deepeyes /
        # image_zoom_in_tool path of tool_agent_loop.py) — needs the model config for the patch geometry and token ids
        self.vl_cfg = model_cfg if (model_cfg is not None and getattr(model_cfg, "vision", None) is not None) else None
        self.image_pixels = (image_min_pixels, image_max_pixels)
        self.tools = {t.name: t for t in (tools or [])}
        self.tool_schemas = [t.schema.to_openai() for t in (tools or [])]
        self.interactions = interactions or {}
        self.parser = get_tool_parser(parser, tokenizer)
        self.response_length = response_length
        self.max_assistant_turns, self.max_user_turns = max_assistant_turns, max_user_turns
        self.max_parallel_calls = max_parallel_calls
        self.max_tool_response_length = max_tool_response_length
        self.truncate_side = tool_response_truncate_side
        self.pool = ThreadPoolExecutor(max_workers=tool_workers)
        # token delta of an appended non-assistant turn = template(base + msgs) - template(base)
        self._base = [{"role": "system", "content": "x"}]
        self._base_len = len(self._template(self._base, add_generation_prompt=False))

    # ------------------------------------------------------------------ tokenisation helpers
    def _template(self, messages, add_generation_prompt: bool, tools=None) -> list[int]:
        try:
            out = self.tok.apply_chat_template(messages, add_generation_prompt=add_generation_prompt, tokenize=True,
                                               **({"tools": tools} if tools else {}))
        except TypeError:
            out = self.tok.apply_chat_template(messages, add_generation_prompt=add_generation_prompt, tokenize=True)
        if isinstance(out, dict):
            out = out["input_ids"]
        return list(out)

    def encode_turns(self, messages: list[dict]) -> list[int]:
        """Tokens of tool / user turns followed by the next assistant generation prompt."""
        return self._template(self._base + messages, add_generation_prompt=True)[self._base_len:]

    def _encode_turns_with_images(self, msgs: list[dict]):
        """Like ``encode_turns`` for turns that carry an image: the turn's text gets an ``<image>`` mark, the rendered text is split
        at the marks and each mark becomes ``vision_start + image_pad * n + vision_end`` -> (ids, patch tensors, grids)."""
        from ..models import vision as V

        cfg, mark = self.vl_cfg, "<image>"
        marked = [{"role": m["role"], "content": (m["content"] + mark) if m.get("image") is not None else m["content"]} for m in msgs]
        base = self.tok.apply_chat_template(self._base, add_generation_prompt=False, tokenize=False)
        text = self.tok.apply_chat_template(self._base + marked, add_generation_prompt=True, tokenize=False)[len(base):]
        parts = text.split(mark)
        images = [m["image"] for m in msgs if m.get("image") is not None]
        if len(parts) - 1 != len(images):
            raise ValueError("chat template dropped or duplicated an <image> mark")
        ids = list(self.tok.encode(parts[0], add_special_tokens=False))
        pix, grids = [], []
        for img, seg in zip(images, parts[1:]):
            pv, g = V.preprocess_image(cfg, img, *self.image_pixels)
            pix.append(pv)
            grids.append(g)
            ids += [cfg.vision_start_token_id] + [cfg.image_token_id] * V.num_image_tokens(cfg, [g])[0] + [cfg.vision_end_token_id]
            ids += list(self.tok.encode(seg, add_special_tokens=False))
        return ids, pix, grids

    def encode_prompt(self, messages: list[dict]) -> list[int]:
        return self._template(messages, add_generation_prompt=True, tools=self.tool_schemas or None)

    # ------------------------------------------------------------------ engine plumbing
    def _submit(self, tr: _Traj, sampling: dict) -> None:
        budget = self.response_length - len(tr.response_ids)
        sp = dict(sampling)
        sp["max_new_tokens"] = budget
        sp["n"] = 1
        tr.state = AgentState.GENERATING
        tr.t_submit = time.time()
        if tr.mm is not None:
            req = self.engine.add_request(tr.prompt_ids + tr.response_ids, sp, mm=tr.mm)
        else:
            req = self.engine.add_request(tr.prompt_ids + tr.response_ids, sp)
        self._by_rid[req.rid] = tr

    def _terminate(self, tr: _Traj, reason: str | None = None) -> None:
        tr.state = AgentState.TERMINATED
        if reason:
            tr.finish_reason = reason
        for name, iid in tr.tool_instances.items():
            try:
                tr.tool_rewards.append(float(self.tools[name].calc_reward(iid)))
                self.tools[name].release(iid)
            except Exception:
                pass
        if tr.interaction_id is not None:
            it = self.interactions.get(tr.extra.get("interaction_kwargs", {}).get("name", ""))
            if it is not None:
                it.finalize_interaction(tr.interaction_id)

    @rollout_trace_op
    def _call_tool(self, tr: _Traj, call: FunctionCall) -> tuple[ToolResponse, float, dict]:
        tool = self.tools.get(call.name)
        if tool is None:
            return ToolResponse(text=f"error: unknown tool {call.name!r}"), 0.0, {}
        args, err = tool.schema.validate_arguments(call.arguments)
        if err:
            return ToolResponse(text=f"error: {err}"), 0.0, {}
        if call.name not in tr.tool_instances:
            kw = (tr.extra.get("tools_kwargs", {}).get(call.name, {}) or {}).get("create_kwargs", {})
            tr.tool_instances[call.name], _ = tool.create(**kw)
        try:
            resp, reward, metrics = tool.execute(tr.tool_instances[call.name], args)
        except Exception as e:
            return ToolResponse(text=f"error when executing tool: {e}"), 0.0, {}
        resp.text = _truncate(resp.text or "", self.max_tool_response_length, self.truncate_side)
        return resp, float(reward or 0.0), metrics or {}

    def _run_tools(self, tr: _Traj, calls: list[FunctionCall]) -> list[dict]:
        t0 = time.time()
        msgs = []
        for c in calls[: self.max_parallel_calls]:
            with rollout_trace_attr(sample_index=tr.index, name="tool_call"):
                resp, reward, _ = self._call_tool(tr, c)
            tr.tool_rewards.append(reward)
            msg = {"role": "tool", "content": resp.text}
            if resp.image is not None and self.vl_cfg is not None:
                msg["image"] = resp.image
            msgs.append(msg)
        tr.metrics["tool_s"] += time.time() - t0
        tr.metrics["tool_calls"] += len(msgs)
        return msgs

    def _append_context(self, tr: _Traj, msgs: list[dict]) -> bool:
        """Append non-assistant turns (mask 0).  False when the response budget is exhausted."""
        images = [m["image"] for m in msgs if m.get("image") is not None]
        if images:
            ids, pix, grids = self._encode_turns_with_images(msgs)
        else:
            ids = self.encode_turns(msgs)
        room = self.response_length - len(tr.response_ids)
        if len(ids) >= room:
            return False
        if images:
            if tr.mm is None:
                tr.mm = {"pixel_values": torch.cat(pix), "grid_thw": list(grids)}
            else:
                tr.mm = {"pixel_values": torch.cat([tr.mm["pixel_values"], *pix]), "grid_thw": list(tr.mm["grid_thw"]) + list(grids)}
            msgs = [{k: v for k, v in m.items() if k != "image"} for m in msgs]
        tr.response_ids += ids
        tr.response_mask += [0] * len(ids)
        tr.logprobs += [0.0] * len(ids)
        tr.messages += msgs
        return True

    def _after_generation(self, tr: _Traj, req, sampling: dict, futures: dict) -> None:
        tr.metrics["generate_s"] += time.time() - tr.t_submit
        room = self.response_length - len(tr.response_ids)
        out, lps = req.output_ids[:room], req.logprobs[:room]
        tr.response_ids += out
        tr.response_mask += [1] * len(out)
        tr.logprobs += lps
        tr.assistant_turns += 1
        text, calls = self.parser.extract_tool_calls(out)
        tr.messages.append({"role": "assistant", "content": text})          # also when cut short: recipes read the conversation
        if req.finish_reason in ("length", "abort") or len(tr.response_ids) >= self.response_length:
            return self._terminate(tr, req.finish_reason if req.finish_reason != "stop" else "length")
        if calls and self.tools and tr.assistant_turns < self.max_assistant_turns:
            tr.state = AgentState.PROCESSING_TOOLS
            futures[self.pool.submit(self._run_tools, tr, calls)] = tr
            return
        ik = tr.extra.get("interaction_kwargs") or {}
        it = self.interactions.get(ik.get("name", "")) if ik else None
        if it is not None and tr.user_turns < self.max_user_turns and tr.assistant_turns < self.max_assistant_turns:
            tr.state = AgentState.INTERACTING
            if tr.interaction_id is None:
                tr.interaction_id = it.start_interaction(**{k: v for k, v in ik.items() if k != "name"})
            stop, reply, score, _ = it.generate_response(tr.interaction_id, tr.messages)
            tr.turn_scores.append(float(score))
            tr.user_turns += 1
            if stop or not self._append_context(tr, [{"role": "user", "content": reply}]):
                return self._terminate(tr)
            return self._submit(tr, sampling)
        self._terminate(tr)

    # ------------------------------------------------------------------ public
    def generate(self, samples: list[dict], sampling: dict | None = None) -> list[AgentLoopOutput]:
        sampling = dict(sampling or {})
        self._by_rid: dict[str, _Traj] = {}
        trajs = []
        for i, s in enumerate(samples):
            msgs = list(s.get("messages") or [])
            pids = list(map(int, s["prompt_ids"])) if s.get("prompt_ids") is not None else self.encode_prompt(msgs)
            mm0 = _engine_mm(s.get("multi_modal_inputs"))
            trajs.append(_Traj(index=i, prompt_ids=pids, messages=msgs, extra=s, mm=mm0, interaction_id=s.get("_interaction_id")))
        for tr in trajs:
            self._submit(tr, sampling)
        futures: dict[Future, _Traj] = {}
        while self.engine.has_work() or futures:
            if self.engine.has_work():
                for req in self.engine.step():
                    tr = self._by_rid.pop(req.rid, None)
                    if tr is not None:
                        self._after_generation(tr, req, sampling, futures)
            elif futures:
                next(iter(futures)).result()              # nothing to decode: block on a tool
            for fut in [f for f in futures if f.done()]:
                tr = futures.pop(fut)
                try:
                    msgs = fut.result()
                except Exception as e:                      # a crashing tool ends the trajectory, not the batch
                    msgs = [{"role": "tool", "content": f"error: {e}"}]
                if self._append_context(tr, msgs):
                    self._submit(tr, sampling)
                else:
                    self._terminate(tr, "length")
        outs = []
        self.last_messages = [tr.messages for tr in trajs]           # final conversation of every trajectory (recipes score them)
        for tr in trajs:
            if tr.state != AgentState.TERMINATED:
                self._terminate(tr)
            outs.append(AgentLoopOutput(prompt_ids=tr.prompt_ids, response_ids=tr.response_ids, response_mask=tr.response_mask,
                                        response_logprobs=tr.logprobs, num_turns=tr.assistant_turns + tr.user_turns + 1
                                        + tr.metrics["tool_calls"], tool_rewards=tr.tool_rewards, turn_scores=tr.turn_scores,
                                        metrics=tr.metrics, finish_reason=tr.finish_reason,
                                        multi_modal_inputs=None if tr.mm is None else
                                        {"pixel_values": tr.mm["pixel_values"], "image_grid_thw": torch.tensor(tr.mm["grid_thw"])}))
        return outs


class ToolAgentLoop(MultiTurnRollout):
    """Name used by the reference's registry (`tool_agent`)."""


class SingleTurnAgentLoop(MultiTurnRollout):
    """`single_turn_agent`: one assistant turn, no tools (reference: single_turn_agent_loop.py)."""

    def __init__(self, engine, tokenizer, response_length: int = 512, **_):
        super().__init__(engine, tokenizer, tools=None, interactions=None, response_length=response_length, max_assistant_turns=1)


def outputs_to_dataproto(outs: list[AgentLoopOutput], prompt_length: int, response_length: int, pad_id: int,
                         non_tensors: dict[str, Any] | None = None, model_cfg=None) -> DataProto:
    """Left-pad prompts / right-pad responses into the trainer batch layout (keys of SURVEY.md Appendix D)."""
    b = len(outs)
    prompts = torch.full((b, prompt_length), pad_id, dtype=torch.long)
    p_mask = torch.zeros(b, prompt_length, dtype=torch.long)
    resp = torch.full((b, response_length), pad_id, dtype=torch.long)
    r_att = torch.zeros(b, response_length, dtype=torch.long)
    r_loss = torch.zeros(b, response_length, dtype=torch.long)
    lps = torch.zeros(b, response_length, dtype=torch.float32)
    for i, o in enumerate(outs):
        p = o.prompt_ids[-prompt_length:]
        prompts[i, prompt_length - len(p):] = torch.tensor(p, dtype=torch.long)
        p_mask[i, prompt_length - len(p):] = 1
        n = min(len(o.response_ids), response_length)
        resp[i, :n] = torch.tensor(o.response_ids[:n], dtype=torch.long)
        r_att[i, :n] = 1
        r_loss[i, :n] = torch.tensor(o.response_mask[:n], dtype=torch.long)
        lps[i, :n] = torch.tensor(o.response_logprobs[:n], dtype=torch.float32)
    attention_mask = torch.cat([p_mask, r_att], dim=1)
    position_ids = (attention_mask.cumsum(-1) - 1).clamp_min(0)
    vl = model_cfg is not None and getattr(model_cfg, "vision", None) is not None
    if vl:
        # 3-axis M-RoPE positions over prompt + response: tool images inside the response shift everything after them
        from ..models.vision import mrope_positions

        pos3 = position_ids.unsqueeze(1).repeat(1, 3, 1)
        for i, o in enumerate(outs):
            if o.multi_modal_inputs is None:
                continue
            p = o.prompt_ids[-prompt_length:]
            if len(p) < len(o.prompt_ids):
                raise ValueError("an image prompt does not fit prompt_length")
            n = min(len(o.response_ids), response_length)
            grids = [tuple(int(x) for x in g) for g in o.multi_modal_inputs["image_grid_thw"].tolist()]
            seq = p + o.response_ids[:n]
            n_img_tok = sum(t in (model_cfg.image_token_id, model_cfg.video_token_id) for t in seq)
            m2 = model_cfg.vision.get("spatial_merge_size", 2) ** 2
            keep, acc = [], 0
            for g in grids:                                   # a truncated response may have lost its last image run
                k = g[0] * g[1] * g[2] // m2
                if acc + k > n_img_tok:
                    break
                keep.append(g)
                acc += k
            if acc != n_img_tok:
                raise ValueError("response truncation cut through an image token run")
            pm, _ = mrope_positions(model_cfg, seq, keep)
            pos3[i, :, prompt_length - len(p): prompt_length + n] = pm
        position_ids = pos3
    nt = {"__num_turns__": np.array([o.num_turns for o in outs], dtype=np.int64),
          "tool_rewards": np.array([json.dumps(o.tool_rewards) for o in outs], dtype=object),
          "turn_scores": np.array([json.dumps(o.turn_scores) for o in outs], dtype=object)}
    if vl or any(o.multi_modal_inputs is not None for o in outs):
        mm_arr = np.empty(b, dtype=object)
        for i, o in enumerate(outs):
            mm_arr[i] = o.multi_modal_inputs
        nt["multi_modal_inputs"] = mm_arr
    nt.update(non_tensors or {})
    tensors = {"prompts": prompts, "responses": resp, "input_ids": torch.cat([prompts, resp], dim=1), "attention_mask": attention_mask,
               "position_ids": position_ids, "response_mask": r_loss, "rollout_log_probs": lps}
    if any(o.reward_score is not None for o in outs):
        # an agent loop that scores its own trajectories (e.g. recipes/collabllm.py): the scalar sits on the last response token,
        # the layout reward models use (``rm_scores``), so the reward managers / trainers pick it up unchanged
        rm = torch.zeros(b, response_length, dtype=torch.float32)
        for i, o in enumerate(outs):
            n = min(len(o.response_ids), response_length)
            if o.reward_score is not None and n:
                rm[i, n - 1] = float(o.reward_score)
        tensors["rm_scores"] = rm
        keys = sorted({k for o in outs for k, v in o.metrics.items() if "/" in k and isinstance(v, (int, float))})
        for k in keys:                                            # recipe metrics ride along per sample ("collabllm/accuracy", ...)
            nt[k] = np.array([float(o.metrics.get(k, 0.0)) for o in outs], dtype=np.float64)
    return DataProto.from_dict(tensors, non_tensors=nt)


_AGENT_LOOPS: dict[str, Any] = {}


def register_agent_loop(name: str, target) -> None:
    """``target``: a MultiTurnRollout subclass or a dotted path to one (imported on first use)."""
    _AGENT_LOOPS[name] = target


def get_agent_loop_cls(name: str | None):
    """Agent-loop class by registry name (``tool_agent``, ``single_turn_agent``, ``collabllm_agent``, ``react_agent``) or dotted path
    (reference: ``rollout.agent.default_agent_loop`` + the ``_target_`` entries of agent_loop_config_path files)."""
    import importlib

    target = _AGENT_LOOPS.get(name or "tool_agent", name)
    if isinstance(target, str):
        mod, _, attr = target.rpartition(".")
        if not mod:
            raise KeyError(f"unknown agent loop {name!r} (registered: {sorted(_AGENT_LOOPS)})")
        target = getattr(importlib.import_module(mod), attr)
    return target


register_agent_loop("tool_agent", ToolAgentLoop)
register_agent_loop("single_turn_agent", SingleTurnAgentLoop)
register_agent_loop("collabllm_agent", "polyrl_b200.recipes.collabllm.CollabLLMAgentLoop")
register_agent_loop("react_agent", "polyrl_b200.agent.graph.ReactAgentLoop")
