"""Graph-structured agents: a small state-graph executor and a ReAct agent loop built on it.

Capability parity: 3rdparty/verl/recipe/langgraph_agent/ (react_agent_loop.py: a LangGraph ``StateGraph`` with a model node, a
tool node and a conditional edge, driven through a chat-model wrapper around the rollout server — chat_model.py — so that every
model call of the graph becomes token-level training data).  LangGraph itself is not a dependency here: ``StateGraph`` below is
the subset agents of that shape need (named nodes, static and conditional edges, a message-appending state), and user code written
against it is plain Python.

B200-first execution: the graphs of ALL trajectories of a batch advance concurrently, one lightweight thread each, and every
``ctx.chat(messages)`` call of a node is a request into the SAME continuous batch of the rollout engine, served by the scheduler
thread that owns the engine (``GraphAgentLoop.generate``).  A node blocks only its own trajectory; the GPU keeps decoding the
others.  Tokens the engine produced are the training response (mask 1); whatever a graph appends between two model calls (tool
results, injected user turns) is tokenised as context (mask 0) — the same layout as ``MultiTurnRollout``.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from typing import Any, Callable

from .agent_loop import AgentLoopOutput, AgentState, MultiTurnRollout, _engine_mm, _Traj

END = "__end__"
START = "__start__"


class StateGraph:
    """Nodes are ``fn(state, ctx) -> dict | None``; the returned dict is merged into the state (``messages`` are APPENDED, other keys
    overwrite).  Edges: ``add_edge(a, b)`` or ``add_conditional_edges(a, router[, mapping])`` with ``router(state) -> key``."""

    def __init__(self):
        self.nodes: dict[str, Callable] = {}
        self.edges: dict[str, str] = {}
        self.routers: dict[str, tuple[Callable, dict | None]] = {}
        self.entry: str | None = None

    def add_node(self, name: str, fn: Callable) -> "StateGraph":
        if name in (END, START) or name in self.nodes:
            raise ValueError(f"invalid or duplicate node name {name!r}")
        self.nodes[name] = fn
        return self

    def add_edge(self, src: str, dst: str) -> "StateGraph":
        if src == START:
            self.entry = dst
        else:
            self.edges[src] = dst
        return self

    def add_conditional_edges(self, src: str, router: Callable, mapping: dict | None = None) -> "StateGraph":
        self.routers[src] = (router, mapping)
        return self

    def set_entry_point(self, name: str) -> "StateGraph":
        self.entry = name
        return self

    def compile(self) -> "CompiledGraph":
        if self.entry not in self.nodes:
            raise ValueError("the graph has no entry point")
        for src, dst in self.edges.items():
            if src not in self.nodes or (dst != END and dst not in self.nodes):
                raise ValueError(f"edge {src!r} -> {dst!r} references an unknown node")
        for src, (_, mapping) in self.routers.items():
            if src not in self.nodes or any(d != END and d not in self.nodes for d in (mapping or {}).values()):
                raise ValueError(f"conditional edges of {src!r} reference an unknown node")
        dangling = [n for n in self.nodes if n not in self.edges and n not in self.routers]
        if dangling:
            raise ValueError(f"nodes without an outgoing edge: {dangling} (use END)")
        return CompiledGraph(self)


class CompiledGraph:
    def __init__(self, g: StateGraph):
        self.g = g

    def invoke(self, state: dict, ctx: Any = None, recursion_limit: int = 64) -> dict:
        state = dict(state)
        state["messages"] = list(state.get("messages") or [])
        node = self.g.entry
        for _ in range(recursion_limit):
            upd = self.g.nodes[node](state, ctx) or {}
            for k, v in upd.items():
                if k == "messages":
                    state["messages"] = state["messages"] + list(v)
                else:
                    state[k] = v
            if node in self.g.routers:
                router, mapping = self.g.routers[node]
                key = router(state)
                node = (mapping or {}).get(key, key)
            else:
                node = self.g.edges[node]
            if node == END:
                return state
        raise RecursionError(f"graph did not reach END within {recursion_limit} steps")


def react_graph() -> CompiledGraph:
    """model -> (tool calls? -> tools -> model | END): the ReAct loop of the reference's example agent."""

    def model(state, ctx):
        msg = ctx.chat(state["messages"])
        return {"messages": [msg]} if msg is not None else {"exhausted": True}

    def tools(state, ctx):
        return {"messages": ctx.run_tools(state["messages"][-1].get("tool_calls") or [])}

    def route(state):
        last = state["messages"][-1] if state["messages"] else {}
        return "tools" if (not state.get("exhausted") and last.get("role") == "assistant" and last.get("tool_calls")) else END

    return (StateGraph().add_node("model", model).add_node("tools", tools).set_entry_point("model")
            .add_conditional_edges("model", route, {"tools": "tools", END: END}).add_edge("tools", "model").compile())


class GraphContext:
    """What a node sees of the runtime: ``chat`` (one model call through the shared engine), ``run_tools`` (the loop's tool registry
    with per-trajectory tool instances), ``sample`` (the dataset row's extras)."""

    def __init__(self, loop: "GraphAgentLoop", tr: _Traj):
        self._loop, self._tr = loop, tr
        self.sample = tr.extra

    def chat(self, messages: list[dict]) -> dict | None:
        """One assistant turn for the conversation ``messages`` (which must extend the conversation of the previous call: the
        new non-assistant messages are appended to the token context).  None once the turn / token budget is used up."""
        fut: Future = Future()
        self._loop._inbox.put((self._tr, list(messages), fut))
        return fut.result()

    def run_tools(self, calls: list) -> list[dict]:
        return self._loop._run_tools(self._tr, list(calls))


class GraphAgentLoop(MultiTurnRollout):
    """Runs ``graph_factory()`` (a ``CompiledGraph``, built once per loop) for every sample of the batch."""

    def __init__(self, engine, tokenizer, graph_factory: Callable[[], CompiledGraph] | str | None = None, recursion_limit: int = 64, **kw):
        super().__init__(engine, tokenizer, **kw)
        if isinstance(graph_factory, str):
            import importlib

            mod, _, attr = graph_factory.rpartition(".")
            graph_factory = getattr(importlib.import_module(mod), attr)
        self.graph = (graph_factory or react_graph)()
        self.recursion_limit = recursion_limit

    # ---- graph side (one thread per trajectory)
    def _run_graph(self, tr: _Traj) -> None:
        try:
            final = self.graph.invoke({"messages": list(tr.messages)}, GraphContext(self, tr), self.recursion_limit)
            tr.extra["_graph_state"] = {k: v for k, v in final.items() if k != "messages"}
        except Exception as e:                         # a crashing node ends its trajectory, not the batch
            tr.finish_reason = f"error: {type(e).__name__}: {e}"
        finally:
            self._inbox.put((tr, None, None))          # "this graph is done"

    # ---- scheduler side (owns the engine)
    def _serve_chat(self, tr: _Traj, messages: list[dict], fut: Future, sampling: dict) -> None:
        known = len(tr.messages)
        if len(messages) < known or any(a is not b and a != b for a, b in zip(messages[:known], tr.messages)):
            fut.set_exception(ValueError("ctx.chat: the conversation must extend the one of the previous call"))
            return
        new = messages[known:]
        if any(m.get("role") == "assistant" for m in new):
            fut.set_exception(ValueError("ctx.chat: assistant messages come from the model, they cannot be injected"))
            return
        if tr.assistant_turns >= self.max_assistant_turns or (new and not self._append_context(tr, new)) \
                or len(tr.response_ids) >= self.response_length:
            tr.finish_reason = "length"
            fut.set_result(None)
            return
        self._waiting[id(tr)] = fut
        self._submit(tr, sampling)

    def _finish_chat(self, tr: _Traj, req) -> None:
        tr.metrics["generate_s"] += time.time() - tr.t_submit
        room = self.response_length - len(tr.response_ids)
        out, lps = req.output_ids[:room], req.logprobs[:room]
        tr.response_ids += out
        tr.response_mask += [1] * len(out)
        tr.logprobs += lps
        tr.assistant_turns += 1
        text, calls = self.parser.extract_tool_calls(out)
        msg = {"role": "assistant", "content": text}
        if calls and self.tools:
            msg["tool_calls"] = calls
        tr.messages.append(msg)
        if req.finish_reason in ("length", "abort"):
            tr.finish_reason = req.finish_reason
        tr.state = AgentState.PENDING
        self._waiting.pop(id(tr)).set_result(msg)

    def generate(self, samples: list[dict], sampling: dict | None = None) -> list[AgentLoopOutput]:
        import torch

        sampling = dict(sampling or {})
        self._by_rid, self._waiting, self._inbox = {}, {}, queue.Queue()
        trajs = []
        for i, s in enumerate(samples):
            msgs = list(s.get("messages") or [])
            pids = list(map(int, s["prompt_ids"])) if s.get("prompt_ids") is not None else self.encode_prompt(msgs)
            trajs.append(_Traj(index=i, prompt_ids=pids, messages=msgs, extra=dict(s), mm=_engine_mm(s.get("multi_modal_inputs"))))
        threads = [threading.Thread(target=self._run_graph, args=(tr,), daemon=True) for tr in trajs]
        for t in threads:
            t.start()
        running = len(trajs)
        while running or self.engine.has_work():
            try:
                # block on the inbox only when there is nothing to decode
                item = self._inbox.get(timeout=0.0 if self.engine.has_work() else 0.05)
                while True:
                    tr, messages, fut = item
                    if messages is None:
                        running -= 1
                    else:
                        self._serve_chat(tr, messages, fut, sampling)
                    item = self._inbox.get_nowait()
            except queue.Empty:
                pass
            if self.engine.has_work():
                for req in self.engine.step():
                    tr = self._by_rid.pop(req.rid, None)
                    if tr is not None:
                        self._finish_chat(tr, req)
        for t in threads:
            t.join()
        self.last_messages = [tr.messages for tr in trajs]
        outs = []
        for tr in trajs:
            reason = tr.finish_reason
            self._terminate(tr)
            tr.finish_reason = reason
            outs.append(AgentLoopOutput(
                prompt_ids=tr.prompt_ids, response_ids=tr.response_ids, response_mask=tr.response_mask, response_logprobs=tr.logprobs,
                num_turns=tr.assistant_turns + tr.user_turns + 1 + tr.metrics["tool_calls"], tool_rewards=tr.tool_rewards,
                turn_scores=tr.turn_scores, metrics=tr.metrics, finish_reason=tr.finish_reason,
                reward_score=(tr.extra.get("_graph_state") or {}).get("reward_score"),
                multi_modal_inputs=None if tr.mm is None else {"pixel_values": tr.mm["pixel_values"], "image_grid_thw": torch.tensor(tr.mm["grid_thw"])}))
        return outs


class ReactAgentLoop(GraphAgentLoop):
    """``react_agent``: the ReAct graph (model <-> tools) — reference name: recipe/langgraph_agent/react_agent_loop.py."""

    def __init__(self, engine, tokenizer, **kw):
        kw.pop("graph_factory", None)
        super().__init__(engine, tokenizer, graph_factory=react_graph, **kw)
