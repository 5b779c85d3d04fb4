"""User-side interactions for multi-turn rollouts (an environment / simulated user that answers the assistant).

Capability parity: 3rdparty/verl/verl/interactions/base.py:20-72 (start / generate_response / calculate_score /
finalize) and interactions/gsm8k_interaction.py (answer feedback until correct)."""
from __future__ import annotations

import importlib
import uuid
from typing import Any


class BaseInteraction:
    def __init__(self, config: dict | None = None):
        self.config = config or {}
        self.name = self.config.get("name", type(self).__name__.lower())
        self._instances: dict[str, dict] = {}

    def start_interaction(self, instance_id: str | None = None, **kwargs) -> str:
        iid = instance_id or uuid.uuid4().hex
        self._instances[iid] = dict(kwargs)
        return iid

    def generate_response(self, instance_id: str, messages: list[dict[str, Any]], **kwargs) -> tuple[bool, str, float, dict]:
        """Returns (should_terminate, user-turn text, turn score, extra)."""
        return True, "", 0.0, {}

    def calculate_score(self, instance_id: str, **kwargs) -> float:
        return 0.0

    def finalize_interaction(self, instance_id: str, **kwargs) -> None:
        self._instances.pop(instance_id, None)


class Gsm8kInteraction(BaseInteraction):
    def start_interaction(self, instance_id=None, ground_truth: str | None = None, **kwargs):
        iid = super().start_interaction(instance_id, **kwargs)
        self._instances[iid] = {"ground_truth": ground_truth, "response": "", "reward": 0.0}
        return iid

    def generate_response(self, instance_id, messages, **kwargs):
        last = next((m["content"] for m in reversed(messages) if m.get("role") == "assistant"), "")
        self._instances[instance_id]["response"] = last
        reward = self.calculate_score(instance_id)
        if reward >= 1.0:
            return True, "Your response is correct!", 1.0, {}
        return False, "Your response is incorrect! You need to reflect on your answer and try again.", 0.0, {}

    def calculate_score(self, instance_id, **kwargs) -> float:
        from ..data.reward_score import gsm8k

        st = self._instances[instance_id]
        return float(gsm8k.compute_score(st["response"], st["ground_truth"], method="strict", format_score=0.0, score=1.0))


class WeatherInteraction(BaseInteraction):
    """Demo environment for tool-use trajectories (reference: verl/interactions/weather_interaction.py): the simulated user is
    satisfied — and the turn scores 1 — once the assistant's reply contains a weather report (a temperature reading), otherwise
    it asks again, up to ``max_turns`` (config)."""

    def start_interaction(self, instance_id=None, **kwargs):
        iid = super().start_interaction(instance_id, **kwargs)
        self._instances[iid] = {"turns": 0, "reward": 0.0, **kwargs}
        return iid

    def generate_response(self, instance_id, messages, **kwargs):
        import re

        st = self._instances[instance_id]
        st["turns"] += 1
        last = next((str(m["content"]) for m in reversed(messages) if m.get("role") == "assistant"), "")
        if re.search(r"-?\d+(\.\d+)?\s*(°|degrees|deg)\s*[CF]?", last, flags=re.I):
            st["reward"] = 1.0
            return True, "Thanks, that is the weather report I was looking for.", 1.0, {}
        if st["turns"] >= int(self.config.get("max_turns", 3)):
            return True, "Never mind.", 0.0, {}
        return False, "I still need the current temperature. Please call the weather tool and report it.", 0.0, {}

    def calculate_score(self, instance_id, **kwargs) -> float:
        return float(self._instances[instance_id]["reward"])


_INTERACTIONS = {"gsm8k": Gsm8kInteraction, "weather": WeatherInteraction, "base": BaseInteraction}


def register_interaction(name: str, cls) -> None:
    """Make an interaction class loadable by a short name in ``interaction_config_path`` files."""
    _INTERACTIONS[name] = cls


_LAZY_INTERACTIONS = {"collabllm": "polyrl_b200.recipes.collabllm.CollabLLMInteraction"}      # registered by their recipe modules


def get_interaction_cls(name: str):
    if name in _INTERACTIONS:
        return _INTERACTIONS[name]
    name = _LAZY_INTERACTIONS.get(name, name)
    mod, _, attr = name.rpartition(".")
    return getattr(importlib.import_module(mod), attr)


def load_interactions(spec) -> dict[str, BaseInteraction]:
    """``{"interaction": [{"name": ..., "class_name": ..., "config": {...}}]}`` -> {name: instance}"""
    if spec is None:
        return {}
    if isinstance(spec, str):
        import yaml

        with open(spec) as f:
            spec = yaml.safe_load(f)
    out = {}
    for it in (spec.get("interaction", []) if isinstance(spec, dict) else spec):
        cfg = dict(it.get("config") or {})
        cfg.setdefault("name", it.get("name") or it.get("class_name"))
        out[cfg["name"]] = get_interaction_cls(it.get("class_name") or it["name"])(cfg)
    return out
