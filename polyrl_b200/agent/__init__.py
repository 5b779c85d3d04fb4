"""Multi-turn / tool-using rollouts: tools, tool-call parsers, user-side interactions and the agent loops that drive
them against the native rollout engine."""
from .agent_loop import (AgentLoopOutput, MultiTurnRollout, SingleTurnAgentLoop, ToolAgentLoop, get_agent_loop_cls,  # noqa: F401
                         register_agent_loop)
from .interactions import BaseInteraction, Gsm8kInteraction, get_interaction_cls, register_interaction  # noqa: F401
from .tool_parser import FunctionCall, GptOssToolParser, HermesToolParser, ToolParser, get_tool_parser  # noqa: F401
from .tools import BaseTool, CalculatorTool, Gsm8kTool, ToolResponse, ToolSchema, load_tools  # noqa: F401
from . import remote_tools  # noqa: F401,E402 — registers mcp / sandbox_fusion / geo3k / image tools
