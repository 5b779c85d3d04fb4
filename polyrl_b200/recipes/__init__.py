"""Recipe trainers: algorithm-specific control flow on top of the single-controller PPO driver
(``trainer/controller_ppo.py``), mirroring the reference's ``recipe/`` tree (3rdparty/verl/recipe/{dapo,sppo,spin,entropy,gspo,flowrl,prime,collabllm}; the langgraph_agent recipe's graph-structured agents
live in ``agent/graph.py``).
The loss functions themselves are registered policy losses in ``algos/core_algos.py``; a recipe is the sampling / batching /
advantage policy around them plus an entry point with the recipe's defaults."""
from .dapo import DAPOTrainer
from .entropy import EntropyMechanismTrainer
from .flowrl import FlowRLTrainer
from .gspo import GSPOTrainer
from .prime import PRIMETrainer
from .spin import SPINTrainer
from .sppo import SPPOTrainer

RECIPES = {"prime": PRIMETrainer, "flowrl": FlowRLTrainer, "dapo": DAPOTrainer, "sppo": SPPOTrainer, "spin": SPINTrainer, "entropy": EntropyMechanismTrainer, "gspo": GSPOTrainer}

__all__ = ["PRIMETrainer", "FlowRLTrainer", "DAPOTrainer", "SPPOTrainer", "SPINTrainer", "EntropyMechanismTrainer", "GSPOTrainer", "RECIPES"]
