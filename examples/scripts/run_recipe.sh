#!/bin/bash
# Any recipe trainer (dapo | sppo | spin | entropy | gspo) on N GPUs:  RECIPE=sppo NGPU=4 bash examples/scripts/run_recipe.sh key=value ...
set -e
python -m polyrl_b200.recipes.${RECIPE:-gspo} controller.backend=${BACKEND:-mp} trainer.n_gpus_per_node=${NGPU:-8} \
  actor_rollout_ref.model.path=${MODEL:-llama3-8b} "$@"
