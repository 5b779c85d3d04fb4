#!/bin/bash
# DAPO recipe (dynamic sampling + clip-higher + token-mean loss + over-long shaping) through the single-controller driver:
# one worker process per GPU (controller.backend=mp).  Counterpart of 3rdparty/verl/recipe/dapo/run_dapo_qwen2.5_32b.sh.
set -e
NGPU=${NGPU:-8}
python -m polyrl_b200.recipes.dapo controller.backend=mp trainer.n_gpus_per_node=$NGPU \
  actor_rollout_ref.model.path=${MODEL:-Qwen/Qwen2.5-7B} \
  data.train_files=${TRAIN:-data/dapo-math-17k.parquet} data.val_files=${VAL:-data/aime-2024.parquet} \
  data.train_batch_size=512 data.gen_batch_size=1536 data.max_prompt_length=2048 data.max_response_length=8192 \
  actor_rollout_ref.rollout.n=16 actor_rollout_ref.rollout.temperature=1.0 \
  actor_rollout_ref.actor.ppo_mini_batch_size=32 actor_rollout_ref.actor.use_dynamic_bsz=True \
  actor_rollout_ref.actor.ppo_max_token_len_per_gpu=20480 actor_rollout_ref.actor.optim.lr=1e-6 \
  reward_model.overlong_buffer.enable=True reward_model.overlong_buffer.len=4096 reward_model.overlong_buffer.penalty_factor=1.0 \
  trainer.total_epochs=1 trainer.test_freq=5 trainer.save_freq=5 "$@"
