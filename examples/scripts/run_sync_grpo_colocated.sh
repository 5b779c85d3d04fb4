#!/bin/bash
# Synchronous colocated GRPO (every GPU rolls out, then trains) — the "sync baseline" of the reference
# (examples/scripts + rlboost/verl_stream/trainer/main_ppo.py).
set -e
NGPU=${NGPU:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NGPU --master-addr 127.0.0.1 --master-port 29500 \
  -m polyrl_b200.trainer.main_stream \
  algorithm.adv_estimator=grpo data.synthetic=True data.train_batch_size=128 data.max_prompt_length=256 data.max_response_length=1024 \
  actor_rollout_ref.model.path=llama3-8b actor_rollout_ref.rollout.n=8 actor_rollout_ref.rollout.min_stream_batch_size=128 \
  actor_rollout_ref.actor.ppo_mini_batch_size=128 actor_rollout_ref.actor.use_dynamic_bsz=True \
  actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True trainer.n_gpus_per_node=$NGPU trainer.n_rollout_gpus=0 "$@"
