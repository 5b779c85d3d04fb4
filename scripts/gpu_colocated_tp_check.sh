#!/bin/bash
# 2-GPU colocated GRPO with a TP-2 rollout engine (ZeRO over both ranks for training, one TP-2 replica for rollout).
set -o pipefail
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 \
  -m polyrl_b200.trainer.main_stream \
  algorithm.adv_estimator=grpo data.synthetic=True data.synthetic_num_samples=256 data.train_batch_size=16 \
  data.max_prompt_length=128 data.max_response_length=128 \
  actor_rollout_ref.model.path=${MODEL:-qwen2.5-0.5b} actor_rollout_ref.actor.optim.lr=1e-6 \
  actor_rollout_ref.actor.ppo_mini_batch_size=16 actor_rollout_ref.actor.use_dynamic_bsz=True \
  actor_rollout_ref.actor.ppo_max_token_len_per_gpu=8192 actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True \
  actor_rollout_ref.rollout.n=4 actor_rollout_ref.rollout.tensor_model_parallel_size=2 actor_rollout_ref.rollout.max_num_seqs=64 \
  trainer.n_gpus_per_node=2 trainer.n_rollout_gpus=0 trainer.logger=[console] \
  trainer.total_training_steps=3 trainer.total_epochs=1 trainer.save_freq=-1 trainer.test_freq=-1 "$@" 2>&1
