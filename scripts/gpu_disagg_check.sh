#!/bin/bash
# 2-GPU disaggregated streaming run on real GPUs: 1 rollout GPU (engine server behind the C++ manager) + 1 trainer GPU,
# peer-memory weight push every step.  usage: scripts/gpu_disagg_check.sh [extra overrides]
set -o pipefail
NGPU=${NGPU:-2}; NROLL=${NROLL:-1}
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NGPU --master-addr 127.0.0.1 --master-port 29511 \
  -m polyrl_b200.trainer.main_stream \
  algorithm.adv_estimator=grpo data.synthetic=True data.synthetic_num_samples=256 data.train_batch_size=16 \
  data.max_prompt_length=128 data.max_response_length=128 \
  actor_rollout_ref.model.path=${MODEL:-qwen2.5-0.5b} actor_rollout_ref.actor.optim.lr=1e-6 \
  actor_rollout_ref.actor.ppo_mini_batch_size=16 actor_rollout_ref.actor.use_dynamic_bsz=True \
  actor_rollout_ref.actor.ppo_max_token_len_per_gpu=8192 actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True \
  actor_rollout_ref.rollout.name=native-disaggregated actor_rollout_ref.rollout.n=4 \
  actor_rollout_ref.rollout.min_stream_batch_size=16 actor_rollout_ref.rollout.max_num_seqs=64 \
  actor_rollout_ref.rollout.rollout_manager.port=0 actor_rollout_ref.rollout.rollout_manager.max_local_gen_s=60 \
  trainer.n_gpus_per_node=$NGPU trainer.n_rollout_gpus=$NROLL trainer.logger=[console] \
  trainer.total_training_steps=3 trainer.total_epochs=1 trainer.save_freq=-1 trainer.test_freq=-1 "$@" 2>&1
