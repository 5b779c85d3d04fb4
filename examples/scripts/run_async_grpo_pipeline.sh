#!/bin/bash
# Streaming (async) GRPO, Qwen3-1.7B-shaped model, disaggregated: 4 trainer GPUs (ZeRO-3, peer-memory comm) +
# 4 rollout GPUs behind the C++ rollout manager.  Mirrors the reference's examples/scripts/run_async_grpo_pipeline.sh
# (same knobs; rollout.name=sglang-disaggregated is accepted as an alias of native-disaggregated).
set -e
NGPU=${NGPU:-8}
NROLL=${NROLL:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NGPU --master-addr 127.0.0.1 --master-port 29500 \
  -m polyrl_b200.trainer.main_stream \
  algorithm.adv_estimator=grpo \
  data.synthetic=True data.train_batch_size=128 data.max_prompt_length=512 data.max_response_length=14336 \
  actor_rollout_ref.model.path=qwen3-1.7b \
  actor_rollout_ref.actor.optim.lr=1e-6 \
  actor_rollout_ref.actor.ppo_mini_batch_size=128 \
  actor_rollout_ref.actor.use_dynamic_bsz=True \
  actor_rollout_ref.actor.ppo_max_token_len_per_gpu=16384 \
  actor_rollout_ref.actor.use_kl_loss=True actor_rollout_ref.actor.kl_loss_coef=0.001 actor_rollout_ref.actor.kl_loss_type=low_var_kl \
  actor_rollout_ref.actor.fsdp_config.comm_backend=p2p \
  actor_rollout_ref.rollout.name=sglang-disaggregated \
  actor_rollout_ref.rollout.n=8 \
  actor_rollout_ref.rollout.min_stream_batch_size=16 \
  actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True \
  actor_rollout_ref.rollout.max_num_seqs=256 \
  actor_rollout_ref.rollout.rollout_manager.port=5000 \
  trainer.n_gpus_per_node=$NGPU trainer.n_rollout_gpus=$NROLL trainer.local_rollout=True \
  trainer.logger=[console,file] trainer.project_name=polyrl_b200 trainer.experiment_name=qwen3_1p7b_async_grpo \
  trainer.save_freq=20 trainer.test_freq=10 trainer.total_epochs=1 "$@"
