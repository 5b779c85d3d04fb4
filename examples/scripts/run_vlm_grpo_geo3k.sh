#!/bin/bash
# GRPO on a vision-language model (Qwen2.5-VL-7B, Geometry3K) — counterpart of the reference stack's
# 3rdparty/verl/examples/grpo_trainer/run_qwen2_5_vl-7b.sh.  MODEL is an HF checkpoint directory (config.json + safetensors +
# tokenizer); DATA / VAL come from examples/preprocess_geo3k.py.  Colocated by default; NROLL>0 dedicates rollout GPUs.
set -e
NGPU=${NGPU:-8}
NROLL=${NROLL:-0}
MODEL=${MODEL:?path to a Qwen2-VL / Qwen2.5-VL checkpoint directory}
DATA=${DATA:?geo3k train parquet}
VAL=${VAL:-$DATA}
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NGPU --master-addr 127.0.0.1 --master-port 29500 \
  -m polyrl_b200.trainer.main_stream \
  algorithm.adv_estimator=grpo data.train_files=$DATA data.val_files=$VAL data.train_batch_size=512 \
  data.max_prompt_length=1024 data.max_response_length=2048 data.filter_overlong_prompts=True data.image_key=images \
  data.image_max_pixels=1003520 \
  actor_rollout_ref.model.path=$MODEL actor_rollout_ref.rollout.n=5 actor_rollout_ref.rollout.min_stream_batch_size=80 \
  actor_rollout_ref.actor.optim.lr=1e-6 actor_rollout_ref.actor.ppo_mini_batch_size=128 actor_rollout_ref.actor.use_dynamic_bsz=True \
  actor_rollout_ref.actor.use_kl_loss=True actor_rollout_ref.actor.kl_loss_coef=0.01 actor_rollout_ref.actor.kl_loss_type=low_var_kl \
  actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True trainer.test_freq=5 trainer.save_freq=20 \
  trainer.n_gpus_per_node=$NGPU trainer.n_rollout_gpus=$NROLL "$@"
