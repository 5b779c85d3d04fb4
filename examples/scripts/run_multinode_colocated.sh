#!/bin/bash
# Colocated GRPO on several 8xB200 nodes.  Run on EVERY node with the same MASTER_ADDR; NODE_RANK = 0..NNODES-1.
# ZeRO shards stay inside a node (NVLink peer-memory kernels), replicas across nodes all-reduce over NCCL
# (workers/hybrid_worker.py::auto_fsdp_size); every GPU hosts its own rollout engine replica.
set -e
NNODES=${NNODES:-2}; NODE_RANK=${NODE_RANK:-0}; MASTER_ADDR=${MASTER_ADDR:?set MASTER_ADDR}; MASTER_PORT=${MASTER_PORT:-29500}
python -m torch.distributed.run --nnodes=$NNODES --node-rank=$NODE_RANK --nproc-per-node=${NGPU:-8} \
  --master-addr $MASTER_ADDR --master-port $MASTER_PORT -m polyrl_b200.trainer.main_ppo \
  trainer.nnodes=$NNODES trainer.n_gpus_per_node=${NGPU:-8} actor_rollout_ref.model.path=${MODEL:-llama3-8b} \
  algorithm.adv_estimator=grpo actor_rollout_ref.rollout.n=8 data.train_batch_size=$((16 * NNODES * ${NGPU:-8})) \
  actor_rollout_ref.actor.ppo_mini_batch_size=$((16 * NNODES * ${NGPU:-8})) actor_rollout_ref.actor.use_dynamic_bsz=True \
  actor_rollout_ref.actor.ppo_max_token_len_per_gpu=16384 data.max_response_length=512 "$@"
