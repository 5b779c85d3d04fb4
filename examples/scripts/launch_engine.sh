#!/bin/bash
# Launch a stand-alone native rollout engine that self-registers with a rollout manager as an elastic REMOTE instance
# (the reference's examples/scripts/launch_sglang.sh + rlboost.sglang.launch_server).
python -m polyrl_b200.rollout.launch_server --model-path ${MODEL:-llama3-8b} --host ${HOST:-127.0.0.1} --port ${PORT:-30000} \
  --rollout-manager-address ${MANAGER:-127.0.0.1:5000} --max-num-seqs 256 --max-model-len 2048 "$@"
