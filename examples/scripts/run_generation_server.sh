#!/usr/bin/env bash
# Offline generation through stand-alone rollout servers (the reference's verl/trainer/main_generation_server.py and
# examples/generation/run_deepseek_v2_lite_math.sh): one engine server per GPU speaking the OpenAI chat-completions API,
# n samples per prompt, parquet in / parquet out.  MODEL may be a preset (deepseek-v2-lite, moonlight-16b-a3b, llama3-8b, ...)
# with --load-format dummy, or an HF checkpoint directory.
set -euo pipefail
MODEL=${MODEL:-$HOME/models/DeepSeek-V2-Lite-Chat}
DATA=${DATA:-$HOME/data/gsm8k/test.parquet}
OUT=${OUT:-$HOME/data/gsm8k/deepseek_v2_lite_gen_test.parquet}
python -m polyrl_b200.trainer.main_generation_server \
    --model "$MODEL" --data "$DATA" --output "$OUT" --prompt-key prompt \
    --n-samples "${N_SAMPLES:-1}" --temperature "${TEMPERATURE:-1.0}" --top-p "${TOP_P:-0.7}" \
    --prompt-length 2048 --response-length 1024 --num-replicas "${NUM_REPLICAS:-8}" --tokenizer "${TOKENIZER:-$MODEL}"
# against servers that already run (e.g. launched by examples/scripts/launch_engine.sh with --tokenizer-path):
#   python -m polyrl_b200.trainer.main_generation_server --model m --data d.parquet --output o.parquet --server host:30000 --server host:30001
