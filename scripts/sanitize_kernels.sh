#!/bin/bash
# compute-sanitizer sweep of the single-GPU kernel tests (memcheck, synccheck, racecheck) — SURVEY 5.2 asks for sanitizer
# targets the reference never had.  Output: gpurun_out/sanitize_<tool>.log + a one-line summary per tool.
#   usage: scripts/sanitize_kernels.sh [pytest -k expression]
mkdir -p gpurun_out
K=${1:-"gemm_bf16 or swiglu or linear_ce or rmsnorm or rope or adamw or sample or paged_decode or quant or skinny"}
for tool in memcheck synccheck racecheck; do
  log=gpurun_out/sanitize_${tool}.log
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "$K" > $log 2>&1
  rc=$?
  echo "SANITIZE tool=$tool rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | $(grep -E 'passed|failed' $log | tail -1)"
done
