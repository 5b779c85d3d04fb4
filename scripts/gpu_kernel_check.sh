#!/bin/bash
# Run every GPU kernel test in its own process (a device trap in one test must not poison the rest).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_info.txt 2>&1
python - <<'PY' > gpurun_out/test_ids.txt
import subprocess, sys
r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_kernels_gpu.py", "--collect-only", "-q"], capture_output=True, text=True)
for l in r.stdout.splitlines():
    if "::" in l: print(l.strip())
PY
: > gpurun_out/kernel_tests.log
while read -r tid; do
  echo "=== $tid" >> gpurun_out/kernel_tests.log
  timeout 300 python -m pytest "$tid" -x -q 2>&1 | tail -25 >> gpurun_out/kernel_tests.log
  echo "rc=$?" >> gpurun_out/kernel_tests.log
done < gpurun_out/test_ids.txt
grep -E "^(===|rc=)|passed|failed|error" gpurun_out/kernel_tests.log | paste - - - | sed 's/  */ /g' | tail -60
