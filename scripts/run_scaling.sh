#!/bin/bash
# Reproduce the headline numbers on one 8xB200 box: product arm and stock-library baseline arm at 1/2/4/8 GPUs.
#   scripts/run_scaling.sh [steps] [warmup]        -> gpurun_out/scale_<impl>_n<N>.json
K=${1:-2}; W=${2:-3}
mkdir -p gpurun_out
for impl in own baseline; do
  for n in 1 2 4 8; do
    out=gpurun_out/scale_${impl}_n${n}.json
    if [ "$n" == "1" ]; then
      python bench.py --gpus 1 --steps $K --warmup $W --impl $impl > $out 2> ${out%.json}.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
        bench.py --gpus $n --steps $K --warmup $W --impl $impl > $out 2> ${out%.json}.err
    fi
    python - "$out" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{d.get('impl','own'):9s} N={d['n_gpus']} {d['value']:8.2f} {d['unit']}  {d['ms_per_step']:.0f} ms/step  e2e {d['e2e']['value']:.2f}")
except Exception as e:
    print("failed:", sys.argv[1], e)
PY
  done
done
