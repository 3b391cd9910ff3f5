#!/bin/bash
# A/B of one environment switch: tools/gpu_ab3.sh VAR v1 v2 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu \
    > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.err
  echo "$VAR=$v exit=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_${VAR}_$v.json')); print('ms/step %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))" 2>&1 | tail -1)"
  tail -2 gpurun_out/ab_${VAR}_$v.err
done
