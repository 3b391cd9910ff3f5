#!/bin/bash
# A/B of an environment switch: tools/ab.sh VAR v0 v1 [repeats]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VAR=$1; A=$2; B=$3; R=${4:-2}
for i in $(seq $R); do for v in $A $B; do
  env $VAR=$v timeout 300 python bench.py --steps 300 --warmup 20 --skip-cpu > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$VAR=$v ms/step %.4f e2e %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step']))"
done; done
