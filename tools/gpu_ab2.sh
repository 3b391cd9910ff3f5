#!/bin/bash
# A/B of the weight-gradient cluster size.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cs in ${CS_LIST:-8 4 1}; do
  SEGB200_WGRAD_CLUSTER=$cs timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu \
    > gpurun_out/ab_cs$cs.json 2> gpurun_out/ab_cs$cs.err
  echo "cluster=$cs exit=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_cs$cs.json')); print('ms/step %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))" 2>&1 | tail -1)"
  tail -2 gpurun_out/ab_cs$cs.err
done
