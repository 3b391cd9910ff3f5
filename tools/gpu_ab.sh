#!/bin/bash
# A/B of the scheduling switches: PDL and the weight-gradient side stream.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  SEGB200_PDL=$1 SEGB200_WGRAD_STREAM=$2 timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu \
    > gpurun_out/ab_pdl$1_side$2.json 2> gpurun_out/ab_pdl$1_side$2.err
  echo "pdl=$1 side=$2 exit=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/ab_pdl$1_side$2.json')); print('ms/step %.3f value %.0f e2e %.0f (%.3f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))" 2>&1 | tail -1)"
done
