#!/bin/bash
# ncu --set full capture of one conv kernel instantiation while tools/layer_prof.py runs.
#   tools/ncu_layer.sh <out-name> <layer> '<kernel regex on the demangled name>' [skip]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=$1; LAYER=$2; RX=$3; SKIP=${4:-0}
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"$RX" -s "$SKIP" -c 1 -f -o "gpurun_out/$OUT" python tools/layer_prof.py "$LAYER" \
  > "gpurun_out/$OUT.log" 2>&1
echo "ncu $OUT exit=$?"
