#!/bin/bash
# ncu --set full capture of ONE launch of a kernel instantiation inside the eager first
# train step of bench.py.   tools/ncu_bench.sh <out-name> '<regex on demangled name>' <skip>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=$1; RX=$2; SKIP=${3:-0}
SEGB200_WGRAD_STREAM=0 timeout 600 ncu --set full --clock-control none --import-source on \
  --kernel-name-base demangled -k regex:"$RX" -s "$SKIP" -c 1 -f -o "gpurun_out/$OUT" \
  python bench.py --steps 1 --warmup 3 --skip-cpu > "gpurun_out/$OUT.log" 2>&1
echo "ncu $OUT exit=$?"
