#!/bin/bash
# ncu evidence: launch list of a bench run + one --set full capture of the conv kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --skip-cpu"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
echo "launch list exit=$?"
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-igemm_kernel}" -s ${KSKIP:-0} -c ${KCOUNT:-8} \
    -o gpurun_out/prof_full $CMD > gpurun_out/prof_ncu2.log 2>&1
echo "full capture exit=$?"
ls -la gpurun_out/ | tail -12
