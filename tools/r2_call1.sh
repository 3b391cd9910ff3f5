#!/bin/bash
# round 2, GPU call 1: TRED experimental test + A/B, baseline bench, fresh launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
SEGB200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -k tensor_reduce -x -q 2>&1 | tail -5
for v in 0 1 0 1; do
  SEGB200_TWGRAD_TRED=$v timeout 300 python bench.py --steps 200 --warmup 10 --skip-cpu \
    > gpurun_out/ab_tred_$v.json 2> gpurun_out/ab_tred_$v.err
  echo "TRED=$v exit=$? $(python -c "import json; d=json.load(open('gpurun_out/ab_tred_$v.json')); print('ms/step %.4f value %.0f e2e %.0f loss %s'%(d['ms_per_step'],d['value'],d['e2e']['value'],d.get('loss')))" 2>&1 | tail -1)"
done
cp gpurun_out/timeline.json gpurun_out/timeline_tred1.json 2>/dev/null
SEGB200_WGRAD_STREAM=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches_r2a.csv python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/prof_ncu1.log 2>&1
python tools/launch_list.py gpurun_out/launches_r2a.csv -v > gpurun_out/launch_summary_r2a.txt; head -40 gpurun_out/launch_summary_r2a.txt
