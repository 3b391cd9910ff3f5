#!/bin/bash
# Reduced round-end evidence (no --set full captures): gpu suite in one process, smoke,
# bench + reference arm, ncu launch list.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
SEGB200_WGRAD_STREAM=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/prof_ncu1.log 2>&1
python tools/launch_list.py gpurun_out/launches.csv -v > gpurun_out/launch_summary.txt; head -2 gpurun_out/launch_summary.txt
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
