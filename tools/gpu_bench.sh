#!/bin/bash
# bench + smoke on the GPU box; logs -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
