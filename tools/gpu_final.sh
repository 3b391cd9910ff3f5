#!/bin/bash
# Round-end evidence: full gpu test suite, smoke, bench (+reference arm), ncu launch list and
# --set full captures of the heaviest kernels.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_check.sh > gpurun_out/check.log 2>&1; cat gpurun_out/summary.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
CMD="python bench.py --steps 3 --warmup 3 --skip-cpu"
SEGB200_WGRAD_STREAM=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
python tools/launch_list.py gpurun_out/launches.csv -v > gpurun_out/launch_summary.txt; head -2 gpurun_out/launch_summary.txt
bash tools/ncu_bench.sh r01_twgrad16_conv1_1 "twgrad_kernel<.int.16, .int.32>" 0
bash tools/ncu_bench.sh r01_tconv_conv2_2_fwd "tconv_kernel<.int.64, .int.64, .bool.1, .int.2>" 0
bash tools/ncu_bench.sh r01_pool1_bwd "maxpool_bwd_row8_kernel" 3
bash tools/ncu_bench.sh r01_pool1_fwd "maxpool_fwd_row8_kernel" 0
