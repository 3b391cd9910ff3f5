#!/bin/bash
# Runs the -m gpu suite in separate processes (a trapped kernel poisons its CUDA
# context, so one failing group must not hide the others).  Logs -> gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/diag.jsonl
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
  local name=$1; shift; local to=$1; shift
  timeout "$to" python -m pytest -m gpu -q --no-header -p no:cacheprovider "$@" > "gpurun_out/$name.log" 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 4 "gpurun_out/$name.log"
}
: > gpurun_out/summary.txt
run pointwise 600 tests/test_gpu_pointwise.py
run conv_simt 600 tests/test_gpu_conv.py -k "simt or deconv or crop"
run probe0 300 tests/test_gpu_conv.py -k "probe and 0"
run probe1 300 tests/test_gpu_conv.py -k "probe and 1"
run conv_umma_fwd 600 tests/test_gpu_conv.py -k "umma and fwd and not im2col"
run conv_umma_bwd 600 tests/test_gpu_conv.py -k "umma and bwd and not im2col"
run conv_im2col 600 tests/test_gpu_conv.py -k "im2col"
run unet_simt 900 tests/test_gpu_unet.py -k "simt"
run unet_umma 900 tests/test_gpu_unet.py -k "umma or train or infer"
run unet_full 900 tests/test_gpu_unet.py -k "config1"
run models 1200 tests/test_gpu_models.py
cat gpurun_out/summary.txt
