#!/bin/bash
# FCN-8s config 2: parity tests of the FCN paths, step time, per-launch timeline
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pointwise.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q -k "upscore8 or maxpool or fcn" 2>&1 | tail -4
timeout 300 python tools/configs_check.py 2 2>&1 | grep config
cd tools; timeout 300 python fcn_timeline.py > ../gpurun_out/fcn_timeline.txt 2>&1; head -14 ../gpurun_out/fcn_timeline.txt
