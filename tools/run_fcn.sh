cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pointwise.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q -k "upscore8 or fcn" 2>&1 | tail -4
timeout 300 python tools/configs_check.py 2 2>&1 | grep config
cd tools; timeout 300 python fcn_timeline.py 2>&1 | tail -150 > ../gpurun_out/fcn_timeline.txt; head -3 ../gpurun_out/fcn_timeline.txt; grep "upscore8" ../gpurun_out/fcn_timeline.txt
