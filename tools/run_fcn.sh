cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pointwise.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q -k "maxpool or fcn" 2>&1 | tail -4
timeout 300 python tools/configs_check.py 2 2>&1 | grep config
