cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/configs_check.py 2 4 5u 5d 2>&1 | tail -6
timeout 300 python tools/deconv_timeline.py 2>&1 | tail -45
