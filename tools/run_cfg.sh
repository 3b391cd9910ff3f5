cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_first_layer.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -6
