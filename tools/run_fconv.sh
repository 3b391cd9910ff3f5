cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_first_layer.py -m gpu -x -q 2>&1 | tail -15
