cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
bash tools/ab.sh SEGB200_DYN_TILES 0 1 2
