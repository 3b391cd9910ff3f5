cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pointwise.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q -k "classmap_tail or deconv or config5 or config4" 2>&1 | tail -8
timeout 300 python tools/configs_check.py 4 2>&1 | grep config
SEGB200_TAIL_MMA=0 timeout 300 python tools/configs_check.py 4 2>&1 | grep config
timeout 200 python tools/deconv_timeline.py 2>&1 | grep -i "total\|tail"
grep "classmap_tail" gpurun_out/diag.jsonl | tail -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:classmap_tail -c 1 -f -o gpurun_out/r02_tail2 python tools/deconv_timeline.py > gpurun_out/ncu_tail2.log 2>&1; echo "ncu tail exit=$?"
