cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pointwise.py tests/test_gpu_conv.py -m gpu -x -q -k "maxpool_bn or folded_batchnorm or deconv_fwd_bwd or conv_fwd" 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q -k "deconv or config5 or config4" 2>&1 | tail -6
timeout 300 python tools/configs_check.py 4 5d 2>&1 | grep config
SEGB200_FUSE_BN=0 timeout 300 python tools/configs_check.py 4 2>&1 | grep config
timeout 200 python tools/deconv_timeline.py 2>&1 | tail -30
grep "fused_bn\|conv_affine" gpurun_out/diag.jsonl | tail -8
