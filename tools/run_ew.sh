cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_fullsize.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 300 --warmup 20 --skip-cpu > gpurun_out/ab_ew.json 2> gpurun_out/ab_ew.err; python -c "import json; d=json.load(open('gpurun_out/ab_ew.json')); print('ms/step %.4f e2e %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'])); r=d['roofline']; print(r['kernel'], r['frac'])"
timeout 300 python tools/configs_check.py 2 4 2>&1 | grep config
