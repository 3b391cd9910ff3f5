#!/bin/bash
# data-parallel bench at N GPUs: tools/gpu_dp.sh N [extra env assignments...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; shift
TAG=$(echo "$*" | tr ' =' '__')
env "$@" SEGB200_BENCH_WATCHDOG=200 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
  --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-60} --warmup ${WARMUP:-10} --skip-cpu \
  > gpurun_out/dp${N}_$TAG.json 2> gpurun_out/dp${N}_$TAG.err
echo "N=$N $* exit=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/dp${N}_$TAG.json')); print('ms/step %.3f value %.0f e2e %.0f (%.3f ms) clocks %s'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['clocks']))" 2>&1 | tail -1)"
