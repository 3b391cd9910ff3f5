#!/bin/bash
# 8-GPU data-parallel A/B: SMs reserved for the all-reduce kernels during the backward pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
run() {  # label, env...
  local label=$1; shift
  P=$((29500 + RANDOM % 1000))
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 200 --warmup 10 --skip-cpu > gpurun_out/bench_dp${N}_$label.json 2> gpurun_out/bench_dp${N}_$label.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_dp${N}_$label.json')); print('$label N=%d ms/step %.4f value %.0f e2e %.0f (%.4f ms)'%(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))" 2>&1 | tail -1
}
run reserve0 SEGB200_DP_SM_RESERVE=0
run reserve24 SEGB200_DP_SM_RESERVE=24
run reserve12 SEGB200_DP_SM_RESERVE=12
