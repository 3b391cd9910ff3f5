#!/bin/bash
# N-GPU A/B of one environment switch: tools/r2_dp_ab.sh N VAR v0 v1 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; VAR=$2; shift; shift
for c in "$@"; do
P=$((29500 + RANDOM % 1000))
env $VAR=$c timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 200 --warmup 10 --skip-cpu > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_dp$N.json')); print('$VAR=$c N=%d ms/step %.4f value %.0f e2e %.0f (%.4f ms)'%(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))" 2>&1 | tail -1
done
