#!/bin/bash
# multi-GPU evidence: 2-rank equivalence test, bench at N ranks, concurrent timeline at N ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
[ "$SKIP_TEST" = 1 ] || timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -x -q 2>&1 | tail -3
P=$((29500 + RANDOM % 1000))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 200 --warmup 10 --skip-cpu > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err; echo "bench exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_dp$N.json')); print('N=%d ms/step %.4f value %.0f e2e %.0f (%.4f ms)'%(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
P=$((29500 + RANDOM % 1000))
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/graph_timeline.py > gpurun_out/graph_timeline_dp.log 2>&1; echo "timeline exit=$?"
head -3 gpurun_out/graph_timeline_w${N}_r0.txt
