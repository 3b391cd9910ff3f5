#!/bin/bash
# quick evidence pass: full gpu test suite, bench (no CPU leg), concurrent graph timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/diag.jsonl
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 200 --warmup 10 --skip-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('ms/step %.4f value %.0f e2e %.0f (%.4f ms)'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step']))
print('loss first %.4f last %.4f'%(d['loss_first'],d['loss_last']))
r=d['roofline']
for f in r['families']: print('  %-10s n=%2d %7.1f us %5.1f%% %7.1f TF/s %7.0f GB/s %s %.3f'%(f['family'],f['launches'],f['us'],100*f['share'],f['tflops'],f['gbs'],f['bound'],f['frac']))
PY
timeout 300 python tools/graph_timeline.py > gpurun_out/graph_timeline.log 2>&1; head -3 gpurun_out/graph_timeline.txt
