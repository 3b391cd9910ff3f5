#!/bin/bash
# full evidence pass: gpu test suite, smoke, bench (+reference arm), marked ncu step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/diag.jsonl
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
python tools/ncu_step.py > gpurun_out/ncu_step_plain.log 2>&1 && \
timeout 600 ncu --profile-from-start off --clock-control none \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --csv --log-file gpurun_out/step_ncu.csv python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/ncu_step.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('ms/step %.4f value %.0f e2e %.0f'%(d['ms_per_step'],d['value'],d['e2e']['value']))
print('loss first %.4f last %.4f'%(d['loss_first'],d['loss_last']), 'parity', d['parity_check'])
r=d['roofline']
print('roofline', r['kernel'], r['bound'], '%.3f'%r['frac'], 'serial', r['serial_step_ms'])
for f in r['families']: print('  %-10s n=%2d %7.1f us %5.1f%% %7.1f TF/s %7.0f GB/s %s %.3f'%(f['family'],f['launches'],f['us'],100*f['share'],f['tflops'],f['gbs'],f['bound'],f['frac']))
PY
python tools/ncu_traffic.py gpurun_out/step_ncu.csv gpurun_out/step_calls.json > gpurun_out/ncu_traffic.log 2>&1; echo "traffic exit=$?"; tail -2 gpurun_out/ncu_traffic.log
cp profiles/r02_traffic.json profiles/r02_launches.md gpurun_out/ 2>/dev/null
timeout 300 python tools/configs_check.py 2 4 5u 5d 2>&1 | grep config
