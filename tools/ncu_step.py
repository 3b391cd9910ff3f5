"""One eager, single-stream U-Net train step (bench.py's workload) for ncu, with a marker
kernel in front of every C-ABI call so that tools/ncu_traffic.py can attribute the profiled
launches to (layer, kind):

  python tools/ncu_step.py > gpurun_out/ncu_step_plain.log 2>&1 &&
  ncu --profile-from-start off --clock-control none \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --csv --log-file gpurun_out/step_ncu.csv python tools/ncu_step.py
  python tools/ncu_traffic.py gpurun_out/step_ncu.csv gpurun_out/step_calls.json

Writes gpurun_out/step_calls.json: [[fn, tag, family, flops, bytes], ...] in launch order."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('SEGB200_WGRAD_STREAM', '0')
os.environ.setdefault('SEGB200_NO_GRAPH', '1')

import torch  # noqa: E402

import bench  # noqa: E402
from segmentation_b200 import native as N  # noqa: E402
from segmentation_b200.models.unet import UNetModel  # noqa: E402


def main():
    torch.cuda.set_device(0)
    ds = bench.SyntheticDataSet(bench.BATCH, seed=1000, pool=2, pinned=False)
    model = UNetModel(dataset=ds, n_classes=bench.NCLS, input_dims=bench.S, n_kernels=bench.NK,
                      learning_rate=1e-4, load_snapshot=False, save_dir=None, seed=0)
    ex = model._get_exec(bench.BATCH, True)
    for _ in range(3):
        model.train_step()
    torch.cuda.synchronize()
    ex.stage(*[t.cuda() for t in ds.pool[0]])
    orig = N.call

    def marked(name, *args):
        torch.cuda._sleep(1)                     # marker launch: delimits the calls in the ncu list
        orig(name, *args)

    N.call = marked
    N.TIMELINE = []
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    ex.forward()
    ex.loss(True)
    ex.backward()
    for grp in model.opt_groups:
        N.set_tag('adam')
        model.store.adam_launch(0.0, chunk_range=grp['chunks'])
    torch.cuda._sleep(1)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    calls = [[n, tag, fam, fl, by] for (n, tag, a, b, fam, fl, by) in N.TIMELINE]
    N.TIMELINE = None
    N.call = orig
    out = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, 'step_calls.json'), 'w') as f:
        json.dump(calls, f)
    print('calls', len(calls))


if __name__ == '__main__':
    main()
