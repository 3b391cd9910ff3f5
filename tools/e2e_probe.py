"""Where does the end-to-end (host batch -> loss on host) step time go?  Measures pinned
H2D bandwidth, the sync/readback latency and the graph replay alone on this box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)


def timed(fn, n=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for mb in (1, 4, 13, 64):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    ms = timed(lambda: d.copy_(h, non_blocking=True))
    print('H2D pinned %3d MiB: %.3f ms  %.1f GB/s' % (mb, ms, (mb << 20) / ms / 1e6))
    ms = timed(lambda: h.copy_(d, non_blocking=True))
    print('D2H pinned %3d MiB: %.3f ms  %.1f GB/s' % (mb, ms, (mb << 20) / ms / 1e6))
s = torch.zeros(1, device=dev)
print('item() latency: %.3f ms' % timed(lambda: s.item()))
s.fill_(1.0)
print('fill_+item(): %.3f ms' % timed(lambda: (s.fill_(2.0), s.item())))

import bench
from segmentation_b200.models.unet import UNetModel
ds = bench.SyntheticDataSet(16, seed=1000)
model = UNetModel(dataset=ds, n_classes=2, input_dims=256, n_kernels=32, learning_rate=1e-4,
                  load_snapshot=False, save_dir=None, seed=0)
for _ in range(4):
    model.train_step()
torch.cuda.synchronize()
ex = model._get_exec(16, True)
print('graph replay only: %.3f ms' % timed(lambda: ex.graph.replay()))
x, y = ds.pool[0]
print('stage only: %.3f ms' % timed(lambda: ex.stage(x, y)))
print('train_step(batch) no readback: %.3f ms' % timed(lambda: model.train_step(ds.next_batch())))
print('train_step(batch) + loss: %.3f ms' % timed(lambda: (model.train_step(ds.next_batch()), model.seg_loss_op)))
print('train_step() + loss: %.3f ms' % timed(lambda: (model.train_step(), model.seg_loss_op)))

# distribution of the end-to-end step time: 6 x 30 steps, per-step host wall times
import statistics
for rep in range(6):
    ts = []
    torch.cuda.synchronize()
    for i in range(30):
        t0 = time.perf_counter()
        model.train_step(); _ = model.seg_loss_op
        ts.append((time.perf_counter() - t0) * 1e3)
    print('rep %d: mean %.3f median %.3f max %.3f ms  slow steps(>2ms): %d' %
          (rep, statistics.mean(ts), statistics.median(ts), max(ts), sum(t > 2 for t in ts)))
# same with the NVML sampler thread of bench.py running
cs = bench.ClockSampler(0)
for rep in range(3):
    ts = []
    for i in range(30):
        t0 = time.perf_counter()
        model.train_step(); _ = model.seg_loss_op
        ts.append((time.perf_counter() - t0) * 1e3)
    print('with sampler rep %d: mean %.3f median %.3f max %.3f ms  slow steps(>2ms): %d' %
          (rep, statistics.mean(ts), statistics.median(ts), max(ts), sum(t > 2 for t in ts)))
print(cs.stop())
