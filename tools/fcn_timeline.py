"""Per-launch timeline (CUDA events, eager, one stream) of one FCN-8s 512x512 / 21 classes /
batch 16 train step (BASELINE config 2): which launches the 1.79 ms are made of."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('SEGB200_WGRAD_STREAM', '0')
os.environ.setdefault('SEGB200_NO_GRAPH', '1')
import numpy as np, torch
from segmentation_b200 import native as N
from segmentation_b200.models.fcn import FCNModel
from configs_check import DS

ds = DS(16, 512, 21)
m = FCNModel(None, dataset=ds, n_classes=21, fcn_type='8s', input_dims=512, n_kernels=32,
             learning_rate=1e-4, load_snapshot=False, save_dir=None)
ex = m._get_exec(16, True)
for _ in range(3):
    m.train_step()
torch.cuda.synchronize()
x, y = ds.next_batch()
ex.stage(x.cuda(), y.cuda())
for rep in range(2):
    torch.cuda.synchronize()
    N.TIMELINE = []
    ex.forward_for_step(); ex.loss(True); ex.backward()
    for grp in m.opt_groups:
        N.set_tag('adam')
        m.store.adam_launch(0.0, chunk_range=grp['chunks'])
    torch.cuda.synchronize()
    tl = [(n, tag, a.elapsed_time(b), fam, fl, by) for (n, tag, a, b, fam, fl, by) in N.TIMELINE]
    N.TIMELINE = None
tot = sum(t[2] for t in tl)
print('serial total %.3f ms, %d calls' % (tot, len(tl)))
fam = {}
for n, tag, t, f, fl, by in tl:
    fam.setdefault(f, [0, 0.0]); fam[f][0] += 1; fam[f][1] += t
for f, (c, t) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print('  family %-10s n=%3d %8.3f ms %5.1f%%' % (f, c, t, 100 * t / tot))
for n, tag, t, f, fl, by in tl:
    print('%-26s %-14s %-8s %8.1f us %5.1f%%  %7.1f TF/s %7.0f GB/s' %
          (n.replace('seg_', ''), tag, f, t * 1e3, 100 * t / tot, fl / t / 1e9 if t else 0,
           by / t / 1e6 if t else 0))
