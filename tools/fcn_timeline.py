"""Per-launch timeline of an FCN-8s 512x512 bs16 train step (BASELINE config 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from segmentation_b200 import native as N
from segmentation_b200.models.fcn import FCNModel
from tools.configs_check import DS
ds = DS(16, 512, 21)
m = FCNModel(None, dataset=ds, n_classes=21, fcn_type='8s', input_dims=512, n_kernels=32,
             learning_rate=1e-4, load_snapshot=False, save_dir=None)
ex = m._get_exec(16, True)
ex.use_graph = False
ex.stage(ds.x.cuda(), ds.y.cuda())
for _ in range(2):
    torch.cuda.synchronize()
    torch.cuda._sleep(int(6e-3 * 1.9e9))
    N.TIMELINE = []
    ex.forward(); ex.loss(True); ex.backward()
    torch.cuda.synchronize()
    tl = [(n, tag, a.elapsed_time(b)) for (n, tag, a, b, *_) in N.TIMELINE]
    N.TIMELINE = None
    m.store.grad.zero_()
tot = sum(t for _, _, t in tl)
print('total %.2f ms' % tot)
for n, tag, t in tl:
    print('%-28s %-22s %8.3f ms %5.1f%%' % (n, tag, t, 100 * t / tot))
