"""Per-launch timeline of DeconvModel 1024x1024 bs32 inference (BASELINE config 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from segmentation_b200 import native as N
from segmentation_b200.models.deconvolution import DeconvModel
B, S = int(os.environ.get('B', 32)), int(os.environ.get('S', 1024))
m = DeconvModel(None, dataset=None, n_classes=2, input_dims=S, n_kernels=32, mode='INFERENCE',
                load_snapshot=False, save_dir=None)
x = torch.rand(B, S, S, 3, device='cuda')
ex = m._get_exec(B, False)
ex.infer(x); torch.cuda.synchronize()
for _ in range(2):
    torch.cuda.synchronize()
    torch.cuda._sleep(int(4e-3 * 1.9e9))
    N.TIMELINE = []
    ex.infer(x)
    torch.cuda.synchronize()
    tl = [(n, tag, a.elapsed_time(b)) for (n, tag, a, b, *_) in N.TIMELINE]
    N.TIMELINE = None
tot = sum(t for _, _, t in tl)
print('total %.2f ms' % tot)
for n, tag, t in tl:
    print('%-28s %-12s %8.3f ms %5.1f%%' % (n, tag, t, 100 * t / tot))
