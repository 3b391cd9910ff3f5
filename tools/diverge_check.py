"""Two identically initialised U-Nets stepped on identical batches: how far do they drift?
(run-to-run noise of the unordered fp32 gradient reductions, amplified by Adam, vs a race)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from segmentation_b200.models.unet import UNetModel
from configs_check import DS

S, B, STEPS = int(os.environ.get('S', 188)), int(os.environ.get('B', 2)), int(os.environ.get('STEPS', 3))
LR = float(os.environ.get('LR', 1e-3))
g = np.random.default_rng(0)
xs = [g.random((B, S, S, 3), dtype=np.float32) for _ in range(STEPS)]
ys = [(g.random((B, S, S, 1)) > 0.5).astype(np.uint8) for _ in range(STEPS)]
ms = [UNetModel(dataset=DS(B, S, 2), n_classes=2, input_dims=S, n_kernels=32, learning_rate=LR,
                load_snapshot=False, save_dir=None, seed=0) for _ in range(2)]
for t in range(STEPS):
    for m in ms:
        m.train_step((xs[t], ys[t]))
    la, lb = ms[0].seg_loss_op, ms[1].seg_loss_op
    sa, sb = ms[0].store.master, ms[1].store.master
    d = (sa - sb).abs()
    bad = int((d > 1e-5 + 1e-3 * sb.abs()).sum())
    print('step', t, 'loss', la, lb, 'diff %.3e' % abs(la - lb), 'params off', bad, 'of', d.numel(),
          'max %.3e' % float(d.max()), 'nonzero diffs', int((d > 0).sum()))
