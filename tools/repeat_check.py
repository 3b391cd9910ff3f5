"""Run-to-run repeatability of one U-Net forward + backward at a given size: forward
activations and input gradients must be bit-identical between two runs on the same inputs,
weight gradients equal to the noise of the unordered fp32 reductions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from segmentation_b200.models.unet import UNetModel
from configs_check import DS

S, B = int(os.environ.get('S', 188)), int(os.environ.get('B', 2))
g = np.random.default_rng(0)
x = torch.from_numpy(g.random((B, S, S, 3), dtype=np.float32)).cuda()
y = torch.from_numpy((g.random((B, S, S, 1)) > 0.5).astype(np.uint8)).cuda()
m = UNetModel(dataset=DS(B, S, 2), n_classes=2, input_dims=S, n_kernels=32, learning_rate=1e-3,
              load_snapshot=False, save_dir=None, seed=0)
ex = m._get_exec(B, True)
ex.use_graph = False
runs = []
for r in range(3):
    m.store.grad.zero_()
    ex.stage(x, y)
    ex.forward(); ex.loss(True); ex.backward()
    torch.cuda.synchronize()
    runs.append(({k: v.clone() for k, v in ex.act.items()}, {k: v.clone() for k, v in ex.g.items()},
                 m.store.grad.clone(), float(ex.loss_sum.item())))
a0, g0, w0, l0 = runs[0]
for r in (1, 2):
    a, gg, w, l = runs[r]
    bad_a = [k for k in a0 if not torch.equal(a[k].view(torch.int16), a0[k].view(torch.int16))]
    bad_g = [k for k in g0 if gg[k].dtype == torch.bfloat16 and
             not torch.equal(gg[k].view(torch.int16), g0[k].view(torch.int16))]
    rel = float((w - w0).double().norm() / w0.double().norm())
    print('run', r, 'loss', l, l0, 'acts differing', bad_a, 'grads differing', bad_g, 'wgrad rel', rel)
    for name, p in m.store.params.items():
        d = (w[p.offset:p.offset + p.numel] - w0[p.offset:p.offset + p.numel]).double()
        ref = w0[p.offset:p.offset + p.numel].double()
        e = float(d.norm() / (ref.norm() + 1e-30))
        if e > 1e-5:
            print('   ', name, 'rel %.3e' % e, 'max abs %.3e' % float(d.abs().max()))
