"""BASELINE.json configs 2, 4, 5 at their full sizes on one B200: they must run, and the
size-independent properties must hold (finite loss that goes down, label maps in range,
MC variance >= 0, deterministic repeat).  Prints one JSON line per config with timings
(CUDA events; these are parity-test configurations, not the bench metric)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from segmentation_b200.models.fcn import FCNModel
from segmentation_b200.models.deconvolution import DeconvModel
from segmentation_b200.models.unet import UNetModel


class DS(object):
    use_feed, has_masks = False, True

    def __init__(self, b, s, ncls, seed=0):
        self.batch_size = b
        g = np.random.default_rng(seed)
        self.x = torch.from_numpy(g.random((b, s, s, 3), dtype=np.float32)).pin_memory()
        self.y = torch.from_numpy(g.integers(0, ncls, (b, s, s, 1)).astype(np.uint8)).pin_memory()

    def set_tf_sess(self, s):
        pass

    def next_batch(self):
        return self.x, self.y


def timed(fn, n):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = []
which = (sys.argv[1:] if __name__ == "__main__" else ["none"]) or ['2', '4', '5u', '5d']
if '2' in which:
    ds = DS(16, 512, 21)
    m = FCNModel(None, dataset=ds, n_classes=21, fcn_type='8s', input_dims=512, n_kernels=32,
                 learning_rate=1e-4, load_snapshot=False, save_dir=None)
    m.train_step(); l0 = m.seg_loss_op
    for _ in range(3):
        m.train_step()
    ms = timed(lambda: m.train_step(), 10)
    l1 = m.seg_loss_op
    assert np.isfinite(l0) and np.isfinite(l1) and l1 < l0 + 1e-3, (l0, l1)
    out.append({'config': 2, 'what': 'FCN-8s 512x512 21 classes bs16 train step', 'ms_per_step': ms,
                'img_per_s': 16 / ms * 1e3, 'loss_first': l0, 'loss_last': l1,
                'algorithmic_tflops': 455.0 / ms, 'launches_per_step': None})
    del m; torch.cuda.empty_cache()
if '4' in which:
    g = np.random.default_rng(1)
    x = g.random((32, 1024, 1024, 3), dtype=np.float32)
    m = DeconvModel(None, dataset=None, n_classes=2, input_dims=1024, n_kernels=32, mode='INFERENCE',
                    load_snapshot=False, save_dir=None)
    xd = torch.from_numpy(x).cuda()
    ex = m._get_exec(32, False)
    def run():
        ex.infer(xd)
    ms = timed(run, 5)
    probs, labels = m.infer(x[:32])
    assert labels.shape == (32, 1024, 1024, 1) and set(np.unique(labels)) <= {0.0, 1.0}
    probs2, labels2 = m.infer(x[:32])
    assert np.array_equal(labels, labels2) and np.array_equal(probs, probs2)
    out.append({'config': 4, 'what': 'DeconvModel 1024x1024 bs32 inference (device-resident input, forward+head)',
                'ms_per_batch': ms, 'img_per_s': 32 / ms * 1e3, 'algorithmic_tflops': 211.7 / ms})
    del m, ex, xd; torch.cuda.empty_cache()
if '5u' in which:
    g = np.random.default_rng(2)
    x = g.random((1, 512, 512, 3), dtype=np.float32)
    m = UNetModel(None, dataset=None, n_classes=2, input_dims=512, n_kernels=32, mode='INFERENCE',
                  bayesian=True, load_snapshot=False, save_dir=None)
    t0 = time.time(); mean, var, probs = m.infer_mc(x, passes=16, seed=0); torch.cuda.synchronize()
    mean2, var2, _ = m.infer_mc(x, passes=16, seed=0)
    ms = timed(lambda: m.infer_mc(x, passes=16, seed=0, return_probs=False), 5)
    assert mean.shape == (324, 324, 2) and (var >= 0).all() and np.array_equal(mean, mean2)
    assert np.allclose(mean, probs.mean(0), atol=1e-5) and np.allclose(var, probs.var(0), atol=1e-5)
    out.append({'config': 5, 'what': 'U-Net MC-dropout 16 passes 512x512 tile -> mean/var 324x324x2 (host in, mean+var out)',
                'ms_per_tile': ms, 'var_max': float(var.max())})
    del m; torch.cuda.empty_cache()
if '5d' in which:
    g = np.random.default_rng(3)
    x = g.random((1, 512, 512, 3), dtype=np.float32)
    m = DeconvModel(None, dataset=None, n_classes=2, input_dims=512, n_kernels=32, mode='INFERENCE',
                    bayesian=True, load_snapshot=False, save_dir=None)
    mean, var, probs = m.infer_mc(x, passes=16, seed=0)
    ms = timed(lambda: m.infer_mc(x, passes=16, seed=0, return_probs=False), 5)
    assert mean.shape == (512, 512, 2) and (var >= 0).all()
    assert np.allclose(mean, probs.mean(0), atol=1e-5) and np.allclose(var, probs.var(0), atol=1e-5)
    out.append({'config': 5, 'what': 'DeconvModel MC-dropout 16 passes 512x512 tile (host in, mean+var out)',
                'ms_per_tile': ms, 'var_max': float(var.max())})
os.makedirs('gpurun_out', exist_ok=True)
for o in out:
    print(json.dumps(o), flush=True)
json.dump(out, open('gpurun_out/configs_check.json', 'w'))
