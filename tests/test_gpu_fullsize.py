"""-m gpu parity at the sizes bench.py and BASELINE.json's configs run, with the
production plan defaults (no seg_set_option overrides): kernel / tile-plan selection is
shape- and batch-dependent (umma_conv.cu launch_tconv -> launch_hconv -> launch_igemm,
tile-count-vs-148-SM choices), so the small-shape parity tests do not cover the plans the
benchmarked shapes pick.

  (a) U-Net 256x256 nk32 batch 16 (the bench configuration): fwd + bwd vs oracle.nets
  (b) FCN-8s 512x512, 21 classes, nk32 (BASELINE config 2): batch 2 vs the oracle, and a
      batch-16 run whose first two images reproduce the batch-2 logits
  (c) DeconvModel 1024x1024 inference (config 4): batch 2 label map bit-exact on
      margin-safe pixels vs the oracle; batch-32 run reproduces the batch-2 images
  (d) config 5: MC-dropout, T = 16 passes of a 512x512 tile, U-Net and DeconvModel

Reference graphs: /root/reference/models/unet.py:109-175, models/fcn.py:179-220,
models/deconvolution.py:101-178.  Tolerances as in test_gpu_unet.py (bf16 storage, fp32
accumulate; gradients anchored to the oracle's own bf16 noise floor).
"""
import os

import numpy as np
import pytest
import torch

from oracle import nets, tf_ops as T

from gpu_util import rel_l2, report, sync
from test_gpu_unet import FeedDataSet, oracle_noise_floor
from test_gpu_models import _nonzero_biases, noise_floor

pytestmark = pytest.mark.gpu


def _safe_labels_equal(lab, logits, logits_ref, lab_ref):
    """Label maps must agree on every pixel whose oracle label is stable under a logit
    perturbation of 4 x the largest logit error (each class moved up and down by that much:
    covers both the top-2 margin and the fp32 sigmoid saturation ties that
    argmax(sigmoid(.)) resolves to the first index, models/unet.py:76-77).  Returns
    (mismatches on stable pixels, unstable pixel count)."""
    tol = 4 * float((logits - logits_ref).abs().max())
    safe = torch.ones(lab_ref.shape[:-1], dtype=torch.bool)
    for c in range(logits_ref.shape[-1]):
        for sgn in (-1.0, 1.0):
            pert = logits_ref.clone()
            pert[..., c] += sgn * tol
            _, lab_p = T.sigmoid_argmax(pert)
            safe &= (lab_p[..., 0] == lab_ref[..., 0])
    bad = int((lab[..., 0][safe] != lab_ref[..., 0][safe]).sum())
    return bad, int((~safe).sum())


def test_unet_bench_config_bs16_parity(cuda):
    """(a) the exact bench.py workload: U-Net 256x256, n_kernels 32, 2 classes, batch 16."""
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.unet import UNetModel
    B, S, nk = 16, 256, 32
    ds = FeedDataSet(B, S, S)
    model = UNetModel(dataset=ds, n_classes=2, input_dims=S, n_kernels=nk, learning_rate=1e-4,
                      load_snapshot=False, save_dir=None)
    p = _nonzero_biases(nets.unet_params(n_kernels=nk, n_classes=2, seed=3), suffixes=('/biases',))
    model.load_weights({k: v.numpy() for k, v in p.items()})
    x, y = ds.next_batch()
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    ex = model._get_exec(B, True)
    ex.stage(xt.cuda(), yt.cuda())
    ex.forward()
    ex.loss(True)
    ex.backward()
    sync()
    assert tuple(ex.logits.shape) == (B, 68, 68, 2)
    taps = {}
    fwd = lambda q, xx: nets.unet_forward(q, xx, prec=T.BF16, taps=taps)
    loss_ref, logits_ref, grads_ref = nets.loss_and_grads(fwd, p, xt, yt)
    acts = {}
    for name, t in taps.items():
        if name == 'output':
            continue
        got, ref = ex.act[name].float().cpu(), t.detach()
        if name == 'conv1_2' and getattr(ex, 'c12_crop', False):
            y0, x0, h, w = ex.crop[4]
            got, ref = got[:, y0:y0 + h, x0:x0 + w], ref[:, y0:y0 + h, x0:x0 + w]
        if name == 'conv1_1' and getattr(ex, 'fuse_pool1', False):
            # conv1_1 + pool1 run as one launch: the full-resolution activation exists only
            # inside the window conv1_2 reads (pool1 itself is compared like every other tap)
            y0, x0, h, w = ex.crop[4]
            got, ref = got[:, y0:y0 + h + 2, x0:x0 + w + 2], ref[:, y0:y0 + h + 2, x0:x0 + w + 2]
        acts[name] = rel_l2(got, ref)
    floor = oracle_noise_floor(p, xt, yt, grads_ref)
    loss = float(ex.loss_sum.item()) / ex.loss_pixels
    rec = {'loss': loss, 'loss_ref': float(loss_ref), 'logits': rel_l2(ex.logits.cpu(), logits_ref),
           'acts_max': max(acts.values())}
    bad = []
    for name, gref in grads_ref.items():
        e = rel_l2(model.store.params[name].grad().cpu(), gref)
        rec['grad/' + name] = [e, floor[name]]
        if not e <= 3 * floor[name] + 5e-3:
            bad.append((name, e, floor[name]))
    report('unet_bs16', rec)
    assert max(acts.values()) < 1e-2, acts
    assert abs(loss - float(loss_ref)) < 2e-3 and rec['logits'] < 1e-2, rec
    assert not bad, bad
    _, lab = ex.head()
    sync()
    _, lab_ref = T.sigmoid_argmax(logits_ref)
    nbad, unsafe = _safe_labels_equal(lab.cpu(), ex.logits.cpu(), logits_ref, lab_ref)
    report('unet_bs16_labels', {'unsafe_pixels': unsafe, 'pixels': int(lab_ref.numel())})
    assert nbad == 0


def test_fcn8s_config2_full_size_parity(cuda):
    """(b) BASELINE config 2: FCN-8s 512x512, 21 classes, n_kernels 32."""
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.fcn import FCNModel
    S, nk, nc = 512, 32, 21
    ds2 = FeedDataSet(2, S, S, n_classes=nc, seed=11)
    model = FCNModel(dataset=ds2, n_classes=nc, input_dims=S, n_kernels=nk, fcn_type='8s',
                     load_snapshot=False, save_dir=None)
    p = _nonzero_biases(nets.fcn_params(n_kernels=nk, n_classes=nc, fcn_type='8s', seed=2))
    model.load_weights({k: v.numpy() for k, v in p.items()})
    x, y = ds2.next_batch()
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    ex = model._get_exec(2, True)
    ex.use_graph = False
    ex.stage(xt.cuda(), yt.cuda())
    ex.forward()
    ex.loss(True)
    ex.backward()
    sync()
    assert tuple(ex.logits.shape) == (2, S, S, nc)
    fwd = lambda q, xx: nets.fcn_forward(q, xx, fcn_type='8s', prec=T.BF16)
    loss_ref, logits_ref, grads_ref = nets.loss_and_grads(fwd, p, xt, yt, crop_mask=False)
    logits2 = ex.logits.cpu()
    loss = float(ex.loss_sum.item()) / ex.loss_pixels
    rec = {'logits': rel_l2(logits2, logits_ref), 'loss': loss, 'loss_ref': float(loss_ref)}
    floor = noise_floor(fwd, p, xt, yt, grads_ref, False)
    bad = []
    for name, gref in grads_ref.items():
        e = rel_l2(model.store.params[name].grad().cpu(), gref)
        rec['grad/' + name] = [e, floor[name]]
        if not e <= 3 * floor[name] + 1e-2:
            bad.append((name, e, floor[name]))
    report('fcn8s_config2', rec)
    assert rec['logits'] < 1e-2 and abs(loss - float(loss_ref)) < 2e-3, rec
    assert not bad, bad
    _, lab = ex.head()
    sync()
    _, lab_ref = T.sigmoid_argmax(logits_ref)
    nbad, unsafe = _safe_labels_equal(lab.cpu(), logits2, logits_ref, lab_ref)
    assert nbad == 0
    # the benchmarked batch: 16 images, the first two being the batch-2 images; images are
    # independent in this graph, so their logits must reproduce the batch-2 run (the plans
    # differ with the batch, hence rounding-level differences only)
    g = np.random.default_rng(12)
    x16 = np.concatenate([x, g.random((14, S, S, 3), dtype=np.float32)])
    y16 = np.concatenate([y, g.integers(0, nc, (14, S, S, 1)).astype(np.uint8)])
    ex16 = model._get_exec(16, True)
    ex16.use_graph = False
    model.store.grad.zero_()
    ex16.stage(torch.from_numpy(x16).cuda(), torch.from_numpy(y16).cuda())
    ex16.forward()
    ex16.loss(True)
    ex16.backward()
    sync()
    l16 = ex16.logits[:2].cpu()
    e16 = rel_l2(l16, logits2)
    report('fcn8s_config2_bs16', {'first2_vs_bs2': e16, 'first2_vs_oracle': rel_l2(l16, logits_ref),
                                  'loss16': float(ex16.loss_sum.item()) / ex16.loss_pixels})
    assert e16 < 3e-3 and rel_l2(l16, logits_ref) < 1e-2
    assert bool(torch.isfinite(model.store.grad).all())


def test_deconv_config4_inference_full_size(cuda):
    """(c) BASELINE config 4: DeconvModel 1024x1024 tiled inference, label-map output."""
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.deconvolution import DeconvModel
    S, nk = 1024, 32
    model = DeconvModel(mode='INFERENCE', n_classes=2, input_dims=S, n_kernels=nk,
                        load_snapshot=False, save_dir=None)
    p = _nonzero_biases(nets.deconv_params(n_kernels=nk, n_classes=2, seed=4))
    g = np.random.default_rng(3)
    for k in p:                                   # non-trivial moving statistics
        if k.endswith('moving_mean'):
            p[k] = torch.from_numpy(g.normal(0.2, 0.05, p[k].shape).astype(np.float32))
        if k.endswith('moving_variance'):
            p[k] = torch.from_numpy(g.uniform(0.05, 0.3, p[k].shape).astype(np.float32))
    model.load_weights({k: v.numpy() for k, v in p.items()})
    x = g.random((2, S, S, 3), dtype=np.float32)
    probs, lab = model.infer(x)
    assert probs.shape == (2, S, S, 2) and lab.shape == (2, S, S, 1)
    with torch.no_grad():
        logits_ref = nets.deconv_forward(p, torch.from_numpy(x), training=False, prec=T.BF16)
    sig_ref, lab_ref = T.sigmoid_argmax(logits_ref)
    logits = model._get_exec(2, False).logits.cpu()
    e = rel_l2(logits, logits_ref)
    nbad, unsafe = _safe_labels_equal(torch.from_numpy(lab), logits, logits_ref, lab_ref)
    report('deconv_config4', {'logits': e, 'unsafe_pixels': unsafe, 'pixels': int(lab_ref.numel()),
                              'mismatch_total': int((torch.from_numpy(lab) != lab_ref).sum())})
    assert e < 2e-2
    assert nbad == 0, (nbad, unsafe)
    # |d sigmoid / d logit| <= 1/4
    atol = 0.25 * float((logits - logits_ref).abs().max()) + 1e-6
    assert np.allclose(probs, sig_ref.numpy(), atol=atol)
    # the benchmarked batch (32 tiles): the first two tiles reproduce the batch-2 run
    x32 = np.concatenate([x, g.random((30, S, S, 3), dtype=np.float32)])
    probs32, lab32 = model.infer(x32)
    l32 = model._get_exec(32, False).logits[:2].cpu()
    e32 = rel_l2(l32, logits)
    report('deconv_config4_bs32', {'first2_vs_bs2': e32})
    assert e32 < 3e-3
    nbad32, _ = _safe_labels_equal(torch.from_numpy(lab32[:2]), l32, logits_ref, lab_ref)
    assert nbad32 == 0
    assert set(np.unique(lab32)).issubset({0.0, 1.0})


def test_config5_mc_dropout_512_T16(cuda):
    """(d) BASELINE config 5: 16 stochastic passes of one 512x512 tile as one batch, mean
    and variance maps; U-Net (build-defined sites) and DeconvModel (reference sites)."""
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.deconvolution import DeconvModel
    from segmentation_b200.models.unet import UNetModel
    S, Tn = 512, 16
    g = np.random.default_rng(5)
    x = g.random((1, S, S, 3), dtype=np.float32)
    xt = torch.from_numpy(x)
    # ---- U-Net: three of the sixteen passes against the oracle (56 GFLOP each on the CPU)
    um = UNetModel(mode='INFERENCE', n_classes=2, input_dims=S, n_kernels=32, bayesian=True,
                   load_snapshot=False, save_dir=None)
    pu = nets.unet_params(n_kernels=32, n_classes=2, seed=3)
    um.load_weights({k: v.numpy() for k, v in pu.items()})
    mean, var, probs = um.infer_mc(x, passes=Tn, seed=11, pass_offset=2)
    assert mean.shape == (324, 324, 2) and var.shape == (324, 324, 2) and probs.shape[0] == Tn
    errs = {}
    with torch.no_grad():
        for t in (0, 7, 15):
            ref = torch.sigmoid(nets.unet_forward(pu, xt, prec=T.BF16, dropout=(11, 2 + t)))[0]
            errs[t] = rel_l2(torch.from_numpy(probs[t]), ref)
    report('config5_unet', {'pass_err': errs})
    assert max(errs.values()) < 1e-2, errs
    assert np.allclose(mean, probs.mean(0), atol=1e-6) and np.allclose(var, probs.var(0), atol=1e-6)
    assert float(var.max()) > 0
    # ---- DeconvModel: all sixteen passes
    dm = DeconvModel(mode='INFERENCE', n_classes=2, input_dims=S, n_kernels=32, bayesian=True,
                     load_snapshot=False, save_dir=None)
    pd_ = _nonzero_biases(nets.deconv_params(n_kernels=32, n_classes=2, seed=4))
    for k in pd_:
        if k.endswith('moving_mean'):
            pd_[k] = torch.from_numpy(g.normal(0.2, 0.05, pd_[k].shape).astype(np.float32))
        if k.endswith('moving_variance'):
            pd_[k] = torch.from_numpy(g.uniform(0.05, 0.3, pd_[k].shape).astype(np.float32))
    dm.load_weights({k: v.numpy() for k, v in pd_.items()})
    mean, var, probs = dm.infer_mc(x, passes=Tn, seed=7, pass_offset=1)
    assert mean.shape == (S, S, 2) and probs.shape == (Tn, S, S, 2)
    with torch.no_grad():
        ref = torch.stack([torch.sigmoid(nets.deconv_forward(pd_, xt, training=False, bayesian=True,
                                                             prec=T.BF16, dropout=(7, 1 + t)))[0]
                           for t in range(Tn)])
    e = rel_l2(torch.from_numpy(probs), ref)
    report('config5_deconv', {'probs': e})
    assert e < 1e-2
    assert np.allclose(mean, probs.mean(0), atol=1e-6) and np.allclose(var, probs.var(0), atol=1e-6)
    assert float(var.max()) > 0
