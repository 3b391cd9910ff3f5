"""-m gpu whole-network parity for FCNModel, DeconvModel (incl. Bayesian MC mode) and
the utils/ops.py helpers against the CPU oracle.  Gradient tolerances are anchored
to the oracle's own bf16 noise floor (see test_gpu_unet.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import nets, tf_ops as T
from segmentation_b200 import native as N

from gpu_util import bfr, rel_l2, report, sync
from test_gpu_unet import FeedDataSet

pytestmark = pytest.mark.gpu


def noise_floor(fwd, p, xt, yt, grads_ref, crop_mask):
    c32, d32 = T.conv2d, T.conv2d_transpose

    def c64(x, w, b=None, stride=1, padding='SAME'):
        return c32(x.double(), w.double(), None if b is None else b.double(), stride,
                   padding).float()

    def d64(x, w, b=None, stride=2, padding='VALID'):
        return d32(x.double(), w.double(), None if b is None else b.double(), stride,
                   padding).float()

    T.conv2d, T.conv2d_transpose = c64, d64
    try:
        _, _, g64 = nets.loss_and_grads(fwd, p, xt, yt, crop_mask)
    finally:
        T.conv2d, T.conv2d_transpose = c32, d32
    return {k: rel_l2(g64[k], grads_ref[k]) for k in grads_ref}


def _nonzero_biases(p, seed=5, suffixes=('/biases', '/beta')):
    gen = np.random.default_rng(seed)
    for k in p:
        if k.endswith(suffixes):
            p[k] = torch.from_numpy(gen.normal(0, 0.05, p[k].shape).astype(np.float32))
    return p


def _run_fwd_bwd(model, B, xt, yt):
    ex = model._get_exec(B, True)
    ex.use_graph = False
    ex.stage(xt.cuda(), yt.cuda())
    ex.forward()
    ex.loss(True)
    ex.backward()
    sync()
    return ex


def _check_grads(model, grads_ref, floor, slack, tag):
    rec, bad = {}, []
    for name, gref in grads_ref.items():
        e = rel_l2(model.store.params[name].grad().cpu(), gref)
        rec[name] = [e, floor[name]]
        if not e <= 3 * floor[name] + slack:
            bad.append((name, e, floor[name]))
    report(tag, rec)
    return bad


@pytest.mark.parametrize('fcn_type', ['8s', '16s', '32s'])
def test_fcn_forward_backward_parity(cuda, fcn_type):
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.fcn import FCNModel
    B, S, nk, nc = 2, 64, 16, 5
    ds = FeedDataSet(B, S, S, n_classes=nc)
    model = FCNModel(dataset=ds, n_classes=nc, input_dims=S, n_kernels=nk, fcn_type=fcn_type,
                     load_snapshot=False, save_dir=None)
    p = _nonzero_biases(nets.fcn_params(n_kernels=nk, n_classes=nc, fcn_type=fcn_type, seed=2))
    model.load_weights({k: v.numpy() for k, v in p.items()})
    x, y = ds.next_batch()
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    ex = _run_fwd_bwd(model, B, xt, yt)
    taps = {}
    fwd = lambda q, xx: nets.fcn_forward(q, xx, fcn_type=fcn_type, prec=T.BF16, taps=taps)
    loss_ref, logits_ref, grads_ref = nets.loss_and_grads(fwd, p, xt, yt, crop_mask=False)
    e_logits = rel_l2(ex.logits.cpu(), logits_ref)
    loss = float(ex.loss_sum.item()) / ex.loss_pixels
    acts = {k: rel_l2(ex.act[k].float().cpu()[..., :t.shape[-1]], t.detach())
            for k, t in taps.items() if k in ex.act}
    report('fcn_parity', {'type': fcn_type, 'logits': e_logits, 'loss': loss,
                          'loss_ref': float(loss_ref), 'acts': acts})
    assert tuple(ex.logits.shape) == (B, S, S, nc)
    assert max(acts.values()) < 1e-2, acts
    assert e_logits < 1e-2 and abs(loss - float(loss_ref)) < 2e-3
    floor = noise_floor(lambda q, xx: nets.fcn_forward(q, xx, fcn_type=fcn_type, prec=T.BF16),
                        p, xt, yt, grads_ref, False)
    bad = _check_grads(model, grads_ref, floor, 1.5e-2, 'fcn_grads_' + fcn_type)
    assert not bad, bad
    # label map, first-index argmax over 5 classes, compared where the margin is safe
    probs, lab = ex.head()
    sync()
    _, lab_ref = T.sigmoid_argmax(logits_ref)
    top2 = torch.topk(logits_ref, 2, dim=-1).values
    safe = (top2[..., 0] - top2[..., 1]) > 4 * float((ex.logits.cpu() - logits_ref).abs().max())
    assert torch.equal(lab.cpu()[..., 0][safe], lab_ref[..., 0][safe])


def test_fcn_train_step_runs_and_tracks_oracle_loss(cuda):
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.fcn import FCNModel
    B, S, nk, nc = 2, 64, 16, 21
    ds = FeedDataSet(B, S, S, n_classes=nc)
    model = FCNModel(dataset=ds, n_classes=nc, input_dims=S, n_kernels=nk, fcn_type='8s',
                     learning_rate=1e-3, load_snapshot=False, save_dir=None)
    p = nets.fcn_params(n_kernels=nk, n_classes=nc, fcn_type='8s', seed=2)
    model.load_weights({k: v.numpy() for k, v in p.items()})
    state = nets.AdamState(p)
    ds_ref = FeedDataSet(B, S, S, n_classes=nc)
    fwd = lambda q, xx: nets.fcn_forward(q, xx, fcn_type='8s', prec=T.BF16)
    for it in range(3):
        model.train_step()
        x, y = ds_ref.next_batch()
        ref = nets.train_step(fwd, p, state, torch.from_numpy(x), torch.from_numpy(y), lr=1e-3,
                              crop_mask=False)
        assert abs(model.seg_loss_op - ref) < 5e-3, (it, model.seg_loss_op, ref)
    assert model.global_step == 3


def test_fcn8s_fused_loss_head_matches_unfused_gradients(cuda):
    """A train step of FCN-8s takes forward(fused_loss=True): upscore x8 + loss + their gradient
    in one launch (seg_upscore8_xent_fwd_bwd).  Same loss and the same parameter gradients as
    forward() + loss() + backward() - the score-map gradient is bit-identical, what follows
    differs by the order of the weight-gradient reductions only."""
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.fcn import FCNModel
    B, S, nk, nc = 2, 128, 16, 21
    ds = FeedDataSet(B, S, S, n_classes=nc, seed=5)
    model = FCNModel(dataset=ds, n_classes=nc, input_dims=S, n_kernels=nk, fcn_type='8s',
                     load_snapshot=False, save_dir=None)
    p = _nonzero_biases(nets.fcn_params(n_kernels=nk, n_classes=nc, fcn_type='8s', seed=2))
    model.load_weights({k: v.numpy() for k, v in p.items()})
    x, y = ds.next_batch()
    ex = _run_fwd_bwd(model, B, torch.from_numpy(x), torch.from_numpy(y))
    loss_u = float(ex.loss_sum.item())
    g_fuse3 = ex.g['fuse3'].clone()
    grads_u = model.store.grad.clone()
    model.store.grad.zero_()
    n0 = N.LAUNCHES
    ex.forward(fused_loss=True)
    assert ex._fused_loss
    ex.loss(True)
    ex.backward()
    sync()
    n_fused = N.LAUNCHES - n0
    loss_f = float(ex.loss_sum.item())
    assert torch.equal(ex.g['fuse3'].view(torch.int16), g_fuse3.view(torch.int16))
    assert abs(loss_f - loss_u) <= 1e-5 * abs(loss_u)
    e = rel_l2(model.store.grad.cpu(), grads_u.cpu())
    report('fcn8s_fused_loss', {'grad_rel_l2': e, 'loss': [loss_u, loss_f], 'launches': n_fused})
    assert e < 1e-4, e
    # and a whole train step takes that route
    model.store.grad.zero_()
    model.train_step((x, y))
    assert model._last_train_exec._fused_loss


def _deconv_model(bayesian, mode='TRAINING', B=2, S=256, nk=16):
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.deconvolution import DeconvModel
    ds = FeedDataSet(B, S, S) if mode == 'TRAINING' else None
    model = DeconvModel(dataset=ds, n_classes=2, input_dims=S, n_kernels=nk, bayesian=bayesian,
                        mode=mode, load_snapshot=False, save_dir=None)
    p = _nonzero_biases(nets.deconv_params(n_kernels=nk, n_classes=2, seed=4))
    model.load_weights({k: v.numpy() for k, v in p.items()})
    return model, ds, p


@pytest.mark.parametrize('bayesian', [False, True])
def test_deconv_forward_backward_parity(cuda, bayesian):
    model, ds, p = _deconv_model(bayesian)
    x, y = ds.next_batch()
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    ex = _run_fwd_bwd(model, 2, xt, yt)
    taps, stats = {}, {}
    kw = dict(training=True, bayesian=bayesian, prec=T.BF16, dropout=(0, 0))
    fwd = lambda q, xx: nets.deconv_forward(q, xx, taps=taps, new_stats=stats, **kw)
    loss_ref, logits_ref, grads_ref = nets.loss_and_grads(fwd, p, xt, yt, crop_mask=False)
    # with dropout the model drops the bn2/bn4/bn5 buffers in place, the oracle taps
    # hold the pre-dropout tensors: skip those three (their consumers are compared)
    skip = ('bn2', 'bn4', 'bn5') if bayesian else ()
    acts = {k: rel_l2(ex.act[k].float().cpu()[..., :t.shape[-1]], t.detach())
            for k, t in taps.items() if k in ex.act and k not in skip}
    e_logits = rel_l2(ex.logits.cpu(), logits_ref)
    loss = float(ex.loss_sum.item()) / ex.loss_pixels
    report('deconv_parity', {'bayesian': bayesian, 'logits': e_logits, 'loss': loss,
                             'loss_ref': float(loss_ref), 'acts': acts})
    assert tuple(ex.logits.shape) == (2, 256, 256, 2)
    assert max(acts.values()) < 2e-2, acts
    assert e_logits < 2e-2 and abs(loss - float(loss_ref)) < 3e-3
    # moving statistics updated like slim (decay .999)
    for name in ('bn1', 'bn4', 'bn8'):
        mm = model.store.state[name + '/moving_mean'].cpu()
        mv = model.store.state[name + '/moving_variance'].cpu()
        assert torch.allclose(mm, stats[name + '/moving_mean'], atol=2e-5)
        assert torch.allclose(mv, stats[name + '/moving_variance'], atol=2e-5)
    floor = noise_floor(lambda q, xx: nets.deconv_forward(q, xx, **kw), p, xt, yt, grads_ref, False)
    bad = _check_grads(model, grads_ref, floor, 2e-2, 'deconv_grads_%s' % bayesian)
    assert not bad, bad


def test_deconv_inference_uses_moving_stats_and_mc_dropout(cuda):
    model, _, p = _deconv_model(True, mode='INFERENCE', S=256)
    g = np.random.default_rng(3)
    # non-trivial moving statistics
    for k in p:
        if k.endswith('moving_mean'):
            p[k] = torch.from_numpy(g.normal(0.2, 0.05, p[k].shape).astype(np.float32))
        if k.endswith('moving_variance'):
            p[k] = torch.from_numpy(g.uniform(0.05, 0.3, p[k].shape).astype(np.float32))
    model.load_weights({k: v.numpy() for k, v in p.items()})
    x = g.random((1, 256, 256, 3), dtype=np.float32)
    T_passes = 4
    mean, var, probs = model.infer_mc(x, passes=T_passes, seed=7, pass_offset=3)
    assert mean.shape == (256, 256, 2) and var.shape == (256, 256, 2)
    ref = []
    for t in range(T_passes):
        lg = nets.deconv_forward(p, torch.from_numpy(x), training=False, bayesian=True, prec=T.BF16,
                                 dropout=(7, 3 + t))
        ref.append(torch.sigmoid(lg)[0])
    ref = torch.stack(ref)
    e = rel_l2(torch.from_numpy(probs), ref)
    report('deconv_mc', {'probs': e})
    assert e < 1e-2
    assert np.allclose(mean, probs.mean(0), atol=1e-6)
    assert np.allclose(var, probs.var(0), atol=1e-6)
    assert float(var.max()) > 0          # passes differ: dropout is live at inference


def test_unet_mc_dropout_mean_variance(cuda):
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.unet import UNetModel
    model = UNetModel(mode='INFERENCE', n_classes=2, input_dims=188, n_kernels=16, bayesian=True,
                      load_snapshot=False, save_dir=None)
    p = nets.unet_params(n_kernels=16, n_classes=2, seed=3)
    model.load_weights({k: v.numpy() for k, v in p.items()})
    x = np.random.default_rng(1).random((1, 188, 188, 3), dtype=np.float32)
    mean, var, probs = model.infer_mc(x, passes=3, seed=11, pass_offset=0)
    ref = torch.stack([torch.sigmoid(nets.unet_forward(p, torch.from_numpy(x), prec=T.BF16,
                                                       dropout=(11, t)))[0] for t in range(3)])
    assert rel_l2(torch.from_numpy(probs), ref) < 1e-2
    assert np.allclose(mean, probs.mean(0), atol=1e-6) and np.allclose(var, probs.var(0), atol=1e-6)


def test_ops_helpers(cuda):
    """utils/ops.py mirror: conv2d / deconv2d / linear / batch_norm / lrelu / concat."""
    from segmentation_b200.utils import ops
    ops.reset_variables(0)
    g = torch.Generator().manual_seed(0)
    x = bfr(torch.randn(2, 12, 12, 8, generator=g))
    y = ops.conv2d(x.cuda(), 24, name='d_h0_conv')
    w = ops.get_variable('d_h0_conv/w').detach().cpu()
    b = ops.get_variable('d_h0_conv/biases').detach().cpu()
    assert tuple(w.shape) == (5, 5, 8, 24) and float(w.abs().max()) <= 0.04 + 1e-6
    ref = T.conv2d(x, bfr(w), b, 2, 'SAME')
    assert tuple(y.shape) == (2, 6, 6, 24) and rel_l2(y.float().cpu(), ref) < 4e-3
    y2 = ops.conv2d(x.cuda(), 24, name='d_h0_conv')                  # reuse
    assert torch.equal(y2, y)
    d, dw, db = ops.deconv2d(y, [2, 12, 12, 16], name='g_h1', with_w=True)
    ref = T.conv2d_transpose(bfr(y.detach().float().cpu()), bfr(dw.detach().cpu()),
                             db.detach().cpu(), 2, 'SAME')
    assert tuple(d.shape) == (2, 12, 12, 16) and rel_l2(d.detach().float().cpu(), ref) < 4e-3
    z = bfr(torch.randn(4, 40, generator=g))
    out, M, bias = ops.linear(z.cuda(), 10, 'g_h0_lin', with_w=True)
    assert rel_l2(out.detach().cpu(), z @ bfr(M.detach().cpu()) + bias.detach().cpu()) < 1e-4
    bn = ops.batch_norm(name='g_bn0')
    yb = bn(x.cuda(), train=True)
    ref, m_ref, v_ref = T.batch_norm(x, torch.zeros(8), torch.zeros(8), torch.ones(8), True,
                                     decay=0.9, eps=1e-5, gamma=torch.ones(8))
    assert rel_l2(yb.detach().float().cpu(), ref) < 4e-3
    assert torch.allclose(ops.get_variable('g_bn0/moving_mean').cpu(), m_ref, atol=1e-5)
    assert torch.equal(ops.lrelu(torch.tensor([-1.0, 2.0])), torch.tensor([-0.2, 2.0]))
    cc = ops.conv_cond_concat(x.cuda(), torch.ones(2, 1, 1, 3).cuda())
    assert tuple(cc.shape) == (2, 12, 12, 11)


def test_ops_helpers_are_differentiable(cuda):
    """The utils/ops.py helpers are differentiable like their TF originals (reference
    utils/ops.py:58-110 build tf ops): a small DCGAN-style graph conv2d (3x3/s1 and the default
    5x5/s2) -> lrelu -> batch_norm -> deconv2d (5x5/s2 SAME and 2x2/s2) -> linear, gradients of
    a scalar w.r.t. every variable and the input against torch autograd on the CPU oracle ops
    with bf16 rounding (straight-through) at the same op boundaries."""
    from segmentation_b200.utils import ops
    ops.reset_variables(1)
    g = torch.Generator().manual_seed(3)
    x = bfr(torch.randn(2, 8, 8, 8, generator=g))
    coef = torch.randn(2, 5, generator=g)

    xd = x.cuda().requires_grad_(True)
    h0 = ops.lrelu(ops.conv2d(xd, 16, k_h=3, k_w=3, d_h=1, d_w=1, stddev=0.2, name='c0'))
    h1 = ops.conv2d(h0, 16, stddev=0.1, name='c1')
    bn = ops.batch_norm(name='bn0')
    h2 = ops.lrelu(bn(h1, train=True))
    h3 = ops.deconv2d(h2, [2, 8, 8, 8], stddev=0.1, name='d0')
    h4 = ops.deconv2d(h3, [2, 16, 16, 4], k_h=2, k_w=2, stddev=0.3, name='d1')
    out = ops.linear(h4.reshape(2, -1).float(), 5, 'lin', stddev=0.1)
    (out * coef.cuda()).sum().backward()
    sync()

    def ste(t):                                   # bf16 rounding, gradient passes through
        return t + (bfr(t.detach()) - t.detach())

    names = ['c0/w', 'c0/biases', 'c1/w', 'c1/biases', 'bn0/gamma', 'bn0/beta', 'd0/w', 'd0/biases',
             'd1/w', 'd1/biases', 'lin/Matrix', 'lin/bias']
    V = {n: ops.get_variable(n).detach().cpu().clone().requires_grad_(True) for n in names}
    xr = x.clone().requires_grad_(True)
    r0 = ste(T.conv2d(xr, ste(V['c0/w']), V['c0/biases'], 1, 'SAME'))
    r0 = torch.maximum(r0, 0.2 * r0)
    r1 = ste(T.conv2d(r0, ste(V['c1/w']), V['c1/biases'], 2, 'SAME'))
    r2, _, _ = T.batch_norm(r1, V['bn0/beta'], torch.zeros(16), torch.ones(16), True, decay=0.9,
                            eps=1e-5, gamma=V['bn0/gamma'])
    r2 = ste(r2)
    r2 = torch.maximum(r2, 0.2 * r2)
    r3 = ste(T.conv2d_transpose(r2, ste(V['d0/w']), V['d0/biases'], 2, 'SAME'))
    r4 = ste(T.conv2d_transpose(r3, ste(V['d1/w']), V['d1/biases'], 2, 'SAME'))
    ro = ste(r4.reshape(2, -1)) @ ste(V['lin/Matrix']) + V['lin/bias']
    (ro * coef).sum().backward()
    assert rel_l2(out.detach().cpu(), ro.detach()) < 1e-2
    rec = {}
    for n in names:
        rec[n] = rel_l2(ops.get_variable(n).grad.cpu(), V[n].grad)
    rec['x'] = rel_l2(xd.grad.cpu(), xr.grad)
    report('ops_backward', rec)
    # the bias in front of a batch-norm has a gradient of exactly zero (the layer removes the
    # mean): both sides hold rounding noise there, compare it with the layer's weight gradient
    rec.pop('c1/biases')
    noise = float(ops.get_variable('c1/biases').grad.abs().max())
    assert noise < 1e-2 * float(ops.get_variable('c1/w').grad.abs().max()), noise
    assert max(rec.values()) < 3e-2, rec


def test_deconv_fused_tail_matches_unfused(cuda):
    """DeconvModel inference: the one-launch class-map tail (seg_classmap_tail_infer: resize ->
    deconv3_0 -> bn8 -> conv_out -> sigmoid/argmax) against the five separate C-ABI calls.
    Every intermediate is rounded to bf16 at the same points, so logits agree to fp32
    summation order and the label maps wherever the class margin exceeds that noise."""
    import os
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.deconvolution import DeconvModel
    S, nk, B = 256, 32, 3
    model = DeconvModel(mode='INFERENCE', n_classes=2, input_dims=S, n_kernels=nk,
                        load_snapshot=False, save_dir=None)
    p = _nonzero_biases(nets.deconv_params(n_kernels=nk, n_classes=2, seed=9))
    g = np.random.default_rng(4)
    for k in p:
        if k.endswith('moving_mean'):
            p[k] = torch.from_numpy(g.normal(0.2, 0.05, p[k].shape).astype(np.float32))
        if k.endswith('moving_variance'):
            p[k] = torch.from_numpy(g.uniform(0.05, 0.3, p[k].shape).astype(np.float32))
    model.load_weights({k: v.numpy() for k, v in p.items()})
    x = g.random((B, S, S, 3), dtype=np.float32)
    ex = model._get_exec(B, False)
    probs_f, lab_f = model.infer(x)
    assert ex._tail_done
    logits_f = ex.logits.cpu().clone()
    os.environ['SEGB200_FUSED_TAIL'] = '0'
    try:
        probs_u, lab_u = model.infer(x)
        assert not ex._tail_done
        logits_u = ex.logits.cpu().clone()
    finally:
        del os.environ['SEGB200_FUSED_TAIL']
    # ... and the one-launch head (conv1_0 + bn1 + pool1 on the (R,G,B,1) input) against the
    # patch-packed convolution + batch-norm + pool launches
    os.environ['SEGB200_FUSED_HEAD1'] = '0'
    try:
        model._exec.clear()
        model.infer(x)
        ex2 = model._get_exec(B, False)
        assert not ex2.head_fused and ex.head_fused
        e_head = rel_l2(ex2.logits.cpu(), logits_f)
    finally:
        del os.environ['SEGB200_FUSED_HEAD1']
    report('deconv_fused_head', {'logits_rel_l2': e_head})
    assert e_head < 1e-2, e_head
    # ... and the batch-norms folded into their neighbours (bn2 / bn3 applied to the pooled
    # tensor, bn4..bn7 as the producing layer's epilogue) against separate normalisation passes
    os.environ['SEGB200_FUSE_BN'] = '0'
    try:
        model._exec.clear()
        n0 = N.LAUNCHES
        model.infer(x)
        n_unfused = N.LAUNCHES - n0
        ex3 = model._get_exec(B, False)
        e_bn = rel_l2(ex3.logits.cpu(), logits_f)
        lab_bn = ex3.labelmap.cpu().clone()
    finally:
        del os.environ['SEGB200_FUSE_BN']
    model._exec.clear()
    n0 = N.LAUNCHES
    _, lab_f2 = model.infer(x)
    n_fused = N.LAUNCHES - n0
    report('deconv_fused_bn', {'logits_rel_l2': e_bn, 'launches': [n_unfused, n_fused],
                               'label_mismatch': float((lab_bn != torch.as_tensor(np.asarray(lab_f2))).float().mean())})
    assert e_bn < 1e-2, e_bn
    assert n_fused < n_unfused                      # two launches fewer (bn2, bn3 passes)
    e = rel_l2(logits_f, logits_u)
    dmax = float((logits_f - logits_u).abs().max())
    margin = (logits_u[..., 0] - logits_u[..., 1]).abs()
    safe = margin > 4 * dmax + 1e-6
    report('deconv_fused_tail', {'logits_rel_l2': e, 'logits_max_abs': dmax,
                                 'unsafe': int((~safe).sum()), 'pixels': int(safe.numel())})
    assert e < 2e-3, e
    assert np.allclose(probs_f, probs_u, atol=0.25 * dmax + 1e-6)
    assert bool((torch.from_numpy(lab_f)[..., 0][safe] == torch.from_numpy(lab_u)[..., 0][safe]).all())
