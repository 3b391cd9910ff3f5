"""-m gpu parity tests for the memory-bound kernels against the CPU oracle.

Bit-exact: max-pool values + argmax slots (first max, row-major window scan),
dropout keep-masks, label maps.  Tolerances for floating point are stated at
each assert."""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from segmentation_b200 import engine as E
from segmentation_b200 import native as N

from gpu_util import bfr, dev_bf16, rel_l2, report, sync

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


@pytest.mark.parametrize('shape,k', [((2, 9, 11, 32), 2), ((1, 14, 14, 64), 2), ((2, 10, 13, 32), 3),
                                     ((1, 7, 7, 5), 2), ((1, 254, 254, 32), 2)])
def test_maxpool_fwd_bwd_bit_exact(cuda, shape, k):
    g = _gen(5)
    x = bfr(torch.randn(shape, generator=g))
    # force ties: quantise + ReLU-like zeros (all-zero windows are the common tie)
    x = torch.relu(torch.round(x * 2) / 2)
    y_ref, slot_ref = T.max_pool_with_argmax(x, k, k)
    x_d = dev_bf16(x)
    Nb, H, W, C = shape
    Ho, Wo = y_ref.shape[1], y_ref.shape[2]
    y_d = torch.zeros(Nb, Ho, Wo, C, dtype=BF16, device='cuda')
    am = torch.zeros(Nb, Ho, Wo, C, dtype=torch.uint8, device='cuda')
    E.maxpool_fwd(x_d, y_d, am, k, k)
    sync()
    assert torch.equal(y_d.float().cpu(), y_ref)                 # bit-exact values
    assert torch.equal(am.cpu(), slot_ref)                       # bit-exact argmax
    # backward: gather by argmax + skip add + relu mask
    dy = bfr(torch.randn(y_ref.shape, generator=g))
    xr = x.clone().requires_grad_(True)
    yr = T.max_pool(xr, k, k)
    (dx_ref,) = torch.autograd.grad(yr, xr, dy)
    add = bfr(torch.randn(Nb, H - 2, W - 2, C, generator=g))
    addp = torch.zeros(shape)
    addp[:, 1:H - 1, 1:W - 1] = add
    ref = bfr((dx_ref + addp) * (x > 0).float())
    dx_d = torch.zeros(shape, dtype=BF16, device='cuda')
    E.maxpool_bwd(dev_bf16(dy), am, dx_d, k, k, add=dev_bf16(add), add_y0=1, add_x0=1, mask=x_d)
    sync()
    assert torch.equal(dx_d.float().cpu(), ref)
    # same result when the mask of the routed part comes from the pool output (the pool input
    # is then read only inside the add window); a smaller add window exercises both branches
    add2 = bfr(torch.randn(Nb, H - 4, W - 4, C, generator=g))
    addp2 = torch.zeros(shape)
    addp2[:, 2:H - 2, 2:W - 2] = add2
    ref2 = bfr((dx_ref + addp2) * (x > 0).float())
    dx_d.fill_(7.0)
    E.maxpool_bwd(dev_bf16(dy), am, dx_d, k, k, add=dev_bf16(add2), add_y0=2, add_x0=2, mask=x_d,
                  pooled=y_d)
    sync()
    assert torch.equal(dx_d.float().cpu(), ref2)


@pytest.mark.parametrize('shape,k', [((2, 8, 12, 32), 2), ((1, 64, 64, 256), 2), ((2, 9, 9, 16), 3)])
def test_maxpool_bwd_two_consumers(cuda, shape, k):
    """seg_maxpool_bwd2 / seg_maxpool_bwd2_y: the pooled tensor feeds two layers (FCN pool3 /
    pool4, /root/reference/models/fcn.py:192-195), dx = ReluGrad(MaxPoolGrad(dy + dy2)).  With
    the forward pool output as the mask source the pool input is not read; same bits."""
    g = _gen(9)
    x = torch.relu(torch.round(bfr(torch.randn(shape, generator=g)) * 2) / 2)
    y_ref, slot_ref = T.max_pool_with_argmax(x, k, k)
    Nb, H, W, C = shape
    Ho, Wo = y_ref.shape[1], y_ref.shape[2]
    x_d = dev_bf16(x)
    y_d = torch.zeros(Nb, Ho, Wo, C, dtype=BF16, device='cuda')
    am = torch.zeros(Nb, Ho, Wo, C, dtype=torch.uint8, device='cuda')
    E.maxpool_fwd(x_d, y_d, am, k, k)
    dy = bfr(torch.randn(y_ref.shape, generator=g))
    dy2 = bfr(torch.randn(y_ref.shape, generator=g))
    xr = x.clone().requires_grad_(True)
    (dx_ref,) = torch.autograd.grad(T.max_pool(xr, k, k), xr, dy + dy2)
    ref = bfr(dx_ref * (x > 0).float())
    a = torch.full(shape, float('nan'), dtype=BF16, device='cuda')
    b = torch.full(shape, float('nan'), dtype=BF16, device='cuda')
    E.maxpool_bwd2(dev_bf16(dy), dev_bf16(dy2), am, a, k, k, mask=x_d)
    E.maxpool_bwd2(dev_bf16(dy), dev_bf16(dy2), am, b, k, k, mask=x_d, pooled=y_d)
    sync()
    # rows / columns beyond the last full window receive no gradient
    assert torch.equal(a.float().cpu(), ref)
    assert torch.equal(b.view(torch.int16).cpu(), a.view(torch.int16).cpu())


def test_softmax_xent_and_head(cuda):
    g = _gen(2)
    for C, cpad in ((2, 16), (21, 32), (5, 16)):
        logits = torch.randn(2, 9, 7, C, generator=g) * 3
        labels = torch.randint(0, C, (2, 13, 11, 1), generator=g).to(torch.uint8)
        lab_crop = T.crop_or_pad(labels, 9, 7)
        lr = logits.clone().requires_grad_(True)
        loss_ref = T.softmax_xent_mean(lr, lab_crop)
        (dl_ref,) = torch.autograd.grad(loss_ref, lr)
        lg_d = logits.cuda()
        lab_d = labels.cuda()
        view = lab_d[:, 2:11, 2:9, :]
        loss_sum = torch.zeros(1, device='cuda')
        dl = torch.full((2, 9, 7, cpad), float('nan'), dtype=BF16, device='cuda')
        E.softmax_xent(lg_d, view, loss_sum, dl)
        probs = torch.zeros(2, 9, 7, C, device='cuda')
        lm = torch.zeros(2, 9, 7, 1, device='cuda')
        E.sigmoid_argmax(lg_d, probs, lm)
        sync()
        loss = float(loss_sum.item()) / (2 * 9 * 7)
        assert abs(loss - float(loss_ref)) < 1e-5 * max(1.0, abs(float(loss_ref)))
        assert rel_l2(dl.float().cpu()[..., :C], dl_ref) < 4e-3         # one bf16 rounding
        assert float(dl.float().cpu()[..., C:].abs().max()) == 0.0
        sig_ref, lab_ref = T.sigmoid_argmax(logits)
        assert torch.allclose(probs.cpu(), sig_ref, atol=2e-7, rtol=1e-6)
        assert torch.equal(lm.cpu(), lab_ref)                           # bit-exact label map


def test_sigmoid_argmax_saturation_ties(cuda):
    """fp32 sigmoid saturates to 1.0 -> argmax(sigmoid) ties resolve to index 0
    (reference applies argmax to y_hat_sig, models/unet.py:76-77)."""
    logits = torch.tensor([[[[20.0, 30.0], [30.0, 20.0], [-1.0, 2.0], [2.0, 2.0]]]])
    probs = torch.zeros(1, 1, 4, 2, device='cuda')
    lm = torch.zeros(1, 1, 4, 1, device='cuda')
    E.sigmoid_argmax(logits.cuda(), probs, lm)
    sync()
    _, lab_ref = T.sigmoid_argmax(logits)
    assert torch.equal(lm.cpu(), lab_ref)
    assert lm.cpu().flatten().tolist() == [0.0, 0.0, 1.0, 0.0]


def test_dropout_mask_bit_exact(cuda):
    g = _gen(9)
    x = bfr(torch.randn(1, 13, 11, 16, generator=g))
    for seed, stream in ((0, 0), (1234567891011, 17), (2 ** 40 + 5, 3)):
        ref = bfr(T.dropout(x, seed, stream))
        x_d = dev_bf16(x)
        y_d = torch.zeros_like(x_d)
        E.dropout(x_d, y_d, seed, stream)
        sync()
        assert torch.equal(y_d.float().cpu(), ref)


def test_bilinear_upsample(cuda):
    g = _gen(4)
    for f, C, H in ((2, 21, 8), (8, 21, 6), (2, 3, 5), (16, 2, 3), (32, 2, 2)):
        x = bfr(torch.randn(2, H, H + 1, C, generator=g))
        add = bfr(torch.randn(2, H * f, (H + 1) * f, C, generator=g))
        xr = x.clone().requires_grad_(True)
        up = T.bilinear_upsample(xr, f)
        ref = up + add
        x_d = dev_bf16(x)
        y32 = torch.zeros(ref.shape, device='cuda')
        E.bilinear_upsample_fwd(x_d, f, y32, add=dev_bf16(add))
        sync()
        assert torch.allclose(y32.cpu(), ref.detach(), atol=1e-5, rtol=1e-5), f
        dy = torch.randn(ref.shape, generator=g)
        (dx_ref,) = torch.autograd.grad(up, xr, dy)
        dx_d = torch.zeros(x.shape, dtype=BF16, device='cuda')
        E.bilinear_upsample_bwd(dy.cuda(), f, dx_d)
        sync()
        assert rel_l2(dx_d.float().cpu(), dx_ref) < 4e-3, f


def test_resize_bilinear(cuda):
    g = _gen(6)
    for (H, W, oh, ow) in ((11, 11, 25, 25), (221 // 4, 50, 128, 128), (16, 16, 8, 8), (7, 9, 7, 9)):
        x = bfr(torch.randn(2, H, W, 8, generator=g))
        xr = x.clone().requires_grad_(True)
        ref = T.resize_bilinear(xr, oh, ow)
        y_d = torch.zeros(2, oh, ow, 8, dtype=BF16, device='cuda')
        E.resize_bilinear_fwd(dev_bf16(x), y_d)
        dy = bfr(torch.randn(ref.shape, generator=g))
        (dx_ref,) = torch.autograd.grad(ref, xr, dy)
        dx_d = torch.zeros(x.shape, dtype=BF16, device='cuda')
        E.resize_bilinear_bwd(dev_bf16(dy), dx_d)
        sync()
        assert rel_l2(y_d.float().cpu(), ref.detach()) < 4e-3
        assert rel_l2(dx_d.float().cpu(), dx_ref) < 4e-3


def test_batchnorm(cuda):
    g = _gen(8)
    for C in (32, 64, 2, 256):
        x = bfr(torch.relu(torch.randn(2, 9, 10, C, generator=g)))
        store = E.ParamStore(torch.device('cuda'))
        bn = E.BatchNorm(store, 'bn', C)
        store.finalize()
        beta = torch.randn(C, generator=g) * 0.1
        bn.beta.value().copy_(beta.cuda())
        xr = x.clone().requires_grad_(True)
        y_ref, m_ref, v_ref = T.batch_norm(xr, beta, torch.zeros(C), torch.ones(C), True)
        cp = E.pad16(C)
        x_d = dev_bf16(x, cp)
        y_d = torch.zeros(2, 9, 10, cp, dtype=BF16, device='cuda')
        bn.forward(x_d, y_d, training=True)
        dy = bfr(torch.randn(y_ref.shape, generator=g))
        (dx_ref,) = torch.autograd.grad(y_ref, xr, dy)
        dx_ref = dx_ref * (x > 0).float()
        dx_d = torch.zeros_like(y_d)
        bn.backward(dev_bf16(dy, cp), x_d, dx_d)
        sync()
        assert rel_l2(y_d.float().cpu()[..., :C], y_ref.detach()) < 4e-3
        assert torch.allclose(bn.moving_mean.cpu(), m_ref, atol=1e-6)
        assert torch.allclose(bn.moving_var.cpu(), v_ref, atol=1e-6)
        assert rel_l2(dx_d.float().cpu()[..., :C], dx_ref) < 6e-3
        assert rel_l2(bn.beta.grad().cpu(), dy.sum(dim=(0, 1, 2))) < 1e-4
        # inference form uses the moving statistics
        y_inf, _, _ = T.batch_norm(x, beta, m_ref, v_ref, False)
        bn.forward(x_d, y_d, training=False)
        sync()
        assert rel_l2(y_d.float().cpu()[..., :C], y_inf) < 4e-3


@pytest.mark.parametrize('shape,k,C', [((2, 12, 15, 64), 3, 64), ((1, 254, 254, 64), 3, 64),
                                       ((3, 10, 8, 32), 2, 32), ((2, 9, 9, 16), 3, 5)])
def test_maxpool_bn_infer_is_bit_identical_to_bn_then_pool(cuda, shape, k, C):
    """seg_maxpool_bn_infer (batch-norm applied to the pooled tensor) against the order the
    model states, seg_batchnorm_infer then seg_maxpool_fwd: the same bits, because the
    normalisation is increasing and every rounding monotonic
    (/root/reference/models/deconvolution.py:126-138)."""
    g = _gen(21)
    Nb, H, W, Cp = shape
    x = torch.zeros(shape)
    x[..., :C] = torch.relu(torch.round(torch.randn(Nb, H, W, C, generator=g) * 4) / 4)
    store = E.ParamStore(torch.device('cuda'))
    bn = E.BatchNorm(store, 'bn', C)
    store.finalize()
    bn.beta.value().copy_((torch.randn(C, generator=g) * 0.2).cuda())
    bn.moving_mean.copy_((torch.rand(C, generator=g) * 0.5).cuda())
    bn.moving_var.copy_((torch.rand(C, generator=g) * 0.5 + 0.01).cuda())
    x_d = dev_bf16(x)
    Ho, Wo = H // k, W // k
    full = torch.zeros(shape, dtype=BF16, device='cuda')
    bn.forward(x_d, full, training=False)
    ref = torch.zeros(Nb, Ho, Wo, Cp, dtype=BF16, device='cuda')
    am = torch.zeros(Nb, Ho, Wo, Cp, dtype=torch.uint8, device='cuda')
    E.maxpool_fwd(full, ref, am, k, k)
    got = torch.full((Nb, Ho, Wo, Cp), float('nan'), dtype=BF16, device='cuda')
    bn.pool_infer(x_d, got, k)
    sync()
    assert torch.equal(got.view(torch.int16).cpu(), ref.view(torch.int16).cpu())
    # and the CPU oracle: pool(bn(x)) in fp32, rounded once
    y_inf, _, _ = T.batch_norm(x[..., :C], bn.beta.value().cpu(), bn.moving_mean.cpu(),
                               bn.moving_var.cpu(), False)
    y_ref = T.max_pool(bfr(y_inf), k, k)
    assert rel_l2(got.float().cpu()[..., :C], y_ref) < 4e-3


@pytest.mark.parametrize('nc,B,hs,ws,rh,rw', [(2, 2, 13, 17, 30, 41), (3, 1, 9, 9, 14, 14),
                                              (4, 2, 11, 7, 29, 15), (2, 1, 55, 55, 128, 128)])
def test_classmap_tail_infer_forms_agree(cuda, nc, B, hs, ws, rh, rw):
    """seg_classmap_tail_infer (resize -> 2x2/s2 transposed conv -> bn -> 3x3 conv -> sigmoid /
    argmax, /root/reference/models/deconvolution.py:163-174): the tensor-core form (option 18)
    against the CUDA-core form and against the five separate C-ABI calls, for 2..4 classes
    and grids that are not multiples of the 14-pixel tile.  All three round to bf16 at the
    same points; they differ by the fp32 summation order of the 32-channel products only."""
    g = _gen(40 + nc)
    store = E.ParamStore(torch.device('cuda'))
    gen = np.random.default_rng(3)
    up = E.ConvLayer(store, 'up', 'deconv', 2, 2, 'VALID', 32, nc, True, gen)
    bn = E.BatchNorm(store, 'bn', nc)
    out = E.ConvLayer(store, 'out', 'conv', 3, 1, 'SAME', nc, nc, False, gen)
    store.finalize()
    up.init_values(); out.init_values()
    up.b.value().copy_((torch.randn(nc, generator=g) * 0.1).cuda())
    out.b.value().copy_((torch.randn(nc, generator=g) * 0.1).cuda())
    bn.beta.value().copy_((torch.randn(nc, generator=g) * 0.2).cuda())
    bn.moving_mean.copy_((torch.rand(nc, generator=g) * 0.3).cuda())
    bn.moving_var.copy_((torch.rand(nc, generator=g) * 0.4 + 0.05).cuda())
    store.refresh_shadow()
    x = dev_bf16(torch.relu(torch.randn(B, hs, ws, 32, generator=g)))
    H, W = 2 * rh, 2 * rw

    def fused(mma):
        N.load().seg_set_option(N.OPT_TAIL_MMA, mma)
        try:
            lg = torch.full((B, H, W, nc), float('nan'), device='cuda')
            pr = torch.full((B, H, W, nc), float('nan'), device='cuda')
            lm = torch.full((B, H, W), float('nan'), device='cuda')
            E.classmap_tail_infer(x, rh, rw, up, bn, out, lg, pr, lm)
            sync()
            name = N.load().seg_last_kernel_name().decode()
        finally:
            N.load().seg_set_option(N.OPT_TAIL_MMA, 1)
        return lg.cpu(), pr.cpu(), lm.cpu(), name

    lg1, pr1, lm1, k1 = fused(1)
    lg0, pr0, lm0, k0 = fused(0)
    assert 'tail_mma' in k1 and 'tail_mma' not in k0, (k1, k0)
    # the separate launches
    cp = E.pad16(nc)
    rs = torch.zeros(B, rh, rw, 32, dtype=BF16, device='cuda')
    E.resize_bilinear_fwd(x, rs)
    dc = torch.zeros(B, H, W, cp, dtype=BF16, device='cuda')
    up.forward(rs, dc[..., :nc])
    nb = torch.zeros_like(dc)
    bn.forward(dc, nb, training=False)
    lg = torch.zeros(B, H, W, nc, device='cuda')
    out.forward(nb, lg, out_f32=True)
    pr = torch.zeros_like(lg)
    lm = torch.zeros(B, H, W, device='cuda')
    E.sigmoid_argmax(lg, pr, lm)
    sync()
    lg, pr, lm = lg.cpu(), pr.cpu(), lm.cpu()
    assert torch.isfinite(lg1).all() and torch.isfinite(pr1).all() and torch.isfinite(lm1).all()
    rec = {'nc': nc, 'mma_vs_cuda_core': rel_l2(lg1, lg0), 'mma_vs_unfused': rel_l2(lg1, lg),
           'cuda_core_vs_unfused': rel_l2(lg0, lg)}
    report('classmap_tail', rec)
    assert rec['mma_vs_cuda_core'] < 2e-3 and rec['mma_vs_unfused'] < 2e-3, rec
    # labels: identical wherever the top-two margin exceeds the summation noise
    dmax = float((lg1 - lg).abs().max())
    top2 = torch.topk(lg, 2, dim=3).values
    safe = (top2[..., 0] - top2[..., 1]) > 4 * dmax + 1e-6
    assert torch.equal(lm1[safe], lm[safe]) and torch.equal(lm1[safe], lm0[safe])
    assert float(safe.float().mean()) > 0.9
    assert float((pr1 - torch.sigmoid(lg1)).abs().max()) < 1e-6


@pytest.mark.parametrize('B,h,w,C,masked', [(2, 8, 8, 21, False), (1, 7, 10, 21, True),
                                            (2, 5, 3, 2, False), (1, 16, 12, 32, True)])
def test_upscore8_xent_is_bit_identical_to_the_three_launches(cuda, B, h, w, C, masked):
    """seg_upscore8_xent_fwd_bwd (FCN-8s training head, /root/reference/models/fcn.py:207-220 +
    models/basemodel.py:59-70) against seg_bilinear_upsample_fwd -> seg_softmax_xent_fwd_bwd ->
    seg_bilinear_upsample_bwd: the same bits for the score-map gradient and the logits, the
    loss to the order of its final sum; grids that are not multiples of the 4-pixel block;
    and against the CPU oracle."""
    g = _gen(60 + C)
    cp = E.pad16(C)
    x = bfr(torch.randn(B, h, w, C, generator=g) * 2)
    lab = torch.randint(0, C, (B, 8 * h, 8 * w, 1), generator=g).to(torch.uint8)
    x_d = dev_bf16(x, cp)
    xv = x_d[..., :C]
    lab_d = lab.cuda()
    mask_d = dev_bf16(torch.randn(B, h, w, C, generator=g), cp)[..., :C] if masked else None
    H, W = 8 * h, 8 * w
    # unfused
    lg = torch.zeros(B, H, W, C, device='cuda')
    E.bilinear_upsample_fwd(xv, 8, lg)
    dl = torch.zeros(B, H, W, cp, dtype=BF16, device='cuda')
    ls = torch.zeros(1, device='cuda')
    E.softmax_xent(lg, lab_d, ls, dl)
    dx = torch.zeros(B, h, w, cp, dtype=BF16, device='cuda')
    E.bilinear_upsample_bwd(dl[..., :C], 8, dx[..., :C], mask=mask_d)
    # fused
    lg2 = torch.full((B, H, W, C), float('nan'), device='cuda')
    ls2 = torch.zeros(1, device='cuda')
    dx2 = torch.zeros(B, h, w, cp, dtype=BF16, device='cuda')
    E.upscore8_xent(xv, lab_d, ls2, dx2[..., :C], mask=mask_d, logits=lg2)
    # and without the logits output
    ls3 = torch.zeros(1, device='cuda')
    dx3 = torch.zeros(B, h, w, cp, dtype=BF16, device='cuda')
    E.upscore8_xent(xv, lab_d, ls3, dx3[..., :C], mask=mask_d)
    sync()
    assert torch.equal(lg2.cpu(), lg.cpu())
    assert torch.equal(dx2.view(torch.int16).cpu(), dx.view(torch.int16).cpu())
    assert torch.equal(dx3.view(torch.int16).cpu(), dx.view(torch.int16).cpu())
    assert float(dx.float().abs().max()) > 0
    assert abs(float(ls2) - float(ls)) <= 1e-5 * abs(float(ls)) + 1e-6
    assert abs(float(ls3) - float(ls)) <= 1e-5 * abs(float(ls)) + 1e-6
    # oracle: loss and gradient of the x8 bilinear transposed conv + mean softmax x-entropy
    xr = x.clone().requires_grad_(True)
    loss_ref = T.softmax_xent_mean(T.bilinear_upsample(xr, 8), lab)
    (dx_ref,) = torch.autograd.grad(loss_ref, xr)
    if masked:
        dx_ref = dx_ref * (mask_d.float().cpu() > 0).float()
    assert abs(float(ls2) / (B * H * W) - float(loss_ref)) < 1e-4
    # the stored dlogits are bf16 (as in the unfused path): relative error ~2^-9 / sqrt(256)
    assert rel_l2(dx2.float().cpu()[..., :C], dx_ref) < 4e-3


def test_adam_matches_tf_formula(cuda):
    g = _gen(10)
    store = E.ParamStore(torch.device('cuda'))
    gen = np.random.default_rng(0)
    lay = E.ConvLayer(store, 'c', 'conv', 3, 1, 'VALID', 3, 20, True, gen)
    lay2 = E.ConvLayer(store, 'd', 'deconv', 2, 2, 'VALID', 20, 5, True, gen)
    store.finalize()
    lay.init_values(); lay2.init_values()
    store.refresh_shadow()
    p = store.master.cpu().clone()
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(p.shape, generator=g) * 0.01
        store.grad.copy_(grad.cuda())
        store.adam_step(1e-3, grad_scale=0.5)
        p, m, v = T.adam_update(p, grad * 0.5, m, v, step, 1e-3)
        sync()
        assert torch.allclose(store.master.cpu(), p, atol=1e-7, rtol=1e-5)
        assert float(store.grad.abs().max()) == 0.0          # zeroed for the next step
    # bf16 shadows follow the masters, with zero channel padding
    w = lay.w.value().cpu(); sh = lay.w.shadow().float().cpu()
    assert torch.equal(sh[:, :, :3, :20], bfr(w))
    assert float(sh[:, :, 3:, :].abs().max()) == 0.0 and float(sh[:, :, :, 20:].abs().max()) == 0.0
    w2 = lay2.w.value().cpu(); sh2 = lay2.w.shadow().float().cpu()
    assert torch.equal(sh2[:, :, :5, :20], bfr(w2))


def test_mc_mean_var_and_pack(cuda):
    g = _gen(12)
    probs = torch.rand(16, 5, 6, 2, generator=g)
    mean = torch.zeros(5, 6, 2, device='cuda'); var = torch.zeros(5, 6, 2, device='cuda')
    E.mc_mean_var(probs.cuda(), mean, var)
    x = torch.rand(2, 6, 7, 3, generator=g)
    y = torch.full((2, 6, 7, 16), float('nan'), dtype=BF16, device='cuda')
    E.pack_input(x.cuda(), y)
    sync()
    assert torch.allclose(mean.cpu(), probs.mean(0), atol=1e-6)
    assert torch.allclose(var.cpu(), probs.var(0, unbiased=False), atol=1e-6)
    assert torch.equal(y.float().cpu()[..., :3], bfr(x))
    assert float(y.float().cpu()[..., 3:].abs().max()) == 0.0


@pytest.mark.parametrize('k,stride,padding', [(3, 1, 'VALID'), (3, 1, 'SAME'), (5, 2, 'SAME')])
def test_pack_patches_bit_exact(cuda, k, stride, padding):
    """seg_pack_patches: channel (r*k+s)*C+c of output pixel (oy,ox) is the bf16 rounding of
    x[oy*stride+r-pad_t, ox*stride+s-pad_l, c] (TF SAME padding: extra pixel bottom/right),
    zero outside the image and in the padding channels."""
    g = _gen(21)
    Nb, H, W, C = 2, 13, 18, 3
    x = torch.rand(Nb, H, W, C, generator=g)
    st = E.ParamStore(torch.device('cuda'))
    lay = E.PatchConvLayer(st, 'c', k, stride, padding, C, 32)
    Ho, Wo = lay.patch_out_hw(H, W)
    y = torch.full((Nb, Ho, Wo, lay.cin_pad), float('nan'), dtype=BF16, device='cuda')
    lay.pack(x.cuda(), y)
    sync()
    pt = E.same_pad(H, k, stride)[0] if padding == 'SAME' else 0
    pl = E.same_pad(W, k, stride)[0] if padding == 'SAME' else 0
    ref = torch.zeros(Nb, Ho, Wo, lay.cin_pad)
    for r in range(k):
        for s in range(k):
            for oy in range(Ho):
                yy = oy * stride + r - pt
                if not 0 <= yy < H:
                    continue
                for ox in range(Wo):
                    xx = ox * stride + s - pl
                    if 0 <= xx < W:
                        ref[:, oy, ox, (r * k + s) * C:(r * k + s + 1) * C] = x[:, yy, xx]
    assert torch.equal(y.float().cpu(), bfr(ref))
