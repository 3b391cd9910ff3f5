"""-m gpu parity tests of the first-layer path: seg_stage_input (fp32 / uint8 -> (R,G,B,1) bf16,
crop, mask, per-step scalars) and the first-layer tcgen05 kernel (csrc/fconv.cuh, forward and
weight gradient) against the CPU oracle, through the C ABI.

Tolerances as in tests/test_gpu_conv.py: bf16 operands and fp32 accumulation on both sides,
rel-L2 <= 4e-3 for bf16 outputs, <= 2e-4 for fp32 outputs; staging is bit-exact.
Reference anchors: models/unet.py:111 / models/fcn.py:110 (the convolution),
utils/datasets.py:176-190 (the /255, the joint crop, uint8(mask/255))."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import tf_ops as T
from segmentation_b200 import engine as E
from segmentation_b200 import native as N

from gpu_util import bfr, conv_pads, desc, rel_l2, report, shadow_conv, sync

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
TOL_BF16, TOL_F32 = 4e-3, 2e-4


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def _x4(x):
    """CPU fp32 [B,H,W,3] -> device bf16 [B,H,W,4] = (R,G,B,1) via seg_stage_input."""
    y4 = torch.zeros(x.shape[0], x.shape[1], x.shape[2], 4, dtype=BF16, device='cuda')
    E.stage_input(x.cuda().contiguous(), y4)
    return y4


@pytest.mark.parametrize('shape', [(3, 37, 29), (2, 16, 18)], ids=['scalar', 'vec4'])
def test_stage_input_fp32_bit_exact(cuda, shape):
    """fp32 -> (R,G,B,1) bf16 and the mask copy: the one-pixel-per-thread path (pixel count
    not a multiple of 4) and the four-pixels-per-thread path."""
    g = _gen(0)
    B, H, W = shape
    x = torch.rand(B, H, W, 3, generator=g)
    m = (torch.rand(B, H, W, 1, generator=g) * 21).to(torch.uint8)
    y4 = torch.zeros(B, H, W, 4, dtype=BF16, device='cuda')
    m_out = torch.full((B, H, W, 1), 99, dtype=torch.uint8, device='cuda')
    E.stage_input(x.cuda(), y4, mask_src=m.cuda(), mask_dst=m_out)
    sync()
    ref = torch.ones(B, H, W, 4, dtype=BF16)
    ref[..., :3] = x.to(BF16)
    assert torch.equal(y4.cpu(), ref)
    assert torch.equal(m_out.cpu(), m)


def test_stage_input_u8_crop_mask_and_scalars(cuda):
    """uint8 image / 255 (fp32 division, then bf16), per-image crop windows, raw 0/255 mask
    -> {0,1}; the ctl scalars: loss published to the pinned ring then zeroed, lr_t, step."""
    rng = np.random.default_rng(1)
    B, Hs, Ws, H, W = 4, 50, 61, 32, 40
    img = rng.integers(0, 256, (B, Hs, Ws, 3), dtype=np.uint8)
    msk = rng.choice(np.array([0, 255, 254, 1], dtype=np.uint8), (B, Hs, Ws, 1))
    crop = np.stack([rng.integers(0, Hs - H + 1, B), rng.integers(0, Ws - W + 1, B)], 1).astype(np.int32)
    y4 = torch.zeros(B, H, W, 4, dtype=BF16, device='cuda')
    m_out = torch.full((B, H, W, 1), 7, dtype=torch.uint8, device='cuda')
    loss = torch.tensor([123.5], device='cuda')
    ring = torch.zeros(8, dtype=torch.float32).pin_memory()
    lr_dev = torch.zeros(1, device='cuda')
    step_dev = torch.zeros(1, dtype=torch.int32, device='cuda')
    ctl = N.SegStageCtl(loss.data_ptr(), ring.data_ptr(), 6, 0.25, lr_dev.data_ptr(), 41,
                        step_dev.data_ptr())
    E.stage_input(torch.from_numpy(img).cuda(), y4, mask_src=torch.from_numpy(msk).cuda(),
                  mask_dst=m_out, crop_yx=torch.from_numpy(crop).cuda(), ctl=ctl)
    sync()
    ref = torch.ones(B, H, W, 4, dtype=BF16)
    mref = torch.zeros(B, H, W, 1, dtype=torch.uint8)
    for n in range(B):
        cy, cx = int(crop[n, 0]), int(crop[n, 1])
        win = img[n, cy:cy + H, cx:cx + W].astype(np.float32) / np.float32(255.0)
        ref[n, :, :, :3] = torch.from_numpy(win).to(BF16)
        mref[n] = torch.from_numpy((msk[n, cy:cy + H, cx:cx + W] == 255).astype(np.uint8))
    assert torch.equal(y4.cpu(), ref)
    assert torch.equal(m_out.cpu(), mref)
    assert float(loss.item()) == 0.0
    assert float(ring[2 * (6 & 3)]) == 123.5 and int(ring.view(torch.int32)[2 * (6 & 3) + 1]) == 6
    assert float(lr_dev.item()) == 0.25 and int(step_dev.item()) == 41


CASES = [
    # name, N, H, W, Cout, padding
    ('valid_small', 2, 30, 26, 32, 'VALID'),
    ('valid_ragged', 3, 41, 37, 32, 'VALID'),        # pixel count not a multiple of 128
    ('same_odd', 1, 15, 17, 32, 'SAME'),
    ('same_64', 2, 24, 40, 64, 'SAME'),
    ('valid_64', 1, 33, 35, 64, 'VALID'),
    ('valid_cout24', 2, 20, 22, 24, 'VALID'),        # logical cout < padded 32
    ('many_tiles', 4, 130, 126, 32, 'VALID'),        # > 2 x 148 tiles: several tiles per CTA
]


def _inputs(case, seed=0):
    name, Nb, H, W, Co, padding = case
    g = _gen(seed)
    x = torch.rand(Nb, H, W, 3, generator=g)
    w = bfr(torch.randn(3, 3, 3, Co, generator=g) * 0.2)
    b = torch.randn(Co, generator=g) * 0.1
    return x, w, b


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_first_layer_forward(cuda, case):
    name, Nb, H, W, Co, padding = case
    x, w, b = _inputs(case)
    cop = 32 if Co <= 32 else 64
    pads = conv_pads(H, W, 3, 1, padding)
    Ho = H + pads[0] + pads[2] - 2
    Wo = W + pads[1] + pads[3] - 2
    x4 = _x4(x)
    sh = shadow_conv(w, 16, cop)
    bd = b.cuda()
    ref = torch.relu(T.conv2d(bfr(x), w, b, 1, padding))
    out = {}
    for impl_name, impl in (('umma', N.IMPL_UMMA), ('simt', N.IMPL_SIMT)):
        y = torch.full((Nb, Ho, Wo, cop), float('nan'), dtype=BF16, device='cuda')
        d = desc(3, 1, pads, 3, Co, 16, cop, N.EPI_BIAS | N.EPI_RELU, impl)
        N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(x4), None, N.ptr(sh), N.ptr(bd), N.vref(y),
               N.stream_ptr())
        sync()
        if impl == N.IMPL_UMMA:
            assert 'fconv' in N.load().seg_last_kernel_name().decode()
        out[impl_name] = y.float().cpu()
        err = rel_l2(out[impl_name][..., :Co], ref)
        report('first_layer_fwd', {'case': name, 'impl': impl_name, 'err': err})
        assert err < TOL_BF16, (name, impl_name, err)
    # padded output channels are exact zeros (bias 0, weights 0, ReLU)
    assert float(out['umma'][..., Co:].abs().max() if Co < cop else 0.0) == 0.0


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_first_layer_wgrad(cuda, case):
    name, Nb, H, W, Co, padding = case
    x, w, b = _inputs(case, seed=3)
    cop = 32 if Co <= 32 else 64
    pads = conv_pads(H, W, 3, 1, padding)
    Ho = H + pads[0] + pads[2] - 2
    Wo = W + pads[1] + pads[3] - 2
    g = _gen(5)
    dz = bfr(torch.randn(Nb, Ho, Wo, Co, generator=g) * (torch.rand(Nb, Ho, Wo, Co, generator=g) > 0.4))
    dz_d = torch.zeros(Nb, Ho, Wo, cop, dtype=BF16, device='cuda')
    dz_d[..., :Co] = dz.to(BF16).cuda()
    x4 = _x4(x)
    # oracle: dW = correlation of the bf16-rounded input with dz; db = sum dz
    xr = bfr(x).double()
    xp = torch.nn.functional.pad(xr, (0, 0, pads[1], pads[3], pads[0], pads[2]))
    dw_ref = torch.zeros(3, 3, 3, Co, dtype=torch.float64)
    for r in range(3):
        for s in range(3):
            win = xp[:, r:r + Ho, s:s + Wo, :]
            dw_ref[r, s] = torch.einsum('nhwc,nhwo->co', win, dz.double())
    db_ref = dz.double().sum((0, 1, 2))
    for impl_name, impl in (('umma', N.IMPL_UMMA), ('simt', N.IMPL_SIMT)):
        dw = torch.zeros(3, 3, 3, Co, dtype=torch.float32, device='cuda')
        db = torch.zeros(Co, dtype=torch.float32, device='cuda')
        d = desc(3, 1, pads, 3, Co, 16, cop, 0, impl)
        N.call('seg_conv2d_wgrad', ctypes.byref(d), N.vref(x4), None, N.vref(dz_d), N.ptr(dw),
               N.ptr(db), N.stream_ptr())
        sync()
        if impl == N.IMPL_UMMA:
            assert 'fconv' in N.load().seg_last_kernel_name().decode()
        e_w = rel_l2(dw.cpu(), dw_ref)
        e_b = rel_l2(db.cpu(), db_ref)
        report('first_layer_wgrad', {'case': name, 'impl': impl_name, 'dw': e_w, 'db': e_b})
        assert e_w < TOL_F32 and e_b < TOL_F32, (name, impl_name, e_w, e_b)


# ---------------------------------------------------------------------------
# first layer fused with its 2x2 max-pool (seg_conv2d_pool_fwd / seg_conv2d_pool_wgrad)
# ---------------------------------------------------------------------------
POOL_CASES = [
    # name, N, H, W, Cout, padding, window (y0, x0, h, w) in output coordinates or None
    ('unet_like', 2, 44, 70, 32, 'VALID', (9, 13, 20, 30)),
    ('wide', 1, 20, 200, 32, 'VALID', (0, 0, 18, 198)),          # window = whole output
    ('fcn_like', 2, 32, 48, 32, 'SAME', None),
    ('cout24', 1, 18, 134, 24, 'VALID', (3, 60, 8, 70)),
]


def _pool_setup(case, seed):
    name, Nb, H, W, Co, padding, win = case
    x, w, b = _inputs((name, Nb, H, W, Co, padding), seed=seed)
    pads = conv_pads(H, W, 3, 1, padding)
    Ho, Wo = H + pads[0] + pads[2] - 2, W + pads[1] + pads[3] - 2
    return x, w, b, pads, Ho, Wo


@pytest.mark.parametrize('case', POOL_CASES, ids=[c[0] for c in POOL_CASES])
def test_first_layer_pool_forward(cuda, case):
    """pooled values and argmax slots bit-exact against the unfused C-ABI path (same MMA per
    pixel, then seg_maxpool_fwd), the window of the activation bit-exact as well; pooled
    values against the oracle within the bf16 tolerance."""
    name, Nb, H, W, Co, padding, win = case
    x, w, b, pads, Ho, Wo = _pool_setup(case, 7)
    x4 = _x4(x)
    sh, bd = shadow_conv(w, 16, 32), b.cuda()
    d = desc(3, 1, pads, 3, Co, 16, 32, N.EPI_BIAS | N.EPI_RELU, N.IMPL_UMMA)
    st = N.stream_ptr()
    y_ref = torch.zeros(Nb, Ho, Wo, 32, dtype=BF16, device='cuda')
    N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(x4), None, N.ptr(sh), N.ptr(bd), N.vref(y_ref), st)
    p_ref = torch.zeros(Nb, Ho // 2, Wo // 2, 32, dtype=BF16, device='cuda')
    a_ref = torch.zeros(Nb, Ho // 2, Wo // 2, 32, dtype=torch.uint8, device='cuda')
    E.maxpool_fwd(y_ref, p_ref, a_ref)
    y = torch.full((Nb, Ho, Wo, 32), -7.0, dtype=BF16, device='cuda')
    pooled = torch.full_like(p_ref, float('nan'))
    amax = torch.full_like(a_ref, 9)
    ywin = None if win is None else y[:, win[0]:win[0] + win[2], win[1]:win[1] + win[3], :]
    N.call('seg_conv2d_pool_fwd', ctypes.byref(d), N.vref(x4), N.ptr(sh), N.ptr(bd), N.vref(ywin),
           win[0] if win else 0, win[1] if win else 0, N.vref(pooled), N.ptr(amax), st)
    sync()
    assert 'fconv' in N.load().seg_last_kernel_name().decode()
    assert torch.equal(pooled.cpu().view(torch.int16), p_ref.cpu().view(torch.int16))
    assert torch.equal(amax.cpu(), a_ref.cpu())
    if win is not None:
        y0, x0, h, ww = win
        assert torch.equal(y[:, y0:y0 + h, x0:x0 + ww].cpu().view(torch.int16),
                           y_ref[:, y0:y0 + h, x0:x0 + ww].cpu().view(torch.int16))
        outside = y.clone()
        outside[:, y0:y0 + h, x0:x0 + ww] = -7.0
        assert bool((outside == -7.0).all())                 # nothing written outside the window
    ref = torch.relu(T.conv2d(bfr(x), w, b, 1, padding))
    pr = T.max_pool(bfr(ref), 2, 2)
    assert rel_l2(pooled.float().cpu()[..., :Co], pr) < TOL_BF16


@pytest.mark.parametrize('case', POOL_CASES, ids=[c[0] for c in POOL_CASES])
def test_first_layer_pool_wgrad(cuda, case):
    """dW / db of the fused pool-backward + weight-gradient launch against the unfused C-ABI
    path (seg_maxpool_bwd_y, then seg_conv2d_wgrad on the materialised gradient)."""
    name, Nb, H, W, Co, padding, win = case
    x, w, b, pads, Ho, Wo = _pool_setup(case, 11)
    x4 = _x4(x)
    sh, bd = shadow_conv(w, 16, 32), b.cuda()
    d = desc(3, 1, pads, 3, Co, 16, 32, N.EPI_BIAS | N.EPI_RELU, N.IMPL_UMMA)
    st = N.stream_ptr()
    y = torch.zeros(Nb, Ho, Wo, 32, dtype=BF16, device='cuda')
    N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(x4), None, N.ptr(sh), N.ptr(bd), N.vref(y), st)
    pooled = torch.zeros(Nb, Ho // 2, Wo // 2, 32, dtype=BF16, device='cuda')
    amax = torch.zeros(Nb, Ho // 2, Wo // 2, 32, dtype=torch.uint8, device='cuda')
    E.maxpool_fwd(y, pooled, amax)
    g = _gen(13)
    dpool = torch.zeros_like(pooled)
    dpool[..., :Co] = (torch.randn(Nb, Ho // 2, Wo // 2, Co, generator=g) * 0.3).to(BF16).cuda()
    # unfused: materialise dz, then the first-layer weight gradient
    dz = torch.zeros_like(y)
    E.maxpool_bwd(dpool, amax, dz, mask=y, pooled=pooled)
    dw_ref = torch.zeros(3, 3, 3, Co, dtype=torch.float32, device='cuda')
    db_ref = torch.zeros(Co, dtype=torch.float32, device='cuda')
    d0 = desc(3, 1, pads, 3, Co, 16, 32, 0, N.IMPL_UMMA)
    N.call('seg_conv2d_wgrad', ctypes.byref(d0), N.vref(x4), None, N.vref(dz), N.ptr(dw_ref),
           N.ptr(db_ref), st)
    dw = torch.zeros_like(dw_ref)
    db = torch.zeros_like(db_ref)
    N.call('seg_conv2d_pool_wgrad', ctypes.byref(d0), N.vref(x4), N.vref(dpool), N.ptr(amax),
           N.vref(pooled), N.ptr(dw), N.ptr(db), st)
    sync()
    assert 'fconv' in N.load().seg_last_kernel_name().decode()
    e_w, e_b = rel_l2(dw.cpu(), dw_ref.cpu()), rel_l2(db.cpu(), db_ref.cpu())
    report('first_layer_pool_wgrad', {'case': name, 'dw': e_w, 'db': e_b})
    assert e_w < 2e-5 and e_b < 2e-5, (name, e_w, e_b)


def test_first_layer_wgrad_on_crop_view(cuda):
    """The plain first-layer weight gradient on a crop VIEW of the staged input (U-Net: the
    window of conv1_1 whose gradient arrives from conv1_2, models/unet.py:118-120): equals
    the gradient computed on a dense copy of the crop."""
    g = _gen(21)
    Nb, H, W, Co = 2, 40, 52, 32
    x = torch.rand(Nb, H, W, 3, generator=g)
    x4 = _x4(x)
    y0, x0, h, w = 7, 11, 22, 30                          # input window; output (h-2) x (w-2)
    dz = (torch.randn(Nb, h - 2, w - 2, Co, generator=g) * 0.2).to(BF16).cuda()
    d0 = desc(3, 1, (0, 0, 0, 0), 3, Co, 16, 32, 0, N.IMPL_UMMA)
    out = []
    for src in (x4[:, y0:y0 + h, x0:x0 + w, :], x4[:, y0:y0 + h, x0:x0 + w, :].contiguous()):
        dw = torch.zeros(3, 3, 3, Co, dtype=torch.float32, device='cuda')
        db = torch.zeros(Co, dtype=torch.float32, device='cuda')
        N.call('seg_conv2d_wgrad', ctypes.byref(d0), N.vref(src), None, N.vref(dz), N.ptr(dw),
               N.ptr(db), N.stream_ptr())
        sync()
        assert 'fconv' in N.load().seg_last_kernel_name().decode()
        out.append((dw.cpu(), db.cpu()))
    assert rel_l2(out[0][0], out[1][0]) < 1e-6 and rel_l2(out[0][1], out[1][1]) < 1e-6
    xr = bfr(x)[:, y0:y0 + h, x0:x0 + w].double()
    dw_ref = torch.zeros(3, 3, 3, Co, dtype=torch.float64)
    for r in range(3):
        for s in range(3):
            dw_ref[r, s] = torch.einsum('nhwc,nhwo->co', xr[:, r:r + h - 2, s:s + w - 2, :],
                                        dz.float().cpu().double())
    assert rel_l2(out[0][0], dw_ref) < TOL_F32


@pytest.mark.parametrize('geom', [(5, 2, 'SAME', 2, 64, 96), (3, 1, 'VALID', 2, 38, 70)],
                         ids=['k5s2_same', 'k3s1_valid'])
def test_first_layer_bn_pool_infer(cuda, geom):
    """conv (5x5/s2 SAME as DeconvModel's conv1_0, or 3x3/s1) + ReLU + batch-norm (moving
    statistics) + 2x2 max-pool in one launch against the oracle (reference
    models/deconvolution.py:109-118), both weight layouts (padded HWIO shadow, dense
    patch-packed matrix)."""
    k, s, padding, Nb, H, W = geom
    Co = 32
    g = _gen(31)
    x = torch.rand(Nb, H, W, 3, generator=g)
    w = bfr(torch.randn(k, k, 3, Co, generator=g) * 0.1)
    b = torch.randn(Co, generator=g) * 0.1
    mm = torch.rand(Co, generator=g) * 0.3
    mv = torch.rand(Co, generator=g) * 0.3 + 0.05
    beta = torch.randn(Co, generator=g) * 0.1
    eps = 1e-3
    pads = conv_pads(H, W, k, s, padding)
    Ho = (H + pads[0] + pads[2] - k) // s + 1
    Wo = (W + pads[1] + pads[3] - k) // s + 1
    x4 = _x4(x)
    conv = bfr(torch.relu(T.conv2d(bfr(x), w, b, s, padding)))
    bn = bfr((conv - mm) * torch.rsqrt(mv + eps) + beta)
    ref = T.max_pool(bn, 2, 2)
    d = desc(k, s, pads, 3, Co, 16, 32, N.EPI_BIAS | N.EPI_RELU, N.IMPL_UMMA)
    dense = torch.zeros(k * k * 3 + 5, 32, dtype=BF16)          # patch-packed [k*k*3 (+pad)][32]
    dense[:k * k * 3, :Co] = w.reshape(k * k * 3, Co).to(BF16)
    bd, mmd, mvd, betad = b.cuda(), mm.cuda(), mv.cuda(), beta.cuda()   # keep the buffers alive
    for name, sh, rows in (('shadow', shadow_conv(w, 16, 32), 0), ('dense', dense.cuda(), 3)):
        pooled = torch.full((Nb, Ho // 2, Wo // 2, 32), float('nan'), dtype=BF16, device='cuda')
        N.call('seg_conv2d_bn_pool_infer', ctypes.byref(d), N.vref(x4), N.ptr(sh), rows,
               N.ptr(bd), N.ptr(mmd), N.ptr(mvd), eps, N.ptr(betad),
               N.vref(pooled), None, N.stream_ptr())
        sync()
        assert 'fconv' in N.load().seg_last_kernel_name().decode()
        e = rel_l2(pooled.float().cpu()[..., :Co], ref)
        report('first_layer_bn_pool', {'geom': list(geom), 'weights': name, 'err': e})
        assert e < TOL_BF16, (geom, name, e)
