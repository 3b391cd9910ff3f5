"""-m gpu parity tests for the convolution family: tcgen05 implicit-GEMM (UMMA)
and CUDA-core (SIMT) paths against the CPU oracle, through the C ABI.

Tolerance: operands are rounded to bf16 on both sides and accumulated in fp32,
so the only difference is summation order -> rel-L2 <= 2e-3 for bf16 outputs
(one bf16 rounding of the result, 2^-9 relative), 1e-4 for fp32 outputs.
"""
import ctypes
import os

import pytest
import torch

from oracle import tf_ops as T
from segmentation_b200 import native as N

from gpu_util import (bfr, conv_pads, desc, dev_bf16, pad16, rel_l2, report, shadow_conv,
                      shadow_deconv, sync)

pytestmark = pytest.mark.gpu
IMPLS = [('simt', N.IMPL_SIMT), ('umma', N.IMPL_UMMA), ('umma_halo', N.IMPL_UMMA),
         ('umma_im2col', N.IMPL_UMMA)]


@pytest.fixture(autouse=True)
def _halo_switch(request):
    """'umma' = tcgen05 with the spatial-tile conv kernel wherever it structurally applies
    (efficiency gate off, so the small test shapes exercise it); 'umma_halo' = the
    position-space halo kernel instead; 'umma_im2col' forces the TMA-im2col kernel."""
    name = request.node.callspec.params.get('impl_name') if hasattr(request.node, 'callspec') else None
    if name == 'umma':
        N.set_option(N.OPT_TILE_CONV_MIN_EFF, 0)
        N.set_option(N.OPT_TILE_WGRAD_MIN_EFF, 0)
    elif name == 'umma_halo':
        N.set_option(N.OPT_TILE_CONV, 0)
        N.set_option(N.OPT_TILE_WGRAD, 0)      # TMA-im2col weight-gradient kernel
    elif name == 'umma_im2col':
        N.set_option(N.OPT_TILE_CONV, 0)
        N.set_option(N.OPT_HALO_CONV, 0)
        N.set_option(N.OPT_TILE_WGRAD, 0)
    yield
    N.set_option(N.OPT_HALO_CONV, 1)
    N.set_option(N.OPT_TILE_CONV, 1)
    N.set_option(N.OPT_TILE_CONV_MIN_EFF, 70)
    N.set_option(N.OPT_TILE_WGRAD, 1)
    N.set_option(N.OPT_TILE_WGRAD_MIN_EFF, 40)
TOL_BF16 = 4e-3
TOL_F32 = 2e-4


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


# ---------------------------------------------------------------------------
# descriptor probes: plain GEMMs through the igemm kernel (no im2col)
# ---------------------------------------------------------------------------
PROBE_SHAPES = [
    # M, N, K           (K picks KC: 64 | 32 | 16; N picks BN / swizzle of MN-major B)
    (128, 64, 64), (128, 128, 128), (384, 256, 192), (200, 128, 64),
    (128, 32, 32), (256, 64, 96), (130, 32, 160),
    (128, 16, 16), (256, 16, 48), (128, 32, 16), (128, 64, 16), (192, 128, 48),
    (128, 16, 64), (128, 256, 32),
]


@pytest.mark.parametrize('mode', [0, 1])
def test_probe_umma_gemm(cuda, mode):
    g = _gen(1)
    fails = []
    for (M, Nn, K) in PROBE_SHAPES:
        a = bfr(torch.randn(M, K, generator=g))
        b = bfr(torch.randn(Nn, K, generator=g))           # [N][K]
        ref = a @ b.t()
        a_d = a.to(torch.bfloat16).cuda()
        b_d = (b.t().contiguous() if mode == 1 else b).to(torch.bfloat16).cuda()
        d = torch.full((M, Nn), float('nan'), dtype=torch.float32, device='cuda')
        N.call_probe('seg_probe_umma', mode, M, Nn, K, N.ptr(a_d), N.ptr(b_d), N.ptr(d),
                     N.stream_ptr())
        sync()
        err = rel_l2(d.cpu(), ref)
        report('probe', {'mode': mode, 'M': M, 'N': Nn, 'K': K, 'err': err})
        if not (err < TOL_F32):
            fails.append((M, Nn, K, err))
    assert not fails, 'probe failures (M,N,K,err): %s' % fails


# ---------------------------------------------------------------------------
# conv forward
# ---------------------------------------------------------------------------
CONV_CASES = [
    # name, N, H, W, C1, C2, Cout, k, stride, padding, relu, out_f32
    ('v3_32_64', 2, 20, 18, 32, 0, 64, 3, 1, 'VALID', True, False),
    ('v3_64_128', 1, 17, 23, 64, 0, 128, 3, 1, 'VALID', True, False),
    ('v3_128_256', 1, 12, 12, 128, 0, 256, 3, 1, 'VALID', True, False),
    ('v3_rgb', 2, 30, 26, 3, 0, 32, 3, 1, 'VALID', True, False),
    ('s3_32_32', 2, 16, 16, 32, 0, 32, 3, 1, 'SAME', True, False),
    ('s3_rgb_odd', 1, 15, 17, 3, 0, 32, 3, 1, 'SAME', True, False),
    ('concat', 2, 14, 14, 64, 64, 64, 3, 1, 'VALID', True, False),
    ('concat32', 1, 20, 20, 32, 32, 32, 3, 1, 'VALID', True, False),
    ('head1x1', 2, 12, 12, 32, 0, 2, 1, 1, 'VALID', False, True),
    ('p1x1_21', 1, 16, 16, 128, 0, 21, 1, 1, 'SAME', True, False),
    ('s5s2_rgb', 1, 32, 32, 3, 0, 32, 5, 2, 'SAME', True, False),
    ('bigM', 3, 40, 40, 32, 0, 32, 3, 1, 'VALID', True, False),
    ('c512', 1, 10, 10, 256, 0, 512, 3, 1, 'VALID', True, False),
    ('head3x3_2_2', 2, 21, 19, 2, 0, 2, 3, 1, 'SAME', False, True),      # streaming class-map head
    ('tiny_4_3', 1, 16, 18, 4, 0, 3, 3, 1, 'SAME', True, False),
]


def _conv_inputs(case, seed=0):
    name, Nb, H, W, C1, C2, Co, k, s, padding, relu, f32 = case
    g = _gen(seed)
    x = bfr(torch.rand(Nb, H, W, C1 + C2, generator=g) - 0.3)
    w = bfr(torch.randn(k, k, C1 + C2, Co, generator=g) * 0.1)
    b = torch.randn(Co, generator=g) * 0.1
    return x, w, b


@pytest.mark.parametrize('impl_name,impl', IMPLS)
@pytest.mark.parametrize('case', CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_fwd(cuda, case, impl_name, impl):
    name, Nb, H, W, C1, C2, Co, k, s, padding, relu, f32 = case
    x, w, b = _conv_inputs(case)
    ref = T.conv2d(x, w, b, s, padding)
    if relu:
        ref = torch.relu(ref)
    cin, cin_pad, cout_pad = C1 + C2, pad16(C1 + C2), pad16(Co)
    if C2:
        x1_d, x2_d = dev_bf16(x[..., :C1]), dev_bf16(x[..., C1:])
    else:
        x1_d, x2_d = dev_bf16(x, cin_pad), None
    w_d = shadow_conv(w, cin_pad, cout_pad)
    b_d = b.cuda()
    Ho, Wo = ref.shape[1], ref.shape[2]
    y_d = torch.full((Nb, Ho, Wo, Co), float('nan'),
                     dtype=torch.float32 if f32 else torch.bfloat16, device='cuda')
    flags = N.EPI_BIAS | (N.EPI_RELU if relu else 0) | (N.EPI_OUT_F32 if f32 else 0)
    d = desc(k, s, conv_pads(H, W, k, s, padding), cin, Co, cin_pad, cout_pad, flags, impl)
    N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(x1_d), N.vref(x2_d), N.ptr(w_d), N.ptr(b_d),
           N.vref(y_d), N.stream_ptr())
    sync()
    err = rel_l2(y_d.float().cpu(), ref)
    report('conv_fwd', {'case': name, 'impl': impl_name, 'err': err})
    assert err < (TOL_F32 if f32 else TOL_BF16), (name, impl_name, err)


CLUSTER_CASES = [c for c in CONV_CASES if c[0] in ('c512', 'v3_128_256', 'concat')]


@pytest.mark.parametrize('opt', ['rowstage'])
@pytest.mark.parametrize('case', CLUSTER_CASES, ids=[c[0] for c in CLUSTER_CASES])
def test_conv_fwd_halo_plan_options(cuda, case, opt):
    """Plans of the halo kernel for streamed weights: seg_set_option key 14 (one filter row =
    three taps per weight stage, the default) against one tap per stage.  Same numbers;
    layers whose weights stay resident in shared memory ignore the option."""
    key = N.OPT_HALO_ROWSTAGE
    name, Nb, H, W, C1, C2, Co, k, s, padding, relu, f32 = case
    x, w, b = _conv_inputs(case)
    cin, cin_pad, cout_pad = C1 + C2, pad16(C1 + C2), pad16(Co)
    if C2:
        x1_d, x2_d = dev_bf16(x[..., :C1]), dev_bf16(x[..., C1:])
    else:
        x1_d, x2_d = dev_bf16(x, cin_pad), None
    w_d = shadow_conv(w, cin_pad, cout_pad)
    b_d = b.cuda()
    Ho, Wo = (H - k + 1, W - k + 1) if padding == 'VALID' else (H, W)
    flags = N.EPI_BIAS | (N.EPI_RELU if relu else 0)
    d = desc(k, s, conv_pads(H, W, k, s, padding), cin, Co, cin_pad, cout_pad, flags, N.IMPL_UMMA)
    outs = []
    try:
        for on in (0, 1):
            N.set_option(key, on)
            y_d = torch.full((Nb, Ho, Wo, Co), float('nan'), dtype=torch.bfloat16, device='cuda')
            N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(x1_d), N.vref(x2_d), N.ptr(w_d),
                   N.ptr(b_d), N.vref(y_d), N.stream_ptr())
            sync()
            outs.append(y_d.float().cpu())
    finally:
        N.set_option(key, 1)
    ref = torch.relu(T.conv2d(x, w, b, s, padding)) if relu else T.conv2d(x, w, b, s, padding)
    assert rel_l2(outs[1], ref) < TOL_BF16, name
    # same K order per output: the plans differ in staging only
    assert rel_l2(outs[1], outs[0]) < 1e-3, name


# ---------------------------------------------------------------------------
# conv dgrad / wgrad / bias grad
# ---------------------------------------------------------------------------
BWD_CASES = [c for c in CONV_CASES if c[0] not in ('s5s2_rgb',)]


@pytest.mark.parametrize('impl_name,impl', IMPLS)
@pytest.mark.parametrize('case', BWD_CASES, ids=[c[0] for c in BWD_CASES])
def test_conv_bwd(cuda, case, impl_name, impl):
    name, Nb, H, W, C1, C2, Co, k, s, padding, relu, f32 = case
    x, w, b = _conv_inputs(case)
    g = _gen(7)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    z = T.conv2d(xr, wr, None, s, padding)
    dz = bfr(torch.randn(z.shape, generator=g) * 0.05)
    dx_ref, dw_ref = torch.autograd.grad(z, [xr, wr], dz)
    db_ref = dz.sum(dim=(0, 1, 2))
    # the producer of x is a ReLU layer: mask where x <= 0
    mask = (x > 0).float()
    dx_ref = dx_ref * mask

    cin, cin_pad, cout_pad = C1 + C2, pad16(C1 + C2), pad16(Co)
    dz_d = dev_bf16(dz, cout_pad)
    w_d = shadow_conv(w, cin_pad, cout_pad)
    d = desc(k, s, conv_pads(H, W, k, s, padding), cin, Co, cin_pad, cout_pad, 0, impl)
    if C2:
        x1_d, x2_d = dev_bf16(x[..., :C1]), dev_bf16(x[..., C1:])
        dx1 = torch.full((Nb, H, W, C1), float('nan'), dtype=torch.bfloat16, device='cuda')
        dx2 = torch.full((Nb, H, W, C2), float('nan'), dtype=torch.bfloat16, device='cuda')
    else:
        x1_d, x2_d = dev_bf16(x, cin_pad), None
        dx1 = torch.full((Nb, H, W, cin_pad), float('nan'), dtype=torch.bfloat16, device='cuda')
        dx2 = None
    dw_d = torch.zeros(k, k, cin, Co, dtype=torch.float32, device='cuda')
    db_d = torch.zeros(Co, dtype=torch.float32, device='cuda')
    st = N.stream_ptr()
    N.call('seg_conv2d_wgrad', ctypes.byref(d), N.vref(x1_d), N.vref(x2_d), N.vref(dz_d),
           N.ptr(dw_d), N.ptr(db_d), st)                 # BiasAddGrad fused
    do_dgrad = cin >= 16          # the RGB layer never needs dgrad
    if do_dgrad:
        N.call('seg_conv2d_dgrad', ctypes.byref(d), N.vref(dz_d), N.ptr(w_d), N.vref(dx1),
               N.vref(dx2), N.vref(x1_d), N.vref(x2_d), st)
    sync()
    e_w = rel_l2(dw_d.cpu(), dw_ref)
    e_b = rel_l2(db_d.cpu(), db_ref)
    rec = {'case': name, 'impl': impl_name, 'dw': e_w, 'db': e_b}
    ok = e_w < TOL_F32 * 5 and e_b < TOL_F32 * 5
    if do_dgrad:
        got = torch.cat([dx1.float().cpu()[..., :C1 if C2 else cin]] +
                        ([dx2.float().cpu()] if C2 else []), dim=-1)
        e_x = rel_l2(got, dx_ref)
        rec['dx'] = e_x
        ok = ok and e_x < TOL_BF16
    report('conv_bwd', rec)
    assert ok, rec
    if do_dgrad and C2 and impl == N.IMPL_UMMA and s == 1:
        # the two halves of the virtual concat as separate channel-slice launches
        # (seg_conv2d_dgrad_slice) are the same numbers as the two-destination launch
        s1 = torch.full_like(dx1, float('nan'))
        s2 = torch.full_like(dx2, float('nan'))
        N.call('seg_conv2d_dgrad_slice', ctypes.byref(d), N.vref(dz_d), N.ptr(w_d), 0,
               N.vref(s1), N.vref(x1_d), st)
        N.call('seg_conv2d_dgrad_slice', ctypes.byref(d), N.vref(dz_d), N.ptr(w_d), C1,
               N.vref(s2), N.vref(x2_d), st)
        sync()
        # (same K order per output channel; a different tile plan may round a few values
        # differently in bf16, hence a tolerance well below TOL_BF16 instead of equality)
        e1 = rel_l2(s1.float().cpu(), dx1.float().cpu())
        e2 = rel_l2(s2.float().cpu(), dx2.float().cpu())
        assert e1 < 1e-3 and e2 < 1e-3, (name, e1, e2)


@pytest.mark.parametrize('case', [c for c in BWD_CASES if c[7] == 3 and c[8] == 1 and c[6] % 32 == 0],
                         ids=lambda c: c[0])
def test_conv_wgrad_tensor_reduce_option(cuda, case):
    """seg_set_option key 15: the spatial-tile weight-gradient kernel's TMA tensor reduce-add
    epilogue (the default) gives the same dW / db as the row-wise bulk reduce-add epilogue."""
    name, Nb, H, W, C1, C2, Co, k, s, padding, relu, f32 = case
    x, w, b = _conv_inputs(case)
    g = _gen(7)
    z = T.conv2d(x, w, None, s, padding)
    dz = bfr(torch.randn(z.shape, generator=g) * 0.05)
    cin, cin_pad, cout_pad = C1 + C2, pad16(C1 + C2), pad16(Co)
    dz_d = dev_bf16(dz, cout_pad)
    d = desc(k, s, conv_pads(H, W, k, s, padding), cin, Co, cin_pad, cout_pad, 0, N.IMPL_UMMA)
    if C2:
        x1_d, x2_d = dev_bf16(x[..., :C1]), dev_bf16(x[..., C1:])
    else:
        x1_d, x2_d = dev_bf16(x, cin_pad), None
    outs = []
    try:
        for on in (0, 1):
            N.set_option(N.OPT_WGRAD_TENSOR_RED, on)
            dw_d = torch.zeros(k, k, cin, Co, dtype=torch.float32, device='cuda')
            db_d = torch.zeros(Co, dtype=torch.float32, device='cuda')
            N.call('seg_conv2d_wgrad', ctypes.byref(d), N.vref(x1_d), N.vref(x2_d), N.vref(dz_d),
                   N.ptr(dw_d), N.ptr(db_d), N.stream_ptr())
            sync()
            outs.append((dw_d.cpu(), db_d.cpu()))
    finally:
        N.set_option(N.OPT_WGRAD_TENSOR_RED, 1)
    assert rel_l2(outs[1][0], outs[0][0]) < TOL_F32 * 5, name
    assert rel_l2(outs[1][1], outs[0][1]) < TOL_F32 * 5, name


# ---------------------------------------------------------------------------
# transposed conv (slim.convolution2d_transpose)
# ---------------------------------------------------------------------------
DECONV_CASES = [
    # name, N, H, W, Cin, Cout, k, stride, padding, impls
    ('up2_64_32', 2, 8, 8, 64, 32, 2, 2, 'VALID', ('simt', 'umma')),
    ('up2_512_256', 1, 8, 8, 512, 256, 2, 2, 'VALID', ('simt', 'umma')),
    ('up2_32_2', 1, 10, 12, 32, 2, 2, 2, 'VALID', ('simt', 'umma')),
    ('up2_odd', 3, 7, 9, 128, 64, 2, 2, 'VALID', ('simt', 'umma')),
    ('k5s2', 1, 11, 11, 64, 32, 5, 2, 'VALID', ('simt', 'umma')),
    ('k5s2_256_64', 2, 13, 9, 256, 64, 5, 2, 'VALID', ('simt', 'umma')),
    ('k5s2_32_32', 2, 27, 25, 32, 32, 5, 2, 'VALID', ('umma',)),
    ('k3s2', 1, 10, 12, 64, 64, 3, 2, 'VALID', ('simt', 'umma')),
    ('k4s2_same', 1, 9, 9, 32, 16, 4, 2, 'SAME', ('simt',)),
]


@pytest.mark.parametrize('case', DECONV_CASES, ids=[c[0] for c in DECONV_CASES])
def test_deconv_fwd_bwd(cuda, case):
    name, Nb, H, W, Ci, Co, k, s, padding, impls = case
    g = _gen(3)
    x = bfr(torch.rand(Nb, H, W, Ci, generator=g) - 0.3)
    w = bfr(torch.randn(k, k, Co, Ci, generator=g) * 0.1)
    b = torch.randn(Co, generator=g) * 0.1
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    z = T.conv2d_transpose(xr, wr, b, s, padding)
    y_ref = torch.relu(z)
    dz = bfr(torch.randn(z.shape, generator=g) * 0.05)
    dx_ref, dw_ref = torch.autograd.grad(z, [xr, wr], dz)
    dx_ref = dx_ref * (x > 0).float()
    OH, OW = z.shape[1], z.shape[2]
    cin_pad, cout_pad = pad16(Ci), pad16(Co)
    tot = max(k - s, 0) if padding == 'SAME' else 0
    pads = (tot // 2, tot // 2, tot - tot // 2, tot - tot // 2)
    for impl_name in impls:
        impl = dict(IMPLS)[impl_name]
        x_d = dev_bf16(x, cin_pad)
        w_d = shadow_deconv(w, cin_pad, cout_pad)
        b_d = b.cuda()
        y_d = torch.full((Nb, OH, OW, Co), float('nan'), dtype=torch.bfloat16, device='cuda')
        d = desc(k, s, pads, Ci, Co, cin_pad, cout_pad, N.EPI_BIAS | N.EPI_RELU, impl)
        st = N.stream_ptr()
        N.call('seg_deconv2d_fwd', ctypes.byref(d), N.vref(x_d), N.ptr(w_d), N.ptr(b_d),
               N.vref(y_d), st)
        dz_d = dev_bf16(dz, cout_pad)
        dx_d = torch.full((Nb, H, W, cin_pad), float('nan'), dtype=torch.bfloat16, device='cuda')
        dw_d = torch.zeros(k, k, Co, Ci, dtype=torch.float32, device='cuda')
        N.call('seg_deconv2d_dgrad', ctypes.byref(d), N.vref(dz_d), N.ptr(w_d), N.vref(dx_d),
               N.vref(x_d), st)
        N.call('seg_deconv2d_wgrad', ctypes.byref(d), N.vref(x_d), N.vref(dz_d), N.ptr(dw_d), st)
        sync()
        rec = {'case': name, 'impl': impl_name,
               'y': rel_l2(y_d.float().cpu(), y_ref),
               'dx': rel_l2(dx_d.float().cpu()[..., :Ci], dx_ref),
               'dw': rel_l2(dw_d.cpu(), dw_ref)}
        report('deconv', rec)
        assert rec['y'] < TOL_BF16 and rec['dx'] < TOL_BF16 and rec['dw'] < TOL_F32 * 5, rec


AFFINE_CASES = [
    # name, kind, N, H, W, Cin, Cout, k, stride
    ('conv3_valid_32_64', 'conv', 2, 30, 26, 32, 64, 3, 1),
    ('conv3_valid_128_256', 'conv', 2, 27, 27, 128, 256, 3, 1),
    ('k5s2_256_64', 'deconv', 2, 13, 9, 256, 64, 5, 2),
    ('k5s2_32_32', 'deconv', 2, 27, 25, 32, 32, 5, 2),
    ('k5s2_64_24', 'deconv', 1, 11, 11, 64, 24, 5, 2),
]


@pytest.mark.parametrize('case', AFFINE_CASES, ids=[c[0] for c in AFFINE_CASES])
def test_conv_deconv_fwd_with_folded_batchnorm(cuda, case):
    """seg_conv2d_fwd_affine / seg_deconv2d_fwd_affine + seg_batchnorm_fold (layer -> ReLU ->
    inference batch-norm as one launch, /root/reference/models/deconvolution.py:120-170)
    against the oracle's conv -> relu -> batch_norm(is_training=False) and against the unfused
    C-ABI pair.  The fused form normalises the fp32 accumulator (one bf16 rounding instead of
    two), so it is compared within the bf16 tolerance, not bitwise; padded output channels
    must stay zero."""
    name, kind, Nb, H, W, Ci, Co, k, s = case
    g = _gen(31)
    x = bfr(torch.rand(Nb, H, W, Ci, generator=g) - 0.3)
    b = torch.randn(Co, generator=g) * 0.1
    beta = torch.randn(Co, generator=g) * 0.2
    mean = torch.rand(Co, generator=g) * 0.3
    var = torch.rand(Co, generator=g) * 0.5 + 0.02
    cin_pad, cout_pad = pad16(Ci), pad16(Co)
    if kind == 'conv':
        w = bfr(torch.randn(k, k, Ci, Co, generator=g) * 0.1)
        z = T.conv2d(x, w, b, s, 'VALID')
        w_d = shadow_conv(w, cin_pad, cout_pad)
    else:
        w = bfr(torch.randn(k, k, Co, Ci, generator=g) * 0.1)
        z = T.conv2d_transpose(x, w, b, s, 'VALID')
        w_d = shadow_deconv(w, cin_pad, cout_pad)
    y_ref, _, _ = T.batch_norm(torch.relu(z), beta, mean, var, False)
    OH, OW = z.shape[1], z.shape[2]
    x_d, b_d = dev_bf16(x, cin_pad), b.cuda()
    mean_d, var_d, beta_d = mean.cuda(), var.cuda(), beta.cuda()
    fold = torch.full((2, cout_pad), float('nan'), dtype=torch.float32, device='cuda')
    st = N.stream_ptr()
    N.call('seg_batchnorm_fold', N.ptr(mean_d), N.ptr(var_d), 1e-3, N.ptr(beta_d), Co, cout_pad,
           N.ptr(fold[0]), N.ptr(fold[1]), st)
    d = desc(k, s, (0, 0, 0, 0), Ci, Co, cin_pad, cout_pad, N.EPI_BIAS | N.EPI_RELU, N.IMPL_UMMA)
    y_d = torch.full((Nb, OH, OW, cout_pad), float('nan'), dtype=torch.bfloat16, device='cuda')
    entry = 'seg_conv2d_fwd_affine' if kind == 'conv' else 'seg_deconv2d_fwd_affine'
    y_d[..., Co:] = 0
    N.call(entry, ctypes.byref(d), N.vref(x_d), N.ptr(w_d), N.ptr(b_d), N.ptr(fold[0]),
           N.ptr(fold[1]), N.vref(y_d[..., :Co]), st)
    assert 'hconv' in N.load().seg_last_kernel_name().decode()
    # the unfused pair
    t_d = torch.zeros(Nb, OH, OW, cout_pad, dtype=torch.bfloat16, device='cuda')
    u_d = torch.zeros_like(t_d)
    plain = 'seg_conv2d_fwd' if kind == 'conv' else 'seg_deconv2d_fwd'
    if kind == 'conv':
        N.call(plain, ctypes.byref(d), N.vref(x_d), None, N.ptr(w_d), N.ptr(b_d),
               N.vref(t_d[..., :Co]), st)
    else:
        N.call(plain, ctypes.byref(d), N.vref(x_d), N.ptr(w_d), N.ptr(b_d),
               N.vref(t_d[..., :Co]), st)
    N.call('seg_batchnorm_infer', N.vref(t_d[..., :Co]), N.ptr(mean_d), N.ptr(var_d), 1e-3,
           N.ptr(beta_d), N.vref(u_d[..., :Co]), st)
    sync()
    f = fold.cpu()
    sc = torch.rsqrt(var + 1e-3)
    assert torch.allclose(f[0, :Co], sc, rtol=1e-5) and torch.allclose(f[1, :Co], beta - mean * sc,
                                                                      rtol=1e-5, atol=1e-6)
    assert torch.all(f[:, Co:] == 0)
    got = y_d.float().cpu()
    rec = {'case': name, 'vs_oracle': rel_l2(got[..., :Co], y_ref),
           'vs_unfused': rel_l2(got[..., :Co], u_d.float().cpu()[..., :Co]),
           'unfused_vs_oracle': rel_l2(u_d.float().cpu()[..., :Co], y_ref)}
    report('conv_affine', rec)
    assert torch.all(got[..., Co:] == 0)                       # padded channels untouched
    assert rec['vs_oracle'] < TOL_BF16 and rec['vs_unfused'] < TOL_BF16, rec


def test_conv_crop_view_and_slice_output(cuda):
    """Crop views as inputs (tf.image.resize_image_with_crop_or_pad as a TMA base
    offset) and channel-slice outputs."""
    g = _gen(11)
    big = bfr(torch.rand(2, 24, 24, 32, generator=g) - 0.3)
    w = bfr(torch.randn(3, 3, 32, 32, generator=g) * 0.1)
    b = torch.zeros(32)
    crop = T.crop_or_pad(big, 16, 16)
    ref = torch.relu(T.conv2d(crop, w, b, 1, 'VALID'))
    w_d, b_d = shadow_conv(w, 32, 32), b.cuda()      # keep alive across the async launch
    for impl_name, impl in IMPLS:
        big_d = dev_bf16(big)
        view = big_d[:, 4:20, 4:20, :]
        out = torch.zeros(2, 14, 14, 64, dtype=torch.bfloat16, device='cuda')
        d = desc(3, 1, (0, 0, 0, 0), 32, 32, 32, 32, N.EPI_BIAS | N.EPI_RELU, impl)
        N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(view), None, N.ptr(w_d),
               N.ptr(b_d), N.vref(out[..., 32:]), N.stream_ptr())
        sync()
        assert rel_l2(out[..., 32:].float().cpu(), ref) < TOL_BF16, impl_name
        assert float(out[..., :32].abs().max()) == 0.0
