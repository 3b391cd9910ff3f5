"""CPU tests (-m "not gpu") that pin the oracle as far as the reference allows:

  * against tests/golden/upsampling.npz — outputs of the reference's OWN
    utils/upsampling.py run in the build container (tests/golden/make_golden.py);
  * against the closed-form / printed shape chains of the reference graphs
    (SURVEY.md §8c);
  * against an independent naive-loop restatement (oracle/naive.py);
  * Philox against the Random123 known-answer vectors.
Everything living in TensorFlow itself stays "parity unpinned" (oracle/__init__.py).
"""
import os

import numpy as np
import pytest
import torch

from oracle import naive, nets, tf_ops as T

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'upsampling.npz'))


def test_upsampling_matches_reference_golden():
    for f in (1, 2, 3, 4, 8, 16, 32):
        assert T.get_kernel_size(f) == int(GOLD['ksize_f%d' % f])
    for k in (1, 2, 3, 4, 5, 16, 32, 64):
        assert np.array_equal(np.asarray(T.upsample_filt(k), dtype=np.float64), GOLD['filt_k%d' % k])
    for f, c in ((2, 2), (2, 21), (8, 21), (16, 3), (32, 2)):
        w = T.bilinear_upsample_weights(f, c)
        g = GOLD['weights_f%d_c%d' % (f, c)]
        assert w.dtype == np.float32 and w.shape == g.shape and np.array_equal(w, g)


def test_upsampling_closed_forms():
    # utils/upsampling.py:10 and :17-24 evaluated by hand
    assert [T.get_kernel_size(f) for f in (2, 8, 16, 32)] == [4, 16, 32, 64]
    assert np.allclose(T.upsample_filt(4), np.outer([.25, .75, .75, .25], [.25, .75, .75, .25]))
    w = T.bilinear_upsample_weights(2, 3)
    assert w[:, :, 0, 1].max() == 0 and w[:, :, 2, 2].max() > 0      # channel-diagonal


def test_product_host_helpers_match_reference_golden():
    """segmentation_b200.utils.upsampling (the product's host mirror) against
    the same golden file."""
    from segmentation_b200.utils import upsampling as U
    for f in (1, 2, 3, 4, 8, 16, 32):
        assert U.get_kernel_size(f) == int(GOLD['ksize_f%d' % f])
    for k in (1, 2, 3, 4, 5, 16, 32, 64):
        assert np.array_equal(np.asarray(U.upsample_filt(k), dtype=np.float64), GOLD['filt_k%d' % k])
    for f, c in ((2, 2), (2, 21), (8, 21), (16, 3), (32, 2)):
        assert np.array_equal(U.bilinear_upsample_weights(f, c), GOLD['weights_f%d_c%d' % (f, c)])


def test_shape_chains_and_param_counts():
    assert nets.unet_out_size(256) == 68 and nets.unet_out_size(512) == 324
    p = nets.unet_params()
    assert sum(v.numel() for v in p.values()) == 7760130 and len(p) == 46
    taps = {}
    y = nets.unet_forward(nets.unet_params(n_kernels=2), torch.rand(1, 256, 256, 3), taps=taps)
    assert tuple(y.shape) == (1, 68, 68, 2)
    hs = [taps[k].shape[1] for k in ('conv1_1', 'conv1_2', 'pool1', 'conv2_2', 'pool2', 'conv3_2',
                                     'pool3', 'conv4_2', 'pool4', 'conv5_2', 'upconv1', 'conv6_2',
                                     'upconv2', 'conv7_2', 'upconv3', 'conv8_2', 'upconv4', 'conv9_2')]
    assert hs == [254, 252, 127, 123, 61, 57, 28, 24, 12, 8, 16, 12, 24, 20, 40, 36, 72, 68]
    pf = nets.fcn_params()
    assert sum(v.numel() for v in pf.values()) == 2320895
    yf = nets.fcn_forward(nets.fcn_params(n_kernels=2, n_classes=5), torch.rand(1, 64, 64, 3))
    assert tuple(yf.shape) == (1, 64, 64, 5)
    for t in ('32s', '16s'):
        yf = nets.fcn_forward(nets.fcn_params(n_kernels=2, n_classes=5, fcn_type=t),
                              torch.rand(1, 64, 64, 3), fcn_type=t)
        assert tuple(yf.shape) == (1, 64, 64, 5)
    pd = nets.deconv_params()
    assert sum(v.numel() for k, v in pd.items() if k.endswith(('weights', 'biases'))) == 876776
    taps = {}
    yd = nets.deconv_forward(nets.deconv_params(n_kernels=2), torch.rand(1, 512, 512, 3), taps=taps)
    assert tuple(yd.shape) == (1, 512, 512, 2)
    assert [taps[k].shape[1] for k in ('conv1_0', 'pool1', 'conv2_0', 'pool2', 'conv3_0', 'pool3',
                                       'conv4_0', 'deconv1_0', 'deconv2_0', 'deconv2_1', 'resize',
                                       'deconv3_0')] == [256, 128, 126, 42, 40, 13, 11, 25, 53, 109,
                                                         256, 512]


def test_same_padding_arithmetic():
    # DeconvModel conv1_0 at 1024: k5 s2 SAME -> 1 before / 2 after (SURVEY §8c [TF-sem 3])
    assert T.same_pad(1024, 5, 2) == (1, 2)
    assert T.same_pad(512, 3, 1) == (1, 1)
    assert T.same_pad(7, 3, 2) == (1, 1)
    assert T.same_pad(8, 3, 2) == (0, 1)
    assert T.deconv_out_size(25, 5, 2, 'VALID') == 53 and T.deconv_out_size(8, 2, 2, 'VALID') == 16
    assert T.deconv_out_size(16, 4, 2, 'SAME') == 32


@pytest.mark.parametrize('k,s,padding', [(3, 1, 'VALID'), (3, 1, 'SAME'), (5, 2, 'SAME'),
                                         (1, 1, 'SAME'), (3, 2, 'VALID')])
def test_conv_vs_naive(k, s, padding):
    g = np.random.default_rng(0)
    x = g.normal(size=(2, 9, 8, 3)).astype(np.float32)
    w = g.normal(size=(k, k, 3, 4)).astype(np.float32)
    b = g.normal(size=(4,)).astype(np.float32)
    y = T.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), s, padding)
    assert np.allclose(y.numpy(), naive.conv2d(x, w, b, s, padding), atol=1e-4)


@pytest.mark.parametrize('k,s,padding', [(2, 2, 'VALID'), (5, 2, 'VALID'), (4, 2, 'SAME'),
                                         (16, 8, 'SAME'), (3, 1, 'SAME')])
def test_conv_transpose_vs_naive(k, s, padding):
    g = np.random.default_rng(1)
    x = g.normal(size=(2, 4, 5, 3)).astype(np.float32)
    w = g.normal(size=(k, k, 2, 3)).astype(np.float32)
    b = g.normal(size=(2,)).astype(np.float32)
    y = T.conv2d_transpose(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), s, padding)
    ref = naive.conv2d_transpose(x, w, b, s, padding)
    assert y.shape == ref.shape and np.allclose(y.numpy(), ref, atol=1e-4)


def test_conv_transpose_is_conv_input_gradient():
    """[TF-sem 2]: conv2d_transpose == gradient of conv2d w.r.t. its input."""
    g = torch.Generator().manual_seed(0)
    for k, s, padding, n in ((5, 2, 'VALID', 13), (4, 2, 'SAME', 12), (2, 2, 'VALID', 10)):
        w = torch.randn(k, k, 3, 4, generator=g)          # as conv: HWIO with I=3 (big side)
        xin = torch.randn(1, n, n, 3, generator=g, requires_grad=True)
        y = T.conv2d(xin, w, None, s, padding)
        dy = torch.randn(y.shape, generator=g)
        (dx,) = torch.autograd.grad(y, xin, dy)
        # the same filter seen as a transposed-conv filter [kh,kw,Cout=3,Cin=4]
        up = T.conv2d_transpose(dy, w, None, s, padding)
        if up.shape[1] != n:       # VALID when (n-k) % s != 0: tail rows receive no gradient
            pad = torch.zeros(1, n, n, 3)
            pad[:, :up.shape[1], :up.shape[2]] = up
            up = pad
        assert torch.allclose(up, dx, atol=1e-4)


def test_maxpool_first_max_and_tail_drop():
    x = np.zeros((1, 5, 5, 1), np.float32)
    x[0, 0, 1, 0] = 1.0; x[0, 1, 0, 0] = 1.0           # tie inside window (0,0): slot 1 wins
    y, slot = T.max_pool_with_argmax(torch.from_numpy(x), 2, 2)
    assert tuple(y.shape) == (1, 2, 2, 1)              # 5 -> 2: last row/col dropped
    assert int(slot[0, 0, 0, 0]) == 1 and int(slot[0, 1, 1, 0]) == 0
    g = np.random.default_rng(2)
    xr = np.round(g.normal(size=(2, 7, 9, 3)) * 2).astype(np.float32) / 2
    for k in (2, 3):
        y, s = T.max_pool_with_argmax(torch.from_numpy(xr), k, k)
        yn, sn = naive.max_pool_with_argmax(xr, k, k)
        assert np.array_equal(y.numpy(), yn) and np.array_equal(s.numpy(), sn)
    # gradient is routed to the first max only
    xt = torch.from_numpy(x).requires_grad_(True)
    T.max_pool(xt, 2, 2).sum().backward()
    assert xt.grad[0, 0, 1, 0] == 1 and xt.grad[0, 1, 0, 0] == 0
    flat = T.argmax_slot_to_flat(slot, 2, 2, 5, 5)
    assert int(flat[0, 0, 0, 0]) == 1                  # ((0*5+0)*5+1)*1+0


def test_resize_bilinear_legacy_vs_naive():
    g = np.random.default_rng(3)
    x = g.normal(size=(1, 5, 7, 2)).astype(np.float32)
    for oh, ow in ((12, 16), (5, 7), (3, 4), (11, 11)):
        y = T.resize_bilinear(torch.from_numpy(x), oh, ow).numpy()
        assert np.allclose(y, naive.resize_bilinear(x, oh, ow), atol=1e-5)
    # legacy (no half-pixel centres): output pixel 0 == input pixel 0 exactly
    assert np.array_equal(T.resize_bilinear(torch.from_numpy(x), 12, 16).numpy()[:, 0, 0], x[:, 0, 0])


def test_crop_or_pad():
    x = torch.arange(2 * 6 * 6).float().view(2, 6, 6, 1)
    assert torch.equal(T.crop_or_pad(x, 2, 2), x[:, 2:4, 2:4])
    assert torch.equal(T.crop_or_pad(x, 3, 3), x[:, 1:4, 1:4])       # offset floor((6-3)/2)=1
    p = T.crop_or_pad(x, 9, 6)
    assert tuple(p.shape) == (2, 9, 6, 1) and torch.equal(p[:, 1:7], x) and p[:, 0].abs().sum() == 0
    assert p[:, 7:].abs().sum() == 0                                   # before=1, after=2


def test_softmax_xent_and_adam_formulas():
    g = np.random.default_rng(4)
    lg = g.normal(size=(2, 3, 3, 5)).astype(np.float32)
    lab = g.integers(0, 5, (2, 3, 3, 1)).astype(np.uint8)
    a = float(T.softmax_xent_mean(torch.from_numpy(lg), torch.from_numpy(lab)))
    assert abs(a - naive.softmax_xent_mean(lg, lab)) < 1e-5
    # gradient = (softmax - onehot)/pixels   [TF-sem 9]
    t = torch.from_numpy(lg).requires_grad_(True)
    T.softmax_xent_mean(t, torch.from_numpy(lab)).backward()
    sm = torch.softmax(torch.from_numpy(lg), -1)
    oh = torch.nn.functional.one_hot(torch.from_numpy(lab).long().squeeze(-1), 5).float()
    assert torch.allclose(t.grad, (sm - oh) / 18, atol=1e-6)
    # TF Adam, epsilon outside the bias correction [TF-sem 11]; first step = -lr*sign(g) (approx)
    p, m, v = T.adam_update(torch.zeros(3), torch.tensor([1.0, -2.0, 0.5]), torch.zeros(3),
                            torch.zeros(3), 1, 1e-3)
    assert torch.allclose(p, torch.tensor([-1e-3, 1e-3, -1e-3]), atol=1e-8)


def test_sigmoid_argmax_ties_to_class0():
    lg = torch.tensor([[[[25.0, 40.0]]]])
    sig, lab = T.sigmoid_argmax(lg)
    assert float(sig[0, 0, 0, 0]) == 1.0 and float(sig[0, 0, 0, 1]) == 1.0 and float(lab) == 0.0


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    assert naive.philox4x32_10_scalar([0, 0, 0, 0], [0, 0]) == \
        [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert naive.philox4x32_10_scalar([0xffffffff] * 4, [0xffffffff] * 2) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert naive.philox4x32_10_scalar([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                      [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # vectorised oracle == scalar
    out = T.philox4x32_10(np.arange(5), np.zeros(5), np.full(5, 7), np.zeros(5), 123, 456)
    for i in range(5):
        assert [int(w[i]) for w in out] == naive.philox4x32_10_scalar([i, 0, 7, 0], [123, 456])
    m = T.dropout_keep_mask(10000, seed=1, stream=2)
    assert 0.47 < m.mean() < 0.53
    assert not np.array_equal(m, T.dropout_keep_mask(10000, seed=1, stream=3))


def test_batch_norm_slim_defaults():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 3, 3, 2, generator=g)
    y, m, v = T.batch_norm(x, torch.tensor([0.5, -0.5]), torch.zeros(2), torch.ones(2), True)
    mu = x.mean((0, 1, 2)); var = x.var((0, 1, 2), unbiased=False)
    assert torch.allclose(y, (x - mu) / torch.sqrt(var + 1e-3) + torch.tensor([0.5, -0.5]), atol=1e-5)
    assert torch.allclose(m, 0.001 * mu, atol=1e-7) and torch.allclose(v, 0.999 + 0.001 * var, atol=1e-6)


def test_bf16_emulation_close_to_fp32_and_train_step_decreases_loss():
    p = nets.unet_params(n_kernels=2, seed=1)
    g = np.random.default_rng(0)
    x = torch.from_numpy(g.random((1, 188, 188, 3), dtype=np.float32))
    y = torch.from_numpy(g.integers(0, 2, (1, 188, 188, 1)).astype(np.uint8))
    l32, lg32, g32 = nets.loss_and_grads(lambda q, xx: nets.unet_forward(q, xx), p, x, y)
    l16, lg16, g16 = nets.loss_and_grads(lambda q, xx: nets.unet_forward(q, xx, prec=T.BF16), p, x, y)
    assert abs(float(l32) - float(l16)) < 5e-3
    assert float((lg32 - lg16).norm() / lg32.norm()) < 5e-2
    st = nets.AdamState(p)
    losses = [nets.train_step(lambda q, xx: nets.unet_forward(q, xx), p, st, x, y, lr=1e-2)
              for _ in range(4)]
    assert losses[-1] < losses[0] and st.step == 4


def test_inference_batchnorm_commutes_with_maxpool_bitwise():
    """The identity seg_maxpool_bn_infer relies on: slim.batch_norm with moving
    statistics is an increasing map (slope rsqrt(var + eps) > 0) and every rounding on the way
    (fp32 arithmetic, bf16 storage) is monotonic, so bn(pool(x)) == pool(bn(x)) bit for bit —
    the order the reference states is conv -> bn -> pool (models/deconvolution.py:126-138)."""
    g = torch.Generator().manual_seed(3)
    for k in (2, 3):
        C = 16
        x = T.bf16_round(torch.relu(torch.randn(2, 12, 18, C, generator=g)))
        beta = torch.randn(C, generator=g) * 0.2
        mean = torch.rand(C, generator=g) * 0.5
        var = torch.rand(C, generator=g) * 0.5 + 0.01

        def bn(t):
            y, _, _ = T.batch_norm(t, beta, mean, var, False)
            return T.bf16_round(y)

        a = T.max_pool(bn(x), k, k)
        b = bn(T.max_pool(x, k, k))
        assert torch.equal(a, b)


def test_fcn_upscore_gradient_is_a_16x16_window_gather():
    """The formulation seg_upscore8_xent_fwd_bwd uses for the input gradient of the x8 bilinear
    transposed conv (models/fcn.py:207-220): dx[i, j, c] = sum over the 16 x 16 output window
    starting at (8i - 4, 8j - 4) of dy * filt, taps outside the image contributing nothing —
    checked against autograd through the oracle's dense conv2d_transpose."""
    g = torch.Generator().manual_seed(5)
    B, h, w, C, f = 1, 5, 4, 3, 8
    k = T.get_kernel_size(f)
    x = torch.randn(B, h, w, C, generator=g, requires_grad=True)
    y = T.bilinear_upsample(x, f)
    assert tuple(y.shape) == (B, f * h, f * w, C)
    dy = torch.randn(y.shape, generator=g)
    (dx_ref,) = torch.autograd.grad(y, x, dy)
    filt = torch.from_numpy(np.asarray(T.upsample_filt(k), dtype=np.float32))
    dx = torch.zeros_like(dx_ref)
    before = (k - f) // 2
    for i in range(h):
        for j in range(w):
            for a in range(k):
                oy = i * f + a - before
                if not 0 <= oy < f * h:
                    continue
                for b in range(k):
                    ox = j * f + b - before
                    if 0 <= ox < f * w:
                        dx[0, i, j] += dy[0, oy, ox] * filt[a, b]
    assert torch.allclose(dx, dx_ref, rtol=1e-5, atol=1e-6)
    # and the forward: every output pixel reads exactly two input rows and two input columns
    yy = torch.zeros_like(y)
    xd = x.detach()
    for oy in range(f * h):
        for ox in range(f * w):
            ty, tx = oy + before, ox + before
            for i in (ty // f, ty // f - 1):
                for j in (tx // f, tx // f - 1):
                    if 0 <= i < h and 0 <= j < w:
                        yy[0, oy, ox] += xd[0, i, j] * filt[ty - i * f, tx - j * f]
    assert torch.allclose(yy, y.detach(), rtol=1e-5, atol=1e-6)
