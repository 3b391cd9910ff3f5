"""Host mirror of /root/reference/utils/ops.py:28-110 (DCGAN-style layer helpers).

Same names, argument order, defaults and variable naming (`<name>/w`,
`<name>/biases`, `<scope>/Matrix`, `<scope>/bias`) as the reference; variables
live in a module-level registry that plays the role of TF's variable scopes
(calling a helper twice with the same `name` reuses the variables, like
`reuse=True`).  Inputs are NHWC device tensors (fp32 or bf16); outputs are bf16
NHWC device tensors computed by the C-ABI kernels (include/segb200.h):

  conv2d    -> seg_conv2d_fwd   (SAME, bias, NO activation — reference :58-69)
  deconv2d  -> seg_deconv2d_fwd (SAME, caller-supplied output_shape — :71-94)
  linear    -> seg_conv2d_fwd on a [B,1,1,K] view (x @ Matrix + bias — :99-110)
  batch_norm-> seg_batchnorm_* (decay .9, eps 1e-5, scale=True — :35-49)

No segmentation model of the reference imports these helpers;
`lrelu`, `concat` and `conv_cond_concat` are trivial elementwise / layout
helpers outside the kernel hot path and use torch tensor ops.
"""
import ctypes
import math

import numpy as np
import torch

from .. import engine as E
from .. import native as N

BF16 = torch.bfloat16
_VARS = {}
_RNG = np.random.default_rng(0)


def reset_variables(seed=0):
    """Forget all variables (a fresh TF graph)."""
    global _RNG
    _VARS.clear()
    _RNG = np.random.default_rng(seed)


def get_variable(name):
    return _VARS[name]


def _truncated_normal(shape, stddev):
    """tf.truncated_normal_initializer: resample beyond 2 sigma."""
    x = _RNG.normal(0.0, stddev, size=shape)
    bad = np.abs(x) > 2 * stddev
    while bad.any():
        x[bad] = _RNG.normal(0.0, stddev, size=int(bad.sum()))
        bad = np.abs(x) > 2 * stddev
    return x.astype(np.float32)


def _var(name, shape, init):
    if name not in _VARS:
        _VARS[name] = torch.from_numpy(np.asarray(init(shape), dtype=np.float32)).cuda()
    v = _VARS[name]
    assert tuple(v.shape) == tuple(shape), 'variable %s exists with another shape' % name
    return v


def _as_bf16_padded(x):
    """NHWC tensor -> bf16 with channels zero-padded to a multiple of 16."""
    c = x.shape[-1]
    cp = E.pad16(c)
    if x.dtype == BF16 and cp == c and x.is_contiguous():
        return x
    out = torch.zeros(x.shape[:-1] + (cp,), dtype=BF16, device=x.device)
    out[..., :c] = x.to(BF16)
    return out


def _shadow(w, pad_dims):
    s = torch.zeros(w.shape[:2] + tuple(pad_dims), dtype=BF16, device=w.device)
    s[:, :, :w.shape[2], :w.shape[3]] = w.to(BF16)
    return s


class batch_norm(object):
    """`tf.contrib.layers.batch_norm(decay=momentum, updates_collections=None,
    epsilon, scale=True, is_training=train)` (reference :35-49)."""

    def __init__(self, epsilon=1e-5, momentum=0.9, name="batch_norm"):
        self.epsilon = epsilon
        self.momentum = momentum
        self.name = name

    def __call__(self, x, train=True):
        c = x.shape[-1]
        beta = _var(self.name + '/beta', (c,), np.zeros)
        gamma = _var(self.name + '/gamma', (c,), np.ones)
        mm = _var(self.name + '/moving_mean', (c,), np.zeros)
        mv = _var(self.name + '/moving_variance', (c,), np.ones)
        xb = _as_bf16_padded(x)
        y = torch.zeros_like(xb)
        st = N.stream_ptr()
        xs, ys = xb[..., :c], y[..., :c]
        if train:
            s = torch.zeros(4, c, dtype=torch.float32, device=x.device)
            N.call('seg_batchnorm_stats', N.vref(xs), N.ptr(s[0]), N.ptr(s[1]), st)
            count = x.shape[0] * x.shape[1] * x.shape[2]
            N.call('seg_batchnorm_finalize', N.ptr(s[0]), N.ptr(s[1]), count, c, self.epsilon,
                   self.momentum, N.ptr(s[2]), N.ptr(s[3]), N.ptr(mm), N.ptr(mv), st)
            scale = s[3] * gamma                     # gamma folded into rstd (C elements)
            N.call('seg_batchnorm_apply', N.vref(xs), N.ptr(s[2]), N.ptr(scale), N.ptr(beta),
                   N.vref(ys), st)
        else:
            scale = gamma * torch.rsqrt(mv + self.epsilon)
            N.call('seg_batchnorm_apply', N.vref(xs), N.ptr(mm), N.ptr(scale), N.ptr(beta),
                   N.vref(ys), st)
        return ys


def concat(tensors, axis, *args, **kwargs):
    return torch.cat(list(tensors), dim=axis)


def conv_cond_concat(x, y):
    """Concatenate conditioning vector on feature map axis (reference :51-56)."""
    return concat([x, y.to(x.dtype) * torch.ones(x.shape[0], x.shape[1], x.shape[2], y.shape[3],
                                                  dtype=x.dtype, device=x.device)], 3)


def conv2d(input_, output_dim, k_h=5, k_w=5, d_h=2, d_w=2, stddev=0.02, name="conv2d"):
    assert k_h == k_w and d_h == d_w, 'square kernels / strides only'
    cin = input_.shape[-1]
    w = _var(name + '/w', (k_h, k_w, cin, output_dim), lambda s: _truncated_normal(s, stddev))
    b = _var(name + '/biases', (output_dim,), np.zeros)
    x = _as_bf16_padded(input_)
    cin_pad, cout_pad = x.shape[-1], E.pad16(output_dim)
    Nb, H, W = x.shape[0], x.shape[1], x.shape[2]
    pt, pb = E.same_pad(H, k_h, d_h)
    pl, pr = E.same_pad(W, k_w, d_w)
    Ho, Wo = -(-H // d_h), -(-W // d_w)
    y = torch.zeros(Nb, Ho, Wo, cout_pad, dtype=BF16, device=x.device)
    d = N.SegConvDesc(k_h, k_w, d_h, pt, pl, pb, pr, cin, output_dim, cin_pad, cout_pad,
                      N.EPI_BIAS, N.IMPL_UMMA)
    sh = _shadow(w, (cin_pad, cout_pad))
    N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(x), None, N.ptr(sh), N.ptr(b),
           N.vref(y[..., :output_dim]), N.stream_ptr())
    torch.cuda.current_stream().synchronize()      # `sh` is a temporary
    return y[..., :output_dim]


def deconv2d(input_, output_shape, k_h=5, k_w=5, d_h=2, d_w=2, stddev=0.02, name="deconv2d",
             with_w=False):
    assert k_h == k_w and d_h == d_w, 'square kernels / strides only'
    cin, cout = input_.shape[-1], output_shape[-1]
    # filter : [height, width, output_channels, in_channels]
    w = _var(name + '/w', (k_h, k_w, cout, cin), lambda s: _RNG.normal(0, stddev, s))
    b = _var(name + '/biases', (cout,), np.zeros)
    x = _as_bf16_padded(input_)
    cin_pad, cout_pad = x.shape[-1], E.pad16(cout)
    assert output_shape[1] == x.shape[1] * d_h and output_shape[2] == x.shape[2] * d_w, \
        'SAME transposed conv: output must be input * stride'
    tot = max(k_h - d_h, 0)
    y = torch.zeros(x.shape[0], output_shape[1], output_shape[2], cout_pad, dtype=BF16,
                    device=x.device)
    impl = N.IMPL_UMMA if (k_h == d_h) else N.IMPL_SIMT
    d = N.SegConvDesc(k_h, k_w, d_h, tot // 2, tot // 2, tot - tot // 2, tot - tot // 2, cin, cout,
                      cin_pad, cout_pad, N.EPI_BIAS, impl)
    sh = torch.zeros(k_h, k_w, cout_pad, cin_pad, dtype=BF16, device=x.device)
    sh[:, :, :cout, :cin] = w.to(BF16)
    N.call('seg_deconv2d_fwd', ctypes.byref(d), N.vref(x), N.ptr(sh), N.ptr(b),
           N.vref(y[..., :cout]), N.stream_ptr())
    torch.cuda.current_stream().synchronize()
    deconv = y[..., :cout]
    if with_w:
        return deconv, w, b
    return deconv


def lrelu(x, leak=0.2, name="lrelu"):
    return torch.maximum(x, leak * x)


def linear(input_, output_size, scope=None, stddev=0.02, bias_start=0.0, with_w=False):
    shape = list(input_.shape)
    sc = scope or "Linear"
    matrix = _var(sc + '/Matrix', (shape[1], output_size), lambda s: _RNG.normal(0, stddev, s))
    bias = _var(sc + '/bias', (output_size,), lambda s: np.full(s, bias_start))
    x = _as_bf16_padded(input_.reshape(shape[0], 1, 1, shape[1]))
    kin_pad, n_pad = x.shape[-1], E.pad16(output_size)
    sh = torch.zeros(1, 1, kin_pad, n_pad, dtype=BF16, device=x.device)
    sh[0, 0, :shape[1], :output_size] = matrix.to(BF16)
    y = torch.zeros(shape[0], 1, 1, output_size, dtype=torch.float32, device=x.device)
    d = N.SegConvDesc(1, 1, 1, 0, 0, 0, 0, shape[1], output_size, kin_pad, n_pad,
                      N.EPI_BIAS | N.EPI_OUT_F32, N.IMPL_UMMA)
    N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(x), None, N.ptr(sh), N.ptr(bias), N.vref(y),
           N.stream_ptr())
    torch.cuda.current_stream().synchronize()
    out = y.reshape(shape[0], output_size)
    if with_w:
        return out, matrix, bias
    return out
