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
    """A trainable variable: an fp32 device tensor with requires_grad, so that gradients
    of anything built from these helpers arrive in `get_variable(name).grad` after
    `.backward()` - the torch analogue of tf.gradients over tf.get_variable."""
    if name not in _VARS:
        v = torch.from_numpy(np.asarray(init(shape), dtype=np.float32)).cuda()
        v.requires_grad_(True)
        _VARS[name] = v
    v = _VARS[name]
    assert tuple(v.shape) == tuple(shape), 'variable %s exists with another shape' % name
    return v


def _state(name, shape, init):
    """Non-trainable state (batch-norm moving statistics)."""
    if name not in _VARS:
        _VARS[name] = torch.from_numpy(np.asarray(init(shape), dtype=np.float32)).cuda()
    return _VARS[name]


def _as_bf16_padded(x, cp=None):
    """NHWC tensor -> bf16 with channels zero-padded to a multiple of 16 (or to cp)."""
    c = x.shape[-1]
    cp = E.pad16(c) if cp is None else cp
    if x.dtype == BF16 and cp == c and x.is_contiguous():
        return x
    out = torch.zeros(x.shape[:-1] + (cp,), dtype=BF16, device=x.device)
    out[..., :c] = x.to(BF16)
    return out


def _shadow(w, pad_dims):
    s = torch.zeros(w.shape[:2] + tuple(pad_dims), dtype=BF16, device=w.device)
    s[:, :, :w.shape[2], :w.shape[3]] = w.detach().to(BF16)
    return s


# Temporaries (padded inputs, bf16 shadows) are plain torch tensors: the caching allocator
# reuses their memory in stream order, so no call here synchronises with the host.
class _ConvFn(torch.autograd.Function):
    """tf.nn.conv2d(SAME) + bias_add and its two gradients through the C ABI
    (seg_conv2d_fwd / seg_conv2d_dgrad / seg_conv2d_wgrad)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, out_f32):
        k, cin, cout = w.shape[0], w.shape[2], w.shape[3]
        xb = _as_bf16_padded(x)
        cin_pad, cout_pad = xb.shape[-1], E.pad16(cout)
        Nb, H, W = xb.shape[0], xb.shape[1], xb.shape[2]
        pt, pb = E.same_pad(H, k, stride)
        pl, pr = E.same_pad(W, k, stride)
        Ho, Wo = -(-H // stride), -(-W // stride)
        sh = _shadow(w, (cin_pad, cout_pad))
        flags = N.EPI_BIAS | (N.EPI_OUT_F32 if out_f32 else 0)
        y = torch.zeros(Nb, Ho, Wo, cout if out_f32 else cout_pad,
                        dtype=torch.float32 if out_f32 else BF16, device=xb.device)
        d = N.SegConvDesc(k, k, stride, pt, pl, pb, pr, cin, cout, cin_pad, cout_pad, flags,
                          N.IMPL_UMMA)
        N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(xb), None, N.ptr(sh), N.ptr(b.detach()),
               N.vref(y[..., :cout]), N.stream_ptr())
        ctx.save_for_backward(xb, sh)
        ctx.geom = (k, stride, (pt, pl, pb, pr), cin, cout, cin_pad, cout_pad, x.dtype)
        return y[..., :cout]

    @staticmethod
    def backward(ctx, gy):
        xb, sh = ctx.saved_tensors
        k, stride, pads, cin, cout, cin_pad, cout_pad, xdt = ctx.geom
        # the tcgen05 gradients cover stride 1; strided convolutions use the CUDA-core kernels
        impl = N.IMPL_UMMA if stride == 1 else N.IMPL_SIMT
        d = N.SegConvDesc(k, k, stride, pads[0], pads[1], pads[2], pads[3], cin, cout, cin_pad,
                          cout_pad, 0, impl)
        dz = _as_bf16_padded(gy, cout_pad)
        st = N.stream_ptr()
        dx = None
        if ctx.needs_input_grad[0]:
            dxp = torch.zeros_like(xb)
            N.call('seg_conv2d_dgrad', ctypes.byref(d), N.vref(dz), N.ptr(sh), N.vref(dxp), None,
                   None, None, st)
            dx = dxp[..., :cin].to(xdt)
        dw = torch.zeros(k, k, cin, cout, dtype=torch.float32, device=xb.device)
        db = torch.zeros(cout, dtype=torch.float32, device=xb.device)
        N.call('seg_conv2d_wgrad', ctypes.byref(d), N.vref(xb), None, N.vref(dz), N.ptr(dw),
               N.ptr(db), st)
        return dx, dw, db, None, None


class _DeconvFn(torch.autograd.Function):
    """tf.nn.conv2d_transpose(SAME) + bias_add and its gradients (seg_deconv2d_fwd /
    seg_deconv2d_dgrad / seg_deconv2d_wgrad + seg_bias_grad).  w is [kh, kw, cout, cin]."""

    @staticmethod
    def forward(ctx, x, w, b, stride, out_hw):
        k, cout, cin = w.shape[0], w.shape[2], w.shape[3]
        xb = _as_bf16_padded(x)
        cin_pad, cout_pad = xb.shape[-1], E.pad16(cout)
        tot = max(k - stride, 0)
        pads = (tot // 2, tot // 2, tot - tot // 2, tot - tot // 2)
        y = torch.zeros(xb.shape[0], out_hw[0], out_hw[1], cout_pad, dtype=BF16, device=xb.device)
        # k == stride runs on the tcgen05 kernels; SAME-cropped k > stride on the gather kernel
        impl = N.IMPL_UMMA if k == stride else N.IMPL_SIMT
        d = N.SegConvDesc(k, k, stride, pads[0], pads[1], pads[2], pads[3], cin, cout, cin_pad,
                          cout_pad, N.EPI_BIAS, impl)
        sh = torch.zeros(k, k, cout_pad, cin_pad, dtype=BF16, device=xb.device)
        sh[:, :, :cout, :cin] = w.detach().to(BF16)
        N.call('seg_deconv2d_fwd', ctypes.byref(d), N.vref(xb), N.ptr(sh), N.ptr(b.detach()),
               N.vref(y[..., :cout]), N.stream_ptr())
        ctx.save_for_backward(xb, sh)
        ctx.geom = (k, stride, pads, cin, cout, cin_pad, cout_pad, impl, x.dtype)
        return y[..., :cout]

    @staticmethod
    def backward(ctx, gy):
        xb, sh = ctx.saved_tensors
        k, stride, pads, cin, cout, cin_pad, cout_pad, impl, xdt = ctx.geom
        d = N.SegConvDesc(k, k, stride, pads[0], pads[1], pads[2], pads[3], cin, cout, cin_pad,
                          cout_pad, 0, impl)
        dz = _as_bf16_padded(gy, cout_pad)
        st = N.stream_ptr()
        dx = None
        if ctx.needs_input_grad[0]:
            dxp = torch.zeros_like(xb)
            N.call('seg_deconv2d_dgrad', ctypes.byref(d), N.vref(dz), N.ptr(sh), N.vref(dxp), None,
                   st)
            dx = dxp[..., :cin].to(xdt)
        dw = torch.zeros(k, k, cout, cin, dtype=torch.float32, device=xb.device)
        db = torch.zeros(cout, dtype=torch.float32, device=xb.device)
        N.call('seg_deconv2d_wgrad', ctypes.byref(d), N.vref(xb), N.vref(dz), N.ptr(dw), st)
        N.call('seg_bias_grad', N.vref(dz[..., :cout]), N.ptr(db), st)
        return dx, dw, db, None, None


class _BatchNormTrainFn(torch.autograd.Function):
    """Batch statistics, y = xhat * gamma + beta, and the batch-norm gradient
    (seg_batchnorm_stats / _finalize / _apply, seg_batchnorm_bwd_reduce / _bwd_apply)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mm, mv, eps, momentum):
        c = x.shape[-1]
        xb = _as_bf16_padded(x)
        y = torch.zeros_like(xb)
        st = N.stream_ptr()
        xs, ys = xb[..., :c], y[..., :c]
        s = torch.zeros(4, c, dtype=torch.float32, device=x.device)
        N.call('seg_batchnorm_stats', N.vref(xs), N.ptr(s[0]), N.ptr(s[1]), st)
        count = x.shape[0] * x.shape[1] * x.shape[2]
        N.call('seg_batchnorm_finalize', N.ptr(s[0]), N.ptr(s[1]), count, c, eps, momentum,
               N.ptr(s[2]), N.ptr(s[3]), N.ptr(mm), N.ptr(mv), st)
        scale = s[3] * gamma.detach()                # gamma folded into rstd (C elements)
        N.call('seg_batchnorm_apply', N.vref(xs), N.ptr(s[2]), N.ptr(scale), N.ptr(beta.detach()),
               N.vref(ys), st)
        ctx.save_for_backward(xb, s, gamma.detach())
        ctx.meta = (c, count, x.dtype)
        return ys

    @staticmethod
    def backward(ctx, gy):
        xb, s, gamma = ctx.saved_tensors
        c, count, xdt = ctx.meta
        st = N.stream_ptr()
        dy = _as_bf16_padded(gy, xb.shape[-1])
        red = torch.zeros(2, c, dtype=torch.float32, device=xb.device)   # sum dy, sum dy*xhat
        N.call('seg_batchnorm_bwd_reduce', N.vref(dy[..., :c]), N.vref(xb[..., :c]), N.ptr(s[2]),
               N.ptr(s[3]), N.ptr(red[0]), N.ptr(red[1]), st)
        dx = torch.zeros_like(xb)
        N.call('seg_batchnorm_bwd_apply', N.vref(dy[..., :c]), N.vref(xb[..., :c]), N.ptr(s[2]),
               N.ptr(s[3]), N.ptr(red[0]), N.ptr(red[1]), count, 0, N.vref(dx[..., :c]), st)
        # y = xhat * gamma + beta: the kernels differentiate xhat + beta; gamma scales dx
        dxs = (dx[..., :c].float() * gamma).to(xdt)
        return dxs, red[1].clone(), red[0].clone(), None, None, None, None


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
        mm = _state(self.name + '/moving_mean', (c,), np.zeros)
        mv = _state(self.name + '/moving_variance', (c,), np.ones)
        if train:
            return _BatchNormTrainFn.apply(x, gamma, beta, mm, mv, self.epsilon, self.momentum)
        # inference: a per-channel affine of the moving statistics (plain tensor algebra,
        # differentiable as it stands)
        scale = gamma * torch.rsqrt(mv + self.epsilon)
        return ((x.float() - mm) * scale + beta).to(BF16)


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
    return _ConvFn.apply(input_, w, b, d_h, False)


def deconv2d(input_, output_shape, k_h=5, k_w=5, d_h=2, d_w=2, stddev=0.02, name="deconv2d",
             with_w=False):
    assert k_h == k_w and d_h == d_w, 'square kernels / strides only'
    cin, cout = input_.shape[-1], output_shape[-1]
    # filter : [height, width, output_channels, in_channels]
    w = _var(name + '/w', (k_h, k_w, cout, cin), lambda s: _RNG.normal(0, stddev, s))
    b = _var(name + '/biases', (cout,), np.zeros)
    assert output_shape[1] == input_.shape[1] * d_h and output_shape[2] == input_.shape[2] * d_w, \
        'SAME transposed conv: output must be input * stride'
    deconv = _DeconvFn.apply(input_, w, b, d_h, (output_shape[1], output_shape[2]))
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
    # x @ Matrix + bias as a 1x1 convolution over a [B,1,1,K] view (fp32 output)
    y = _ConvFn.apply(input_.reshape(shape[0], 1, 1, shape[1]),
                      matrix.reshape(1, 1, shape[1], output_size), bias, 1, True)
    out = y.reshape(shape[0], output_size)
    if with_w:
        return out, matrix, bias
    return out
