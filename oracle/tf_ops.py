"""TF-1.x-semantics op restatement on torch-CPU fp32 (TEST INFRASTRUCTURE ONLY).

Every function restates the documented behaviour of the TensorFlow 1.x /
tf.contrib.slim op that a cited reference line lowers to ("[TF-sem]" in
SURVEY.md §8c).  TensorFlow itself is a third-party dependency that is NOT
under /root/reference and whose version the reference never pins, so this is
"parity unpinned" (see oracle/__init__.py).  Tensors are NHWC like the
reference's (`models/basemodel.py:146-150`).

All ops are differentiable through torch autograd so the backward pass of the
reference graph (tf.gradients under `AdamOptimizer.minimize`,
`models/basemodel.py:366`) is obtained from the same restatement.

`Prec` controls the bf16-emulating mode: the CUDA path stores activations,
activation gradients and the weight shadow copies in bf16 and accumulates in
fp32; `Prec(bf16=True)` rounds at those same storage points so the remaining
GPU/oracle difference is fp32 summation order only.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# precision emulation
# --------------------------------------------------------------------------
def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _RoundSTE(torch.autograd.Function):
    """value -> bf16-rounded value; gradient passes straight through."""

    @staticmethod
    def forward(ctx, x):
        return bf16_round(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _GradRound(torch.autograd.Function):
    """identity forward; the incoming gradient is rounded to bf16 (the CUDA
    path stores every activation-gradient tensor in bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return bf16_round(g)


class Prec(object):
    def __init__(self, bf16=False):
        self.bf16 = bf16

    def act(self, x):
        """storage point of a forward activation (also gradient storage)."""
        if not self.bf16:
            return x
        return _GradRound.apply(_RoundSTE.apply(x))

    def wt(self, w):
        """bf16 shadow copy of an fp32 master weight (bias stays fp32)."""
        if not self.bf16:
            return w
        return _RoundSTE.apply(w)


FP32 = Prec(False)
BF16 = Prec(True)


# --------------------------------------------------------------------------
# padding arithmetic  [TF-sem 2,3,7]
# --------------------------------------------------------------------------
def same_pad(n, k, s):
    """TF SAME forward padding: out=ceil(n/s); extra goes bottom/right."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv_out_size(n, k, s, padding):
    if padding == 'SAME':
        return -(-n // s)
    return (n - k) // s + 1


def deconv_out_size(n, k, s, padding):
    """slim.convolution2d_transpose output size [TF-sem 2]."""
    if padding == 'SAME':
        return n * s
    return n * s + max(k - s, 0)


# --------------------------------------------------------------------------
# convolutions
# --------------------------------------------------------------------------
def conv2d(x, w, b=None, stride=1, padding='SAME'):
    """tf.nn.conv2d + bias_add.  x NHWC, w HWIO (`utils/ops.py:58-69`,
    slim.convolution2d sites e.g. `models/unet.py:111-117`)."""
    kh, kw = w.shape[0], w.shape[1]
    xn = x.permute(0, 3, 1, 2)
    if padding == 'SAME':
        pt, pb = same_pad(x.shape[1], kh, stride)
        pl, pr = same_pad(x.shape[2], kw, stride)
        xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), b, stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose(x, w, b=None, stride=2, padding='VALID'):
    """tf.nn.conv2d_transpose / slim.convolution2d_transpose.  w is
    [kh,kw,Cout,Cin] (`utils/ops.py:75-76`); the op is the input-gradient of
    the forward conv with the same padding [TF-sem 2]."""
    kh, kw = w.shape[0], w.shape[1]
    xn = x.permute(0, 3, 1, 2)
    y = F.conv_transpose2d(xn, w.permute(3, 2, 0, 1), None, stride=stride)
    H, W = x.shape[1], x.shape[2]
    oh, ow = deconv_out_size(H, kh, stride, padding), deconv_out_size(W, kw, stride, padding)
    if padding == 'SAME':
        ph = max(kh - stride, 0) // 2
        pw = max(kw - stride, 0) // 2
        y = y[:, :, ph:ph + oh, pw:pw + ow]
    else:
        # k < s would need trailing zeros; not used by the reference.
        assert y.shape[2] == oh and y.shape[3] == ow
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1)


# --------------------------------------------------------------------------
# pooling  [TF-sem 4]
# --------------------------------------------------------------------------
def max_pool_with_argmax(x, k=2, s=2):
    """slim.max_pool2d(x, k, stride=s, padding='VALID').  Returns (y, slot)
    where slot (uint8) = dy*k+dx of the FIRST maximal element in row-major
    window order — the element TF's MaxPoolGrad routes the gradient to."""
    N, H, W, C = x.shape
    xn = x.permute(0, 3, 1, 2)
    win = xn.unfold(2, k, s).unfold(3, k, s)            # N,C,Ho,Wo,k,k
    win = win.reshape(N, C, win.shape[2], win.shape[3], k * k)
    slot = torch.argmax(win.detach(), dim=-1, keepdim=True)   # first max on ties
    y = torch.gather(win, -1, slot).squeeze(-1)
    return y.permute(0, 2, 3, 1), slot.squeeze(-1).permute(0, 2, 3, 1).to(torch.uint8)


def max_pool(x, k=2, s=2):
    return max_pool_with_argmax(x, k, s)[0]


def argmax_slot_to_flat(slot, k, s, H, W):
    """Derive the flat NHWC element index ((n*H+y)*W+x)*C+c from the window
    slot (documented derivation)."""
    N, Ho, Wo, C = slot.shape
    sl = slot.to(torch.int64)
    oy = torch.arange(Ho).view(1, Ho, 1, 1) * s + sl // k
    ox = torch.arange(Wo).view(1, 1, Wo, 1) * s + sl % k
    n = torch.arange(N).view(N, 1, 1, 1)
    c = torch.arange(C).view(1, 1, 1, C)
    return ((n * H + oy) * W + ox) * C + c


# --------------------------------------------------------------------------
# crop / pad / resize  [TF-sem 7,8]
# --------------------------------------------------------------------------
def crop_or_pad(x, th, tw):
    """tf.image.resize_image_with_crop_or_pad (centre)."""
    N, H, W, C = x.shape
    if H > th:
        o = (H - th) // 2
        x = x[:, o:o + th]
    if W > tw:
        o = (W - tw) // 2
        x = x[:, :, o:o + tw]
    H, W = x.shape[1], x.shape[2]
    if H < th or W < tw:
        pt = (th - H) // 2 if H < th else 0
        pl = (tw - W) // 2 if W < tw else 0
        pb = th - H - pt if H < th else 0
        pr = tw - W - pl if W < tw else 0
        x = F.pad(x, (0, 0, pl, pr, pt, pb))
    return x


def _legacy_axis(n_in, n_out):
    scale = np.float32(n_in) / np.float32(n_out)
    src = np.arange(n_out, dtype=np.float32) * scale
    lo = np.floor(src).astype(np.int64)
    hi = np.minimum(np.ceil(src).astype(np.int64), n_in - 1)
    frac = (src - lo.astype(np.float32)).astype(np.float32)
    return torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy(frac)


def resize_bilinear(x, oh, ow):
    """tf.image.resize_bilinear, legacy align_corners=False (no half-pixel
    centres): src = dst * in/out (`models/deconvolution.py:163`)."""
    N, H, W, C = x.shape
    ylo, yhi, yf = _legacy_axis(H, oh)
    xlo, xhi, xf = _legacy_axis(W, ow)
    top, bot = x[:, ylo], x[:, yhi]
    xf_ = xf.view(1, 1, ow, 1)
    yf_ = yf.view(1, oh, 1, 1)
    t = top[:, :, xlo] + (top[:, :, xhi] - top[:, :, xlo]) * xf_
    b = bot[:, :, xlo] + (bot[:, :, xhi] - bot[:, :, xlo]) * xf_
    return t + (b - t) * yf_


# --------------------------------------------------------------------------
# bilinear transposed-conv filters (`utils/upsampling.py:6-46`, restated)
# --------------------------------------------------------------------------
def get_kernel_size(factor):
    return 2 * factor - factor % 2


def upsample_filt(size):
    factor = (size + 1) // 2
    center = factor - 1 if size % 2 == 1 else factor - 0.5
    og = np.ogrid[:size, :size]
    return (1 - abs(og[0] - center) / factor) * (1 - abs(og[1] - center) / factor)


def bilinear_upsample_weights(factor, number_of_classes):
    k = get_kernel_size(factor)
    w = np.zeros((k, k, number_of_classes, number_of_classes), dtype=np.float32)
    filt = upsample_filt(k)
    for i in range(number_of_classes):
        w[:, :, i, i] = filt
    return w


def bilinear_upsample(x, factor):
    """What `tf.nn.conv2d_transpose(x, bilinear_upsample_weights(f, C),
    [N,H*f,W*f,C], [1,f,f,1])` (SAME) computes (`models/fcn.py:142,163,171`)."""
    C = x.shape[-1]
    w = torch.from_numpy(bilinear_upsample_weights(factor, C))
    return conv2d_transpose(x, w, None, stride=factor, padding='SAME')


# --------------------------------------------------------------------------
# batch-norm / dropout  [TF-sem 5,6]
# --------------------------------------------------------------------------
def batch_norm(x, beta, moving_mean, moving_var, training=True, decay=0.999, eps=1e-3,
               gamma=None):
    """slim.batch_norm defaults: center only, biased batch variance, moving
    stats m <- m*decay + batch*(1-decay).  Returns (y, new_mean, new_var).
    `gamma` covers `utils/ops.py:35-49` (scale=True, decay .9, eps 1e-5)."""
    if training:
        mean = x.mean(dim=(0, 1, 2))
        var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
        new_mean = moving_mean * decay + mean.detach() * (1 - decay)
        new_var = moving_var * decay + var.detach() * (1 - decay)
    else:
        mean, var = moving_mean, moving_var
        new_mean, new_var = moving_mean, moving_var
    y = (x - mean) * torch.rsqrt(var + eps)
    if gamma is not None:
        y = y * gamma
    return y + beta, new_mean, new_var


_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = 0x9E3779B9
_PHILOX_W1 = 0xBB67AE85
_U32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 (Salmon et al., SC'11) vectorised over numpy uint64
    arrays holding 32-bit values.  Returns the four output words."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _U32 for c in (c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _U32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _U32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _U32, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _U32, lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def dropout_keep_mask(numel, seed, stream, keep_prob=0.5):
    """Keep-mask (bool) for `numel` elements.  Element e uses Philox counter
    (e//4 low, e//4 high, stream, 0), key (seed low, seed high), output word
    e%4; u = (word >> 8) * 2^-24; keep iff u < keep_prob.  TF's own RNG stream
    cannot be reproduced, so the mask is DEFINED by this function
    and the CUDA kernel implements the same function."""
    nq = (numel + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    w = philox4x32_10(q & _U32, q >> np.uint64(32), np.uint64(stream), np.uint64(0),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(w, axis=1).reshape(-1)[:numel]
    u = (words >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)
    return u < np.float32(keep_prob)


def dropout(x, seed, stream, keep_prob=0.5):
    """slim.dropout(x, keep_prob=.5, is_training=True): x*keep_mask/keep
    (`models/deconvolution.py:129,144,154`) with the injected Philox mask."""
    m = torch.from_numpy(dropout_keep_mask(x.numel(), seed, stream, keep_prob)).view(x.shape)
    return x * m.to(x.dtype) * (1.0 / keep_prob)


# --------------------------------------------------------------------------
# loss / heads / optimizer  [TF-sem 9,10,11,13]
# --------------------------------------------------------------------------
def softmax_xent_mean(logits, labels_u8):
    """mean_{n,h,w} softmax_cross_entropy_with_logits(one_hot(mask), logits)
    (`models/basemodel.py:59-70` commented intent, `:194`, `:360`)."""
    C = logits.shape[-1]
    lab = labels_u8.reshape(-1).to(torch.int64)
    lg = logits.reshape(-1, C)
    lse = torch.logsumexp(lg, dim=-1)
    picked = lg.gather(1, lab.view(-1, 1)).squeeze(1)
    return (lse - picked).mean()


def sigmoid_argmax(logits):
    """`models/unet.py:76-78`: y_hat_sig = sigmoid(y_hat); output =
    float32(expand_dims(argmax(y_hat_sig, 3), -1)); first index on ties —
    ties DO occur once fp32 sigmoid saturates to 1.0."""
    sig = torch.sigmoid(logits)
    lab = torch.argmax(sig, dim=-1, keepdim=True).to(torch.float32)
    return sig, lab


def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (`models/basemodel.py:321`): epsilon sits
    OUTSIDE the bias correction.  `step` is the 1-based step count."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    p = p - lr_t * m / (torch.sqrt(v) + eps)
    return p, m, v


def xavier_uniform(shape, gen):
    """tf.contrib.layers.xavier_initializer() (uniform) [TF-sem 12]:
    fan_in = prod(shape[:-2])*shape[-2], fan_out = prod(shape[:-2])*shape[-1]."""
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in, fan_out = rf * shape[-2], rf * shape[-1]
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (gen.random(shape, dtype=np.float32) * 2 - 1) * np.float32(limit)
