"""Independent naive (pure-loop numpy) restatement of the parity-critical ops
(TEST INFRASTRUCTURE ONLY).  Nothing external pins oracle.tf_ops (parity
unpinned, see oracle/__init__.py), so each op is written a second time from
the TF documentation with scalar loops and cross-checked on tiny shapes by
tests/test_oracle.py.  Small inputs only.
"""
import math

import numpy as np


def same_pad(n, k, s):
    out = int(math.ceil(n / float(s)))
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv2d(x, w, b, stride, padding):
    N, H, W, Ci = x.shape
    kh, kw, _, Co = w.shape
    if padding == 'SAME':
        pt, _ = same_pad(H, kh, stride)
        pl, _ = same_pad(W, kw, stride)
        Ho, Wo = -(-H // stride), -(-W // stride)
    else:
        pt = pl = 0
        Ho, Wo = (H - kh) // stride + 1, (W - kw) // stride + 1
    y = np.zeros((N, Ho, Wo, Co), np.float64)
    for n in range(N):
        for oy in range(Ho):
            for ox in range(Wo):
                for r in range(kh):
                    for s in range(kw):
                        iy, ix = oy * stride + r - pt, ox * stride + s - pl
                        if 0 <= iy < H and 0 <= ix < W:
                            y[n, oy, ox] += x[n, iy, ix].astype(np.float64) @ w[r, s].astype(np.float64)
    if b is not None:
        y += b
    return y.astype(np.float32)


def conv2d_transpose(x, w, b, stride, padding):
    """w [kh,kw,Cout,Cin]; scatter form of the conv input-gradient."""
    N, H, W, Ci = x.shape
    kh, kw, Co, _ = w.shape
    fh, fw = (H - 1) * stride + kh, (W - 1) * stride + kw
    full = np.zeros((N, fh, fw, Co), np.float64)
    for n in range(N):
        for iy in range(H):
            for ix in range(W):
                for r in range(kh):
                    for s in range(kw):
                        full[n, iy * stride + r, ix * stride + s] += \
                            w[r, s].astype(np.float64) @ x[n, iy, ix].astype(np.float64)
    if padding == 'SAME':
        ph, pw = max(kh - stride, 0) // 2, max(kw - stride, 0) // 2
        full = full[:, ph:ph + H * stride, pw:pw + W * stride]
    if b is not None:
        full = full + b
    return full.astype(np.float32)


def max_pool_with_argmax(x, k, s):
    N, H, W, C = x.shape
    Ho, Wo = (H - k) // s + 1, (W - k) // s + 1
    y = np.zeros((N, Ho, Wo, C), x.dtype)
    slot = np.zeros((N, Ho, Wo, C), np.uint8)
    for n in range(N):
        for oy in range(Ho):
            for ox in range(Wo):
                for c in range(C):
                    best, bi = None, 0
                    for dy in range(k):
                        for dx in range(k):
                            v = x[n, oy * s + dy, ox * s + dx, c]
                            if best is None or v > best:      # strict: first max wins
                                best, bi = v, dy * k + dx
                    y[n, oy, ox, c] = best
                    slot[n, oy, ox, c] = bi
    return y, slot


def resize_bilinear(x, oh, ow):
    N, H, W, C = x.shape
    y = np.zeros((N, oh, ow, C), np.float32)
    sy, sx = np.float32(H) / np.float32(oh), np.float32(W) / np.float32(ow)
    for oy in range(oh):
        fy = np.float32(oy) * sy
        y0 = int(math.floor(fy)); y1 = min(int(math.ceil(fy)), H - 1); ly = np.float32(fy - y0)
        for ox in range(ow):
            fx = np.float32(ox) * sx
            x0 = int(math.floor(fx)); x1 = min(int(math.ceil(fx)), W - 1); lx = np.float32(fx - x0)
            top = x[:, y0, x0] + (x[:, y0, x1] - x[:, y0, x0]) * lx
            bot = x[:, y1, x0] + (x[:, y1, x1] - x[:, y1, x0]) * lx
            y[:, oy, ox] = top + (bot - top) * ly
    return y


def softmax_xent_mean(logits, labels):
    lg = logits.reshape(-1, logits.shape[-1]).astype(np.float64)
    lab = labels.reshape(-1).astype(np.int64)
    tot = 0.0
    for i in range(lg.shape[0]):
        m = lg[i].max()
        lse = m + math.log(np.exp(lg[i] - m).sum())
        tot += lse - lg[i, lab[i]]
    return tot / lg.shape[0]


def philox4x32_10_scalar(ctr, key):
    """Scalar Philox-4x32-10 with python ints (reference for the vectorised
    oracle and for the known-answer test from the Random123 distribution)."""
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c
