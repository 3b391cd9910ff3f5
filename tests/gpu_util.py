"""Helpers for the -m gpu parity tests: move CPU tensors to the device, call the C
ABI (through segmentation_b200.native), bring results back."""
import ctypes
import json
import os

import numpy as np
import torch

from segmentation_b200 import native as N

BF16 = torch.bfloat16
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bfr(x):
    """round to bf16, keep fp32 (CPU)."""
    return x.to(BF16).to(torch.float32)


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.norm(a - b) / (torch.norm(b) + 1e-30))


def dev_bf16(x, cpad=None):
    """CPU fp32 NHWC -> device bf16 with channels zero-padded to cpad."""
    x = x.to(BF16)
    if cpad is not None and cpad != x.shape[-1]:
        pad = torch.zeros(x.shape[:-1] + (cpad,), dtype=BF16)
        pad[..., :x.shape[-1]] = x
        x = pad
    return x.cuda().contiguous()


def pad16(c):
    return (c + 15) // 16 * 16


def shadow_conv(w, cin_pad, cout_pad):
    """HWIO fp32 -> padded bf16 device shadow."""
    kh, kw, ci, co = w.shape
    s = torch.zeros(kh, kw, cin_pad, cout_pad, dtype=BF16)
    s[:, :, :ci, :co] = w.to(BF16)
    return s.cuda()


def shadow_deconv(w, cin_pad, cout_pad):
    """HWOI fp32 [kh,kw,cout,cin] -> padded bf16 device shadow."""
    kh, kw, co, ci = w.shape
    s = torch.zeros(kh, kw, cout_pad, cin_pad, dtype=BF16)
    s[:, :, :co, :ci] = w.to(BF16)
    return s.cuda()


def desc(k, stride, pads, cin, cout, cin_pad, cout_pad, flags, impl):
    pt, pl, pb, pr = pads
    return N.SegConvDesc(k, k, stride, pt, pl, pb, pr, cin, cout, cin_pad, cout_pad, flags, impl)


def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv_pads(h, w, k, s, padding):
    if padding == 'VALID':
        return (0, 0, 0, 0)
    pt, pb = same_pad(h, k, s)
    pl, pr = same_pad(w, k, s)
    return (pt, pl, pb, pr)


def sync():
    torch.cuda.synchronize()


def report(name, payload):
    """Append a diagnostic record to gpurun_out/diag.jsonl (travels back from the box)."""
    d = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, 'diag.jsonl'), 'a') as f:
        f.write(json.dumps({'test': name, **payload}) + '\n')
