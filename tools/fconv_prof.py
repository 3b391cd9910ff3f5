"""First-layer kernel at the bench shape (16 x 256 x 256 RGB -> 32 ch), forward and weight
gradient, CUDA-event timed; run under ncu for the kernel's own counters:

  python tools/fconv_prof.py
  ncu --set full --import-source on --clock-control none -k regex:fconv -c 4 \
      -o gpurun_out/fconv python tools/fconv_prof.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from segmentation_b200 import engine as E  # noqa: E402
from segmentation_b200 import native as N  # noqa: E402

B, H, W, CO = 16, int(os.environ.get('S', 256)), int(os.environ.get('S', 256)), 32
PAD = os.environ.get('PAD', 'VALID')
torch.manual_seed(0)
x = torch.rand(B, H, W, 3, device='cuda')
x4 = torch.zeros(B, H, W, 4, dtype=torch.bfloat16, device='cuda')
E.stage_input(x, x4)
w = torch.zeros(3, 3, 16, CO, dtype=torch.bfloat16, device='cuda')
w[:, :, :3] = (torch.randn(3, 3, 3, CO, device='cuda') * 0.2).to(torch.bfloat16)
b = torch.randn(CO, device='cuda') * 0.1
p = 1 if PAD == 'SAME' else 0
Ho, Wo = H - 2 + 2 * p, W - 2 + 2 * p
y = torch.zeros(B, Ho, Wo, CO, dtype=torch.bfloat16, device='cuda')
dz = (torch.randn(B, Ho, Wo, CO, device='cuda') * 0.1).to(torch.bfloat16)
dw = torch.zeros(3, 3, 3, CO, device='cuda')
db = torch.zeros(CO, device='cuda')
d = N.SegConvDesc(3, 3, 1, p, p, p, p, 3, CO, 16, CO, N.EPI_BIAS | N.EPI_RELU, N.IMPL_UMMA)
st = N.stream_ptr()


def fwd():
    N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(x4), None, N.ptr(w), N.ptr(b), N.vref(y), st)


def wgrad():
    N.call('seg_conv2d_wgrad', ctypes.byref(d), N.vref(x4), None, N.vref(dz), N.ptr(dw), N.ptr(db), st)


for fn, name in ((fwd, 'fwd'), (wgrad, 'wgrad')):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    n = int(os.environ.get('REPS', 20))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(n):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    byt = B * Ho * Wo * CO * 2 + B * H * W * 8
    print('%s %s: median %.1f us  min %.1f us  -> %.0f GB/s (algorithmic %d MB)'
          % (name, PAD, ts[len(ts) // 2], ts[0], byt / ts[len(ts) // 2] / 1e3, byt >> 20))

# ---- fused with the 2x2 max-pool (seg_conv2d_pool_fwd / seg_conv2d_pool_wgrad)
if Ho % 2 == 0 and Wo % 2 == 0:
    pooled = torch.zeros(B, Ho // 2, Wo // 2, CO, dtype=torch.bfloat16, device='cuda')
    amax = torch.zeros(B, Ho // 2, Wo // 2, CO, dtype=torch.uint8, device='cuda')
    dpool = (torch.randn(B, Ho // 2, Wo // 2, CO, device='cuda') * 0.1).to(torch.bfloat16)
    wy, wx, wh, ww = (Ho - 74) // 2, (Wo - 74) // 2, 74, 74
    ywin = y[:, wy:wy + wh, wx:wx + ww, :]
    add = (torch.randn(B, wh, ww, CO, device='cuda') * 0.1).to(torch.bfloat16)

    def pfwd():
        N.call('seg_conv2d_pool_fwd', ctypes.byref(d), N.vref(x4), N.ptr(w), N.ptr(b), N.vref(ywin),
               wy, wx, N.vref(pooled), N.ptr(amax), st)

    def pwgrad():
        N.call('seg_conv2d_pool_wgrad', ctypes.byref(d), N.vref(x4), N.vref(dpool), N.ptr(amax),
               N.vref(pooled), N.ptr(dw), N.ptr(db), st)

    for fn, name in ((pfwd, 'pool fwd'), (pwgrad, 'pool wgrad')):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print('%s %s: median %.1f us  min %.1f us' % (name, PAD, ts[len(ts) // 2], ts[0]))
