"""tcgen05.mma rate experiment (csrc/probe.cu probe_rate_kernel): cycles per MMA for
aligned vs row-shifted A descriptors, per swizzle mode / N / operand major, alone and
under concurrent stress (TMA writes into smem, TMEM reads, global stores / loads)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_b200 import native as N
res = []
ITERS = 40
FLAGS = {'data': 1, 'tma': 2, 'tld': 4, 'stg': 8, 'ldg': 16, 'roll1': 32, 'roll4': 64, 'drain': 128}
def run(kc, bn, b_mn, wp, shifted, a_mn=0, ctas=1, stress=()):
    global ITERS
    out = torch.zeros(2 * ctas + 8192 * ctas + 64, dtype=torch.int64, device='cuda')
    fl = sum(FLAGS[s] for s in stress)
    for _ in range(2):
        N.call_probe('seg_probe_mma_rate', kc, bn, b_mn, wp, shifted, ITERS, a_mn | (fl << 4), ctas, N.ptr(out), N.stream_ptr())
    torch.cuda.synchronize()
    o = out[:2 * ctas].cpu().view(ctas, 2).double()
    n_mma = ITERS * 9 * (kc // 16)
    drain = out[2 * ctas + 64 + 2:2 * ctas + 64 + 6].cpu().tolist()
    r = {'drain_cyc_per_4KB_load': [d / 512.0 for d in drain] if 'drain' in stress else None, 'kc': kc, 'bn': bn, 'b_mn': b_mn, 'wp': wp, 'shifted': shifted, 'a_mn': a_mn, 'ctas': ctas, 'stress': '+'.join(stress),
         'issue_cyc_per_mma': float(o[:, 0].mean()) / n_mma, 'retire_cyc_per_mma': float(o[:, 1].mean()) / n_mma,
         'retire_max': float(o[:, 1].max()) / n_mma}
    res.append(r)
    print(r, flush=True)
if 'full' in sys.argv:
    for kc in (64, 32, 16):
        for bn in (32, 64, 128, 256):
            for b_mn in (0, 1):
                run(kc, bn, b_mn, 64, 0)
                run(kc, bn, b_mn, 125, 1)
if 'stress' in sys.argv:
  for kc, bn in ((64, 64), (32, 32), (64, 128)):
    for ctas in (1, 148):
        run(kc, bn, 1, 125, 1, ctas=ctas)
        for st in (('data',), ('tma',), ('tld',), ('stg',), ('ldg',), ('data', 'tma', 'tld', 'stg', 'ldg')):
            run(kc, bn, 1, 125, 1, ctas=ctas, stress=st)
if 'roll' in sys.argv:
  for kc, bn in ((64, 64), (32, 32), (64, 128), (64, 256), (16, 32)):
    run(kc, bn, 1, 125, 1)
    run(kc, bn, 1, 125, 1, stress=('roll1',))
    run(kc, bn, 1, 125, 1, stress=('roll4',))
    run(kc, bn, 1, 125, 1, stress=('roll4', 'tld', 'stg', 'ldg', 'tma'))
# TMEM drain rate with the tensor pipe idle (ITERS tiny) and busy
for kc, bn in ((64, 64), (32, 32), (64, 128), (64, 256)):
    ITERS = 1
    run(kc, bn, 1, 125, 1, stress=('drain',))
    ITERS = 400
    run(kc, bn, 1, 125, 1, stress=('drain',))
    ITERS = 40
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/probe_rate2.json', 'w'))
