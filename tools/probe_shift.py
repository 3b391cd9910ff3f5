"""Row-shifted UMMA descriptor experiment (see csrc/probe.cu)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_b200 import native as N
res = []
g = torch.Generator().manual_seed(0)
for K in (64, 32, 16):
    a = torch.randn(160, K, generator=g).to(torch.bfloat16)
    b = torch.randn(64, K, generator=g).to(torch.bfloat16)
    a_d, b_d = a.cuda(), b.cuda()
    for use_bo in (0, 1):
        for shift in (0, 1, 2, 3, 5, 7, 8, 9, 16, 17, 30):
            d = torch.full((128, 64), float('nan'), device='cuda')
            mode = 0x100 | (use_bo << 9) | (shift << 16)
            N.call_probe('seg_probe_umma', mode, 128, 64, K, N.ptr(a_d), N.ptr(b_d), N.ptr(d), N.stream_ptr())
            torch.cuda.synchronize()
            ref = a[shift:shift + 128].float() @ b.float().t()
            err = float((d.cpu() - ref).norm() / ref.norm())
            res.append({'K': K, 'bo': use_bo, 'shift': shift, 'err': err})
            print(K, use_bo, shift, '%.3e' % err, flush=True)
# two-box loads: second box at a smem offset that is not a multiple of 8 rows
for K in (64, 32, 16):
    a = torch.randn(160, K, generator=g).to(torch.bfloat16)
    b = torch.randn(64, K, generator=g).to(torch.bfloat16)
    a_d, b_d = a.cuda(), b.cuda()
    for split in (8, 13, 61, 100):
        for shift in (0, 3, 17):
            d = torch.full((128, 64), float('nan'), device='cuda')
            mode = 0x100 | (shift << 16) | (split << 24)
            N.call_probe('seg_probe_umma', mode, 128, 64, K, N.ptr(a_d), N.ptr(b_d), N.ptr(d), N.stream_ptr())
            torch.cuda.synchronize()
            ref = a[shift:shift + 128].float() @ b.float().t()
            err = float((d.cpu() - ref).norm() / ref.norm())
            res.append({'K': K, 'split': split, 'shift': shift, 'err': err})
            print('split', K, split, shift, '%.3e' % err, flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/probe_shift.json', 'w'))
