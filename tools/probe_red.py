"""fp32 global-reduction throughput: LSU red.v4 vs TMA bulk reduce-add (csrc/probe.cu
probe_red_kernel), for the weight-gradient pattern: 148 CTAs x 36864 floats."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_b200 import native as N
res = []
def run(mode, ctas, elems, regions, op_bytes=256):
    dst = torch.zeros(regions * elems, dtype=torch.float32, device='cuda')
    out = torch.zeros(2 * ctas, dtype=torch.int64, device='cuda')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        dst.zero_()
        torch.cuda.synchronize()
        e0.record()
        N.call_probe('seg_probe_red_rate', mode, ctas, elems, regions, op_bytes, N.ptr(dst), N.ptr(out), N.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
    o = out.cpu().view(ctas, 2).double()
    expect = float(ctas) / regions if ctas % regions == 0 else None
    ok = expect is None or bool((dst == expect).all())
    r = {'mode': mode, 'ctas': ctas, 'elems': elems, 'regions': regions, 'op_bytes': op_bytes,
         'us': us, 'issue_cyc': float(o[:, 0].max()), 'done_cyc': float(o[:, 1].max()), 'correct': ok,
         'GBps': ctas * elems * 4 / us / 1e3}
    res.append(r); print(r, flush=True)
E = 36864
for regions in (1, 2, 16, 148):
    run(0, 148, E, regions)
    run(1, 148, E, regions)
    run(2, 148, E, regions)
    for ob in (1024, 4096, 16384):
        run(3, 148, E, regions, ob)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/probe_red.json', 'w'))
