"""Per-layer micro-benchmark of the U-Net conv launches (bs16, 256x256) + in-kernel
timeline of CTA 0 of the halo conv kernel.  python tools/layer_prof.py [layer ...]
(the in-kernel marks need a library built with -DSEGB200_KERNEL_PROF=1: SEGB200_KERNEL_PROF=1 python -m segmentation_b200.build)"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_b200 import native as N
from segmentation_b200.models.unet import UNetModel


class DS(object):
    batch_size, use_feed, has_masks = 16, False, True
    def set_tf_sess(self, s): pass
    def next_batch(self): return None


model = UNetModel(dataset=DS(), n_classes=2, input_dims=256, n_kernels=32, load_snapshot=False, save_dir=None)
ex = model._get_exec(16, True)
x = torch.rand(16, 256, 256, 3, device='cuda'); y = torch.randint(0, 2, (16, 256, 256, 1), dtype=torch.uint8, device='cuda')
ex.stage(x, y)
ex.forward(); ex.loss(True); ex.backward(); model.store.grad.zero_()
torch.cuda.synchronize()
L, A, G = model.layers, ex.act, ex.g
want = sys.argv[1:] or ['conv1_2', 'conv2_2', 'conv3_2', 'conv9_2']
prof = torch.zeros(4 * 16 * 4, dtype=torch.int64, device='cuda')
src_of = {'conv1_1': 'x16', 'conv1_2': 'conv1_1', 'conv2_2': 'conv2_1', 'conv3_2': 'conv3_1', 'conv9_2': 'conv9_1',
          'conv2_1': 'pool1', 'conv3_1': 'pool2', 'conv4_2': 'conv4_1', 'conv5_2': 'conv5_1'}
out = {}
for name in want:
    src = A[src_of[name]]
    def run_fwd(): L[name].forward(src, A[name], impl=N.IMPL_UMMA)
    def run_dgrad():
        d = L[name].desc(src.shape[1], src.shape[2], 0, N.IMPL_UMMA)
        N.call('seg_conv2d_dgrad', ctypes.byref(d), N.vref(G[name]), N.ptr(L[name].w.shadow()),
               N.vref(G[src_of[name]]) if src_of[name] in G else N.vref(torch.zeros_like(src)), None, N.vref(src), None, N.stream_ptr())
    for kind, fn in (('fwd', run_fwd), ('dgrad', run_dgrad)):
        if kind == 'dgrad' and name == 'conv1_2':
            continue
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / 20
        prof.zero_()
        # profiling build only (SEGB200_KERNEL_PROF=1 python -m segmentation_b200.build --force)
        N.load().seg_debug_prof_buffer(N.ptr(prof)); fn(); torch.cuda.synchronize()
        N.load().seg_debug_prof_buffer(None)
        pall = prof.cpu()
        taps = pall[3 * 64:].view(4, 16)
        p = pall[:3 * 64].view(3, 16, 4)
        t0 = int(p[p > 0].min()) if (p > 0).any() else 0
        rel = (p - t0).clamp(min=0)
        print('=== %s %s: %.1f us/launch' % (name, kind, us))
        for role, rn in enumerate(('producer[start,a_empty_ok,done]', 'mma[start,tempty_ok,a_full_ok,issued]', 'epilogue[start,tfull_ok,tmem_read,stored]')):
            print(' ', rn)
            for t in range(8):
                print('    tile %d:' % t, [int(v) for v in rel[role, t]])
        ee = pall[3 * 64:3 * 64 + 16]
        if (ee > 0).any():
            print('  epilogue tile 2, per TMEM load [buffer_free, ld_done, processed, stored]:',
                  [[int(v - t0) for v in ee[i * 4:i * 4 + 4]] for i in range(4) if ee[i * 4] > 0])
        out['%s_%s' % (name, kind)] = {'us': us, 'prof': rel.tolist()}
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/layer_prof.json', 'w'))
