"""Times the Adam launch alone on the U-Net parameter store."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segmentation_b200.models.unet import UNetModel
class DS(object):
    batch_size, use_feed, has_masks = 16, False, True
    def set_tf_sess(self, s): pass
    def next_batch(self): return None
model = UNetModel(dataset=DS(), n_classes=2, input_dims=256, n_kernels=32, load_snapshot=False, save_dir=None)
st = model.store
for kind in ('zeros', 'randn1e-3', 'randn1e-20'):
    if kind == 'randn1e-3': st.grad.normal_(0, 1e-3)
    if kind == 'randn1e-20': st.grad.normal_(0, 1e-20)
    for _ in range(3): st.adam_step(1e-4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        if kind != 'zeros': st.grad.normal_(0, float(kind[5:]))
        e0.record(); st.adam_step(1e-4); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    print(kind, 'adam us: min %.1f median %.1f' % (min(ts), sorted(ts)[5]), 'numel', st.numel, 'chunks', st.nchunks, flush=True)
