"""Generate tests/golden/upsampling.npz by RUNNING the reference's own
`utils/upsampling.py` (the only hot-path code of the reference that is pure
numpy and therefore runnable here).  The reference is Python 2 (`xrange`,
/root/reference/utils/upsampling.py:42); we inject `xrange = range` into
builtins and import the module unmodified from /root/reference.

Run in the build container only (/root/reference does not exist on the GPU
box):   python tests/golden/make_golden.py
"""
import builtins
import importlib.util
import os

import numpy as np

REF = '/root/reference/utils/upsampling.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'upsampling.npz')


def main():
    builtins.xrange = range
    spec = importlib.util.spec_from_file_location('ref_upsampling', REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for f in (1, 2, 3, 4, 8, 16, 32):
        out['ksize_f%d' % f] = np.int64(mod.get_kernel_size(f))
    for k in (1, 2, 3, 4, 5, 16, 32, 64):
        out['filt_k%d' % k] = np.asarray(mod.upsample_filt(k), dtype=np.float64)
    for f, c in ((2, 2), (2, 21), (8, 21), (16, 3), (32, 2)):
        out['weights_f%d_c%d' % (f, c)] = mod.bilinear_upsample_weights(f, c)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, 'with', len(out), 'arrays')


if __name__ == '__main__':
    main()
