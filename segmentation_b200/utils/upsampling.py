"""Host mirror of /root/reference/utils/upsampling.py:6-46 (bilinear transposed-
conv filter bank used by the FCN decoders).  Same names, arguments and float32
output; checked against golden values produced by the reference's own module
(tests/golden/upsampling.npz).

On the B200 path the `[k,k,C,C]` bank is never materialised for compute: it is
channel-diagonal, so `seg_bilinear_upsample_fwd` applies the k-tap separable
filter depthwise (include/segb200.h).  This function exists for API parity and
for the tests.
"""
import numpy as np


def get_kernel_size(factor):
    """Kernel size of the transposed conv for an upsampling `factor`."""
    return 2 * factor - factor % 2


def upsample_filt(size):
    """2-D bilinear kernel of the given size (float64, like the reference)."""
    factor = (size + 1) // 2
    center = factor - 1 if size % 2 == 1 else factor - 0.5
    rows, cols = np.ogrid[:size, :size]
    return (1 - abs(rows - center) / factor) * (1 - abs(cols - center) / factor)


def bilinear_upsample_weights(factor, number_of_classes):
    """[k,k,C,C] float32 weights, non-zero only on the channel diagonal."""
    k = get_kernel_size(factor)
    weights = np.zeros((k, k, number_of_classes, number_of_classes), dtype=np.float32)
    kernel = upsample_filt(k)
    idx = np.arange(number_of_classes)
    weights[:, :, idx, idx] = kernel[:, :, None]
    return weights
