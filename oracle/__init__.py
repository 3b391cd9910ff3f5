"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A CPU restatement (torch-CPU fp32 + numpy) of the arithmetic the reference
`nathanin/segmentation` delegates to TensorFlow 1.x / tf.contrib.slim for its
segmentation hot path (U-Net, FCN-8/16/32s, conv/deconvolution model forward +
backward, softmax cross-entropy, Adam).

PARITY STATUS: **parity unpinned** for everything that lives in TensorFlow
(TF 1.x, version unpinned by the reference, is not vendored under
/root/reference and cannot be installed here; the reference has no tests,
fixtures or golden vectors).  The only functions pinned against the reference
itself are the pure-numpy helpers of `utils/upsampling.py`, which ARE imported
from /root/reference by `tests/golden/make_golden.py` to generate
`tests/golden/upsampling.npz`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product
(`segmentation_b200`) must never import it.
"""
