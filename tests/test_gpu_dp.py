"""-m gpu: N-rank data-parallel step == 1-rank step at N x batch, through the real path —
NCCL all-reduce per optimizer group captured inside the step's CUDA graph
(segmentation_b200/parallel.py, models/basemodel.py group_ready).  Needs >= 2 GPUs on the
box (skipped otherwise; run with `gpurun --gpus 2`).  Worker: tests/dp_worker.py."""
import json
import os
import socket
import subprocess
import sys

import pytest

from gpu_util import report

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_step_equals_single_rank_step(cuda, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = str(tmp_path / 'dp.json')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'dp_worker.py'), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    report('dp2_equivalence', res)
    assert res['graph_captured']
    assert res['ranks_identical']
    # gradients: 2 x batch-8 mean-loss gradients averaged == batch-16 gradient.  The 1/pixels
    # factors differ by exactly 2 (a power of two: bf16 rounding is scale invariant), so only
    # fp32 summation order (unordered L2 reductions, NCCL ring) separates the two
    assert res['grad_rel_l2'] < 2e-4, res
    for a, b in zip(res['losses_dp'], res['losses_ref']):
        assert abs(a - b) < 2e-4, res
    assert res['update_cosine'] > 0.98 and abs(res['update_norm_ratio'] - 1) < 0.02, res
