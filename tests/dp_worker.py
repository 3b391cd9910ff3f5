"""Worker of tests/test_gpu_dp.py (one process per GPU, launched by torch.distributed.run):
the data-parallel U-Net step on W ranks x batch B/W must equal the single-GPU step at
batch B (SURVEY §4 tier 3, §8e).  Every rank runs the single-GPU reference itself (same
seed, same full batch) and compares in-process; rank 0 writes the verdict as JSON."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


class SliceDataSet(object):
    use_feed, has_masks = False, True

    def __init__(self, batches, lo, hi):
        self.batches, self.lo, self.hi, self.i = batches, lo, hi, 0
        self.batch_size = hi - lo

    def set_tf_sess(self, s):
        pass

    def next_batch(self):
        x, y = self.batches[self.i % len(self.batches)]
        self.i += 1
        return x[self.lo:self.hi], y[self.lo:self.hi]


def main():
    out_path = sys.argv[1]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from segmentation_b200 import parallel
    from segmentation_b200.models.unet import UNetModel
    from gpu_util import rel_l2

    B, S, nk, steps = 16, 256, 32, 3
    per = B // world
    g = np.random.default_rng(7)
    batches = [(g.random((B, S, S, 3), dtype=np.float32),
                g.integers(0, 2, (B, S, S, 1)).astype(np.uint8)) for _ in range(steps)]

    def build(ds):
        return UNetModel(dataset=ds, n_classes=2, input_dims=S, n_kernels=nk, learning_rate=1e-4,
                         load_snapshot=False, save_dir=None, seed=0)

    # ---- single-GPU reference at the full batch: one fwd+bwd (gradients), then `steps` steps
    ref = build(SliceDataSet(batches, 0, B))
    ex = ref._get_exec(B, True)
    ex.stage(torch.from_numpy(batches[0][0]).cuda(), torch.from_numpy(batches[0][1]).cuda())
    ex.forward(); ex.loss(True); ex.backward()
    torch.cuda.synchronize()
    g_ref = ref.store.grad.clone()
    ref.store.grad.zero_()
    ref_losses = []
    for _ in range(steps):
        ref.train_step()
        ref_losses.append(ref.seg_loss_op)
    p_ref = ref.store.master.clone()
    p0 = build(SliceDataSet(batches, 0, B)).store.master.clone()

    # ---- data-parallel: rank r takes images [r*per, (r+1)*per)
    dp = build(SliceDataSet(batches, rank * per, (rank + 1) * per))
    wrap = parallel.DataParallel(dp)
    ex = dp._get_exec(per, True)
    xs, ys = batches[0][0][rank * per:(rank + 1) * per], batches[0][1][rank * per:(rank + 1) * per]
    ex.stage(torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda())
    ex.forward(); ex.loss(True); ex.backward()
    for i in range(len(dp.opt_groups)):
        wrap.allreduce_group(i)
    torch.cuda.synchronize()
    g_dp = dp.store.grad.clone() / world            # the 1/W the Adam kernel folds in
    dp.store.grad.zero_()
    dp_losses = []
    for _ in range(steps):                          # eager, graph capture (NCCL inside), replay
        dp.train_step()
        dp_losses.append(dp.seg_loss_op)
    t = torch.tensor(dp_losses, dtype=torch.float64, device='cuda')
    dist.all_reduce(t)                              # mean over ranks == loss of the full batch
    dp_losses = (t / world).tolist()
    p_dp = dp.store.master.clone()
    # every rank must hold the same parameters after the steps
    pmax, pmin = p_dp.clone(), p_dp.clone()
    dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
    du, dr = (p_dp - p0).double(), (p_ref - p0).double()
    res = {
        'world': world, 'graph_captured': dp._last_train_exec.graph is not None,
        'grad_rel_l2': rel_l2(g_dp.cpu(), g_ref.cpu()),
        'grad_max_abs': float((g_dp - g_ref).abs().max()), 'grad_ref_max': float(g_ref.abs().max()),
        'losses_dp': dp_losses, 'losses_ref': ref_losses,
        'ranks_identical': bool(torch.equal(pmax, pmin)),
        'update_cosine': float((du * dr).sum() / (du.norm() * dr.norm())),
        'update_norm_ratio': float(du.norm() / dr.norm()),
        'param_rel_l2': rel_l2(p_dp.cpu(), p_ref.cpu()),
    }
    if rank == 0:
        with open(out_path, 'w') as f:
            json.dump(res, f)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)       # no destroy_process_group under captured graphs (see bench.py)


if __name__ == '__main__':
    main()
