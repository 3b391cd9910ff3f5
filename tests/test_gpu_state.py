"""-m gpu tests of the training-loop state around the hot path (SURVEY §8 N1 and the
round-1 advisor findings):

  * snapshot() -> a new model restoring from it: identical state_dict, Adam slots, step
    counters and next-step loss (/root/reference/models/basemodel.py:112-136,494-501)
  * restore is atomic: a mismatched snapshot leaves the fresh model untouched
  * test() == the oracle's loss on the test batch (`:506-518`)
  * the JSONL loss log under log_dir
  * batched dropout (seg_dropout_ex) bit-exact against the oracle's Philox definition, per
    image streams and the device-side step offset
  * bayesian DeconvModel training through train_step() (graph capture + replay): fresh
    dropout masks every step, tracking the oracle
  * seg_loss_op follows the executor of the most recent train_step
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nets, tf_ops as T
from segmentation_b200 import engine as E

from gpu_util import bfr, report, sync
from test_gpu_unet import FeedDataSet

pytestmark = pytest.mark.gpu


def _unet(ds, save_dir, **kw):
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.unet import UNetModel
    return UNetModel(dataset=ds, n_classes=2, input_dims=188, n_kernels=16, learning_rate=1e-3,
                     save_dir=save_dir, **kw)


def test_snapshot_restore_round_trip(cuda, tmp_path):
    sd_dir = str(tmp_path / 'snap')
    ds = FeedDataSet(2, 188, 188, seed=1)
    a = _unet(ds, sd_dir, load_snapshot=False)
    for _ in range(3):
        a.train_step()
    path = a.snapshot()
    assert os.path.exists(path) and path.endswith('unet.ckpt-3.npz')
    a.train_step()
    a.snapshot()                                   # max_to_keep = 1
    assert sorted(os.listdir(sd_dir)) == ['unet.ckpt-4.npz']
    ds_b = FeedDataSet(2, 188, 188, seed=2)
    b = _unet(ds_b, sd_dir, load_snapshot=True, seed=99)     # different init, then restore
    assert b.global_step == 4 and b.store.step == a.store.step
    sa, sb = a.store.state_dict(), b.store.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    assert torch.equal(a.store.m, b.store.m) and torch.equal(a.store.v, b.store.v)
    assert torch.equal(a.store.shadow, b.store.shadow)
    # the next step on the same batch: same loss, same Adam bias correction
    batch = FeedDataSet(2, 188, 188, seed=3).next_batch()
    a.train_step(batch)
    b.train_step(batch)
    assert abs(a.seg_loss_op - b.seg_loss_op) < 1e-5, (a.seg_loss_op, b.seg_loss_op)
    assert b.global_step == 5
    # explicit file (load_snapshot_from)
    p5 = a.snapshot()
    c = _unet(FeedDataSet(2, 188, 188), None, load_snapshot=True, load_snapshot_from=p5, seed=7)
    assert c.global_step == 5
    with pytest.raises(Exception):
        c.snapshot()                               # save_dir=None: refuse, touch nothing


def test_restore_is_atomic_on_mismatch(cuda, tmp_path):
    sd_dir = str(tmp_path / 'snap')
    ds = FeedDataSet(2, 188, 188)
    a = _unet(ds, sd_dir, load_snapshot=False)
    a.train_step()
    a.snapshot()
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.unet import UNetModel
    fresh = UNetModel(dataset=ds, n_classes=2, input_dims=188, n_kernels=32, save_dir=None,
                      load_snapshot=False, seed=5)
    ref = fresh.store.state_dict()
    # n_kernels 32 model pointed at the n_kernels 16 snapshot: every shape mismatches
    b = UNetModel(dataset=ds, n_classes=2, input_dims=188, n_kernels=32, save_dir=sd_dir,
                  load_snapshot=True, seed=5)
    assert b.global_step == 0 and b.store.step == 0
    got = b.store.state_dict()
    for k in ref:
        assert np.array_equal(ref[k], got[k]), k
    assert float(b.store.m.abs().max()) == 0.0


def test_test_loss_matches_oracle_and_log(cuda, tmp_path):
    ds = FeedDataSet(2, 188, 188, seed=1)
    tds = FeedDataSet(2, 188, 188, seed=9)
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.unet import UNetModel
    log_dir = str(tmp_path / 'logs')
    m = UNetModel(dataset=ds, test_dataset=tds, n_classes=2, input_dims=188, n_kernels=16,
                  learning_rate=1e-3, save_dir=None, load_snapshot=False, log_dir=log_dir)
    p = nets.unet_params(n_kernels=16, n_classes=2, seed=3)
    m.load_weights({k: v.numpy() for k, v in p.items()})
    before = m.store.state_dict()
    loss = m.test()
    x, y = FeedDataSet(2, 188, 188, seed=9).next_batch()
    ref = float(T.softmax_xent_mean(
        nets.unet_forward(p, torch.from_numpy(x), prec=T.BF16),
        T.crop_or_pad(torch.from_numpy(y), 4, 4)))
    assert abs(loss - ref) < 2e-3, (loss, ref)
    after = m.store.state_dict()
    for k in before:                               # test() does not touch the parameters
        assert np.array_equal(before[k], after[k])
    assert m.global_step == 0
    m.summary_iter = 2
    for _ in range(5):
        m.train_step()
    lines = [json.loads(l) for l in open(os.path.join(log_dir, 'train_log.jsonl'))]
    assert [l['step'] for l in lines] == [2, 4]
    assert all(np.isfinite(l['loss']) and l['loss'] > 0 for l in lines)


def test_dropout_ex_bit_exact(cuda):
    g = torch.Generator().manual_seed(0)
    x = bfr(torch.randn(4, 9, 7, 16, generator=g))
    x_d = x.to(torch.bfloat16).cuda()
    seed = 0x1234567890ABCDEF
    # one stream over the batch == seg_dropout
    y_d = torch.empty_like(x_d)
    E.dropout_ex(x_d, y_d, seed, 5)
    ref = bfr(T.dropout(x, seed, 5))
    assert torch.equal(y_d.float().cpu(), ref)
    # per-image streams with a device-side step offset
    step = torch.tensor([3], dtype=torch.int32, device='cuda')
    E.dropout_ex(x_d, y_d, seed, 2, per_image_step=8, step_dev=step, step_mul=100)
    sync()
    for n in range(4):
        ref = bfr(T.dropout(x[n:n + 1], seed, 2 + 8 * n + 300))
        assert torch.equal(y_d[n:n + 1].float().cpu(), ref), n
    # in place
    E.dropout_ex(x_d, x_d, seed, 5)
    assert torch.equal(x_d.float().cpu(), bfr(T.dropout(x, seed, 5)))


def test_bayesian_deconv_train_steps_through_graph(cuda):
    """train_step() with no batch (the reference's call) on DeconvModel(bayesian=True): the
    step is captured into a CUDA graph on its second call; dropout streams follow
    global_step through a device scalar, so every replay draws fresh masks and the losses
    track the oracle's, which uses dropout=(seed, step)."""
    os.environ['SEGB200_IMPL'] = 'umma'
    from segmentation_b200.models.deconvolution import DeconvModel
    B, S, nk = 2, 256, 16
    ds = FeedDataSet(B, S, S, seed=4)
    m = DeconvModel(dataset=ds, n_classes=2, input_dims=S, n_kernels=nk, bayesian=True,
                    learning_rate=1e-3, load_snapshot=False, save_dir=None)
    p = nets.deconv_params(n_kernels=nk, n_classes=2, seed=4)
    m.load_weights({k: v.numpy() for k, v in p.items()})
    state = nets.AdamState(p)
    ds_ref = FeedDataSet(B, S, S, seed=4)
    got, ref = [], []
    for it in range(4):
        m.train_step()
        got.append(m.seg_loss_op)
        x, y = ds_ref.next_batch()
        stats = {}
        fwd = lambda q, xx: nets.deconv_forward(q, xx, training=True, bayesian=True, prec=T.BF16,
                                                dropout=(0, it), new_stats=stats)
        ref.append(nets.train_step(fwd, p, state, torch.from_numpy(x), torch.from_numpy(y),
                                   lr=1e-3, crop_mask=False))
        p.update(stats)
    ex = m._last_train_exec
    assert ex.graph is not None                    # captured, not disabled
    report('bayesian_deconv_train', {'loss': got, 'loss_ref': ref})
    for a, b in zip(got, ref):
        assert abs(a - b) < 1e-2, (got, ref)
    # the masks really change between replays: same batch, consecutive steps, lr = 0
    m.learning_rate = 0.0
    batch = ds.next_batch()
    m.train_step(batch)
    d1 = ex.act['bn2'].clone()
    m.train_step(batch)
    d2 = ex.act['bn2'].clone()
    sync()
    z1, z2 = (d1 == 0), (d2 == 0)
    assert float((z1 != z2).float().mean()) > 0.2


def test_loss_follows_last_train_executor(cuda):
    ds = FeedDataSet(2, 188, 188, seed=1)
    m = _unet(ds, None, load_snapshot=False)
    m.train_step()
    l2 = m.seg_loss_op
    big = FeedDataSet(4, 188, 188, seed=2).next_batch()
    m.train_step(big)                              # batch size != dataset.batch_size
    l4 = m.seg_loss_op
    assert np.isfinite(l4) and l4 != l2
    ex4 = m._exec[(4, True)]
    assert l4 == ex4.loss_value(0)
