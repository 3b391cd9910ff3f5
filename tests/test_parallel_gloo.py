"""CPU world_size-2 (gloo) test of the data-parallel gradient path: bucketed
asynchronous all-reduce over a flat gradient buffer + 1/W folded into Adam ==
the gradient of the 2x batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from segmentation_b200 import engine as E
    from segmentation_b200 import parallel as P
    st = E.ParamStore(torch.device('cpu'))
    gen = np.random.default_rng(0)
    names = ['conv1_1', 'conv2_1', 'conv5_1', 'conv5_2', 'upconv1', 'conv6_1', 'output']
    for n in names:
        E.ConvLayer(st, n, 'deconv' if n.startswith('up') else 'conv', 2 if n.startswith('up') else 3,
                    1, 'VALID', 16, 16, True, gen)
    st.finalize()
    g = torch.Generator().manual_seed(100 + rank)
    local = torch.randn(st.numel, generator=g)
    st.grad.copy_(local)
    bounds = P.bucket_boundaries(st, ('conv5_1', 'upconv1'))
    buckets = P.GradBuckets(st.grad, bounds)
    # backward order: decoder bucket first, encoder last
    for i in (2, 1, 0):
        buckets.launch(i)
    buckets.join()
    q.put((rank, bounds, st.grad.clone().numpy(), local.numpy()))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = res[0][3] + res[1][3]
    assert len(res[0][1]) == 4 and res[0][1][0] == 0
    for r in res:
        assert np.allclose(r[2], total, atol=1e-6)       # every rank holds the sum
    # folded 1/W (Adam grad_scale) gives the mean == gradient of the 2x batch mean loss
    assert np.allclose(res[0][2] * 0.5, total / 2)
