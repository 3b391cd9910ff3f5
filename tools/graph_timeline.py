"""Concurrent kernel timeline of the graph-replayed U-Net train step (bench.py's workload),
taken with CUPTI through torch.profiler: every kernel of ONE replayed step with its stream,
start and duration, so that the critical path and the overlap between the main, weight-
gradient and optimizer streams can be read (ncu serialises launches; this does not).

    python tools/graph_timeline.py [--world N]   -> gpurun_out/graph_timeline[_rank].json + .txt

Numbers taken under a profiler are for SHAPE (who overlaps whom), never bench values."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from segmentation_b200 import parallel  # noqa: E402
from segmentation_b200.models.unet import UNetModel  # noqa: E402


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    ds = bench.SyntheticDataSet(bench.BATCH, seed=1000 + rank, pool=2, pinned=False)
    model = UNetModel(dataset=ds, n_classes=bench.NCLS, input_dims=bench.S, n_kernels=bench.NK,
                      learning_rate=1e-4, load_snapshot=False, save_dir=None, seed=0)
    if world > 1:
        parallel.DataParallel(model)
    batches = [(x.to(dev), y.to(dev)) for x, y in ds.pool]
    for i in range(8):
        model.train_step(batches[i % 2])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(6):
            model.train_step(batches[i % 2])
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type is not None and 'cuda' in str(e.device_type).lower():
            tr = e.time_range
            evs.append((tr.start, tr.end - tr.start, e.name))
    # stream ids are only in the chrome trace: export and read it back
    out = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    trace = os.path.join(out, 'graph_trace_%d.json' % rank)
    prof.export_chrome_trace(trace)
    tr = json.load(open(trace))
    ks = [e for e in tr['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
    ks.sort(key=lambda e: e['ts'])
    os.remove(trace)
    # split into steps at the pack kernel (first kernel of a step)
    starts = [i for i, e in enumerate(ks) if 'pack_input' in e['name'] or 'stage_input' in e['name']]
    if len(starts) < 5:
        print('could not find step boundaries (%d pack kernels)' % len(starts))
        starts = [0, len(ks)]
    a, b = starts[3], starts[4]
    step = ks[a:b]
    t0 = step[0]['ts']
    streams = sorted({e['args'].get('stream', -1) for e in step})
    rows = [{'name': e['name'][:60], 'stream': streams.index(e['args'].get('stream', -1)),
             'start_us': round(e['ts'] - t0, 2), 'dur_us': round(e['dur'], 2),
             'grid': e['args'].get('grid'), 'smem': e['args'].get('shared memory')} for e in step]
    tag = '' if world == 1 else '_w%d_r%d' % (world, rank)
    with open(os.path.join(out, 'graph_timeline%s.json' % tag), 'w') as f:
        json.dump(rows, f)
    end = max(r['start_us'] + r['dur_us'] for r in rows)
    nxt = ks[b]['ts'] - t0 if b < len(ks) else end
    with open(os.path.join(out, 'graph_timeline%s.txt' % tag), 'w') as f:
        f.write('one replayed step: %d kernels, %d streams, last kernel ends at %.1f us, next step starts '
                'at %.1f us\n' % (len(rows), len(streams), end, nxt))
        busy = {}
        for r in rows:
            busy[r['stream']] = busy.get(r['stream'], 0.0) + r['dur_us']
        f.write('busy us per stream: %s\n' % {k: round(v, 1) for k, v in busy.items()})
        f.write('%8s %8s %2s %-18s %s\n' % ('start', 'dur', 'st', 'grid', 'kernel'))
        for r in rows:
            f.write('%8.1f %8.1f %2d %-18s %s\n' % (r['start_us'], r['dur_us'], r['stream'],
                                                      str(r['grid']), r['name']))
    if rank == 0:
        print(open(os.path.join(out, 'graph_timeline%s.txt' % tag)).read()[:600])
    sys.stdout.flush()
    os._exit(0)          # no collective teardown under the profiler (it was seen to hang)


if __name__ == '__main__':
    main()
