#!/usr/bin/env python
"""bench.py — U-Net 256x256 training throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input:
forward + softmax-xent + backward + Adam of UNetModel(n_kernels=32, 3->2 ch) at
256x256, 16 images per GPU (BASELINE.json configs[2] at N GPUs; weak scaling,
data-parallel gradient all-reduce over NCCL for N>1).

  value : img/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e   : img/s through UNetModel.train_step() — the reference's own no-argument call —
          with HOST (pinned) batches pulled from the dataset: every step issues one
          batch's H2D (images+masks) and reads the loss back (D2H) inside the timed region
  roofline     : the single heaviest conv-family launch of the step, against the tensor
                 peak or (arithmetic intensity below the ridge) the HBM peak
  cpu_baseline : the oracle (torch-CPU fp32 restatement of the reference graph —
                 the reference's TensorFlow path cannot run) on
                 the host cores, BASELINE configs[0] (batch 4), bounded sample
  --impl reference : that same CPU restatement as the reference arm
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAIN_GFLOP_PER_IMG = 23.360      # BASELINE.md §3 (fwd + dgrad + wgrad, conv1_1 has no dgrad)
S, NK, NCLS, BATCH = 256, 32, 2, 16


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'tf_burst': p['bf16_tflops'],
                'tf_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'src': 'fallback'}


_ALL_CORES = os.sched_getaffinity(0)


def bind_to_gpu_numa(gpu_index):
    """Pin this process to the CPU cores of the GPU's NUMA node before any pinned host memory
    is allocated (first-touch places the pages there), so H2D DMA does not cross sockets.
    Returns a short description for the bench line; silently does nothing where sysfs does
    not expose the topology (single node, containers without /sys access)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(gpu_index)
        bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bdf).read().strip())
        if node < 0:
            return 'numa_node=-1 (no binding)'
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 'numa_node=%d (no usable cores)' % node
        os.sched_setaffinity(0, cpus)
        return 'numa_node=%d, %d cores' % (node, len(cpus))
    except Exception as e:                       # noqa: BLE001 - best effort
        return 'unavailable (%s)' % type(e).__name__


class SyntheticDataSet(object):
    """images fp32 [B,256,256,3] ~ U[0,1), masks uint8 [B,256,256,1] ~ Bernoulli(.5)
    (reference utils/datasets.py:174-190 tensor contract); a small pool of pinned
    host batches is cycled."""
    use_feed, has_masks = False, True

    def __init__(self, batch_size, seed, pool=4, pinned=True):
        import numpy as np
        import torch
        self.batch_size = batch_size
        g = np.random.default_rng(seed)
        self.pool = []
        for _ in range(pool):
            x = torch.from_numpy(g.random((batch_size, S, S, 3), dtype=np.float32))
            y = torch.from_numpy(g.integers(0, 2, (batch_size, S, S, 1)).astype(np.uint8))
            if pinned:
                x, y = x.pin_memory(), y.pin_memory()
            self.pool.append((x, y))
        self.i = 0

    def set_tf_sess(self, sess):
        pass

    def next_batch(self):
        b = self.pool[self.i % len(self.pool)]
        self.i += 1
        return b


class ClockSampler(object):
    """SM clock + throttle reasons sampled DURING the timed regions by an in-process NVML
    thread (a polling `nvidia-smi -lms` child was measured to stall the host-synchronous
    e2e loop for milliseconds at a time); falls back to nvidia-smi if NVML is missing."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index, period=0.01):
        import threading
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.p = self.f = self.thread = None
        self._stop = threading.Event()
        try:
            import pynvml as nv
            import torch
            nv.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                h = nv.nvmlDeviceGetHandleByUUID(('GPU-' + uuid).encode())
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = (('hw_slowdown', nv.nvmlClocksEventReasonHwSlowdown),
                     ('hw_thermal_slowdown', nv.nvmlClocksEventReasonHwThermalSlowdown),
                     ('sw_thermal_slowdown', nv.nvmlClocksEventReasonSwThermalSlowdown),
                     ('sw_power_cap', nv.nvmlClocksEventReasonSwPowerCap))

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for n, bit in names:
                            if r & bit:
                                self.reasons.add(n)
                    except Exception:
                        pass
                    self._stop.wait(period)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
            try:
                self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                           '--format=csv,noheader,nounits', '-lms', '250'],
                                          stdout=self.f, stderr=subprocess.DEVNULL)
            except Exception:
                self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': self.sm_max, 'reasons': [], 'samples': 0}
        sm, reasons = self.sm, self.reasons
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            out['source'] = 'nvml thread'
        elif self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush()
            rows = [l.strip().split(', ') for l in open(self.f.name) if l.strip()]
            os.unlink(self.f.name)
            for r in rows:
                if len(r) < 9:
                    continue
                try:
                    sm.append(float(r[1]))
                    out['sm_max_mhz'] = float(r[2])
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                    'sw_power_cap'), r[5:9]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(name)
            out['source'] = 'nvidia-smi -lms 250'
        if sm:
            s2 = sorted(sm)
            out['samples'] = len(s2)
            # median of the upper half = clocks under load (idle samples sit at the bottom)
            upper = s2[len(s2) // 2:]
            out['sm_mhz'] = upper[len(upper) // 2]
        out['reasons'] = sorted(reasons)
        return out


def cpu_oracle_rate(steps, warmup, batch=4):
    """Oracle U-Net train steps on the host cores -> (img/s, cores, sample text)."""
    import numpy as np
    import torch
    from oracle import nets
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = nets.unet_params(n_kernels=NK, n_classes=NCLS, seed=0)
    st = nets.AdamState(p)
    g = np.random.default_rng(0)
    x = torch.from_numpy(g.random((batch, S, S, 3), dtype=np.float32))
    y = torch.from_numpy(g.integers(0, 2, (batch, S, S, 1)).astype(np.uint8))
    fwd = lambda q, xx: nets.unet_forward(q, xx)
    for _ in range(warmup):
        nets.train_step(fwd, p, st, x, y, lr=1e-4)
    times = []
    for _ in range(steps):
        t0 = time.time()
        nets.train_step(fwd, p, st, x, y, lr=1e-4)
        times.append(time.time() - t0)
    times.sort()
    med = times[len(times) // 2]
    sample = ('%d warm-up + %d timed fp32 train steps (fwd+bwd+Adam) of U-Net 256x256 nk32 at '
              'batch %d, torch-CPU restatement of the reference graph (not TensorFlow), median'
              % (warmup, steps, batch))
    return batch / med, cores, sample, med


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    rate, cores, sample, med = cpu_oracle_rate(steps, 1)
    line = {
        'impl': 'reference', 'metric': 'U-Net train img/s (256x256, bs16/GPU)', 'value': rate,
        'unit': 'img/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': 1,
        'ms_per_step': med * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'U-Net 256x256 nk32 3->2ch train step (fwd+bwd+Adam), CPU sample at '
                               'batch 4 of the bs16/GPU workload'},
        'cpu_baseline': {'value': rate, 'unit': 'img/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': rate, 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    _emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from segmentation_b200 import native as N
    from segmentation_b200.models.unet import UNetModel
    from segmentation_b200 import parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    ds = SyntheticDataSet(BATCH, seed=1000 + rank)
    model = UNetModel(dataset=ds, n_classes=NCLS, input_dims=S, n_kernels=NK, learning_rate=1e-4,
                      load_snapshot=False, save_dir=None, seed=0)
    if world > 1:
        parallel.DataParallel(model)
    ex = model._get_exec(BATCH, True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W, K = max(3, args.warmup), args.steps
    dev_batches = [(x.to(dev), y.to(dev)) for x, y in ds.pool]

    # ---- launches per step (eager first step), then graph capture in warm-up
    before = N.LAUNCHES
    model.train_step(dev_batches[0])
    launches_per_step = N.LAUNCHES - before
    for i in range(1, W):
        model.train_step(dev_batches[i % len(dev_batches)])

    # ---- value: inputs resident in HBM
    # (the sampler starts BEFORE the barrier: NVML initialisation takes tens of milliseconds
    # on rank 0 only, and the other ranks would wait for it inside the first all-reduce)
    clocks = ClockSampler(local) if rank == 0 else None
    # no cyclic-GC pauses inside the timed regions (the e2e loop is host-synchronous: a
    # collection of the model's object graph shows up as one multi-millisecond step)
    import gc
    gc.collect()
    gc.disable()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        model.train_step(dev_batches[i % len(dev_batches)])
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))

    # ---- e2e: the reference's own call, train_step() with no arguments: the model pulls
    # pinned host batches from the dataset; every step issues one batch's H2D copy (the
    # next step's, on a copy stream behind this step's kernels) and reads the loss back
    for i in range(2):
        model.train_step()
    barrier()
    e0.record()
    loss = 0.0
    step_ms = []
    for i in range(K):
        t0 = time.perf_counter()
        model.train_step()
        # every step's loss is copied to pinned host memory behind the step; the host
        # reads step i-1's while step i runs (no stream drain between steps), and the
        # last one after the loop
        loss = model.seg_loss_lagged
        step_ms.append((time.perf_counter() - t0) * 1e3)
    loss = model.seg_loss_op
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    gc.enable()
    clk = clocks.stop() if clocks is not None else None

    # ---- per-launch timeline (un-captured pass) -> dominant kernel roofline
    roof = None
    if rank == 0:
        ex.use_graph = False
        saved = ex.graph
        ex.graph = None
        saved_side, ex.use_side = ex.use_side, False    # one stream: clean per-launch times
        # single-rank pass: forward/loss/backward called directly issue neither collectives
        # nor optimizer launches (those belong to the train step)
        N.TIMELINE = []
        ex.stage(*dev_batches[0])
        for _ in range(3):                       # fwd + loss + bwd only: parameters untouched
            torch.cuda.synchronize()
            # a spin kernel first: the host enqueues the whole pass behind it, so every event
            # pair brackets back-to-back GPU execution (no host launch latency inside)
            torch.cuda._sleep(int(4e-3 * 1.9e9))
            N.TIMELINE.clear()
            ex.forward()
            ex.loss(True)
            ex.backward()
            model.store.grad.zero_()
        torch.cuda.synchronize()
        tl = [(n, tag, a.elapsed_time(b)) for (n, tag, a, b) in N.TIMELINE]
        N.TIMELINE = None
        out_dir = os.path.join(ROOT, 'gpurun_out')
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, 'timeline%s.json' % os.environ.get('SEGB200_TAG', '')), 'w') as f:
                json.dump(tl, f)
        ex.graph, ex.use_graph = saved, True
        ex.use_side = saved_side
        roof = dominant_kernel(tl, model, ex)

    if world > 1:
        dist.barrier()                           # every rank is done with its GPU work
        torch.cuda.synchronize()
    if rank != 0:
        # no destroy_process_group(): tearing NCCL down under captured graphs was seen to
        # hang; the process is done, leave at once
        sys.stderr.flush()
        os._exit(0)
    pk = peaks()
    imgs = BATCH * world * K
    value = imgs / (ms_dev * 1e-3)
    e2e = imgs / (ms_e2e * 1e-3)
    h2d = BATCH * S * S * 3 * 4 + BATCH * S * S
    if args.skip_cpu:
        cpu_rate, cores, sample = None, 0, 'skipped (--skip-cpu)'
    else:
        os.sched_setaffinity(0, _ALL_CORES)      # the CPU baseline gets every host core
        cpu_rate, cores, sample, _ = cpu_oracle_rate(2, 1)
    # executed FLOPs per image: conv1_2 on its skip window (fwd if enabled, bwd always)
    y0c, x0c, hc, wc = ex.crop[4]
    full = float(ex.act['conv1_2'].shape[1] * ex.act['conv1_2'].shape[2])
    c12 = 2.0 * full * NK * NK * 9 / 1e9                      # GFLOP per image, one pass
    saved = (2 + (1 if getattr(ex, 'c12_crop', False) else 0)) * c12 * (1.0 - hc * wc / full)
    exec_gflop = TRAIN_GFLOP_PER_IMG - saved
    line = {
        'metric': 'U-Net train img/s (256x256, bs16/GPU)', 'value': value, 'unit': 'img/s',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_dev / K,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic',
        'config': {'workload': 'U-Net (models/unet.py graph) 256x256 RGB, 2 classes, n_kernels 32, '
                               'batch 16/GPU, fwd+xent+bwd+Adam, bf16 compute fp32 accumulate',
                   'global_batch': BATCH * world, 'parallelism': 'dp%d' % world,
                   'l2': 'per-step working set (>1 GB activations+gradients) exceeds the 126 MB L2; '
                         '4 distinct input batches cycled',
                   'impl': 'umma' if model.impl == 0 else 'simt', 'cuda_graph': True,
                   'dead_code': 'conv1_2 is evaluated on the 72x72 window that feeds concat4 (its '
                                'only consumer): identical outputs, loss and gradients',
                   'cpu_affinity': affinity},
        'e2e': {'value': e2e, 'unit': 'img/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / K,
                'loss_read': 'every step (4-byte async D2H into pinned memory behind the step); '
                             'the host reads it one step behind the launch, the last one '
                             'inside the timed region',
                # host wall time of the individual steps (rank 0): a slow host<->device link
                # or a descheduled host thread shows here, not in the device-timed `value`
                'host_step_ms': {'median': sorted(step_ms)[len(step_ms) // 2],
                                 'p90': sorted(step_ms)[int(len(step_ms) * 0.9)],
                                 'max': max(step_ms),
                                 'argmax': step_ms.index(max(step_ms))}},
        'gpu_launches': launches_per_step * K,
        'launches_per_step': launches_per_step,
        # of_burst / of_sustained: the reference graph's algorithmic FLOPs.
        # executed_*: what this implementation launches - conv1_2 (forward and backward) is
        # evaluated on the 72x72 window that feeds concat4 only; the rest of its output has
        # no consumer in models/unet.py:118-120,159-161 and its gradient there is zero.
        'conv_tensor_frac': {'of_burst': value / world * TRAIN_GFLOP_PER_IMG / 1e3 / pk['tf_burst'],
                             'of_sustained': value / world * TRAIN_GFLOP_PER_IMG / 1e3 /
                             pk['tf_sustained'], 'peaks': pk['src'],
                             'executed_gflop_per_img': exec_gflop,
                             'executed_of_burst': value / world * exec_gflop / 1e3 / pk['tf_burst']},
        'roofline': roof,
        'cpu_baseline': {'value': cpu_rate, 'unit': 'img/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'clocks': clk,
        'loss': loss,
    }
    _emit(line)
    if world > 1:
        sys.stderr.flush()
        os._exit(0)


def conv_work(layer, x_shape, y_shape, which):
    """Algorithmic work of one conv-family launch.
    FLOPs: 2*N*Ho*Wo*Cout*Cin*kh*kw for a conv, 2*N*Hi*Wi*Cin*Cout*kh*kw for a transposed
    conv.  Bytes: each tensor the launch must read or write once, bf16, real channels:
    fwd x+y, dgrad dy+dx, wgrad x+dy (weights / weight gradients are negligible)."""
    k, n = layer.k, x_shape[0]
    px = y_shape[1] * y_shape[2] if layer.kind == 'conv' else x_shape[1] * x_shape[2]
    flops = 2.0 * n * px * layer.cout * layer.cin * k * k
    small = n * x_shape[1] * x_shape[2] * layer.cin * 2.0
    big = n * y_shape[1] * y_shape[2] * layer.cout * 2.0
    return flops, small + big


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures
# summarised in profiles/r01_ncu_kernels.md (bytes); keyed like roofline['kernel']
NCU_TRAFFIC = {'conv1_1 wgrad': 103.24e6, 'conv1_2 fwd': 89.37e6, 'conv2_2 wgrad': 64.24e6,
               'conv2_2 fwd': 32.68e6}


def dominant_kernel(timeline, model, ex):
    """The conv-family launch with the largest duration: achieved TFLOP/s against the
    measured bf16 peak, or - when its arithmetic intensity is below the ridge of the two
    measured peaks - achieved GB/s of algorithmic bytes against the measured HBM peak."""
    pk = peaks()
    shapes = {}
    A = ex.act
    x_in = (ex.B, ex.H, ex.W, model.input_channel)
    for name, layer in model.layers.items():
        if name == 'output':
            shapes[name] = (A['conv9_2'].shape, ex.logits.shape)
        elif name.startswith('upconv'):
            j = int(name[-1])
            below = 'conv%d_2' % (4 + j) if j > 1 else 'conv5_2'
            shapes[name] = (A[below].shape, A[name].shape)
        elif name == 'conv1_1':
            shapes[name] = (x_in, A[name].shape)
        elif name == 'conv1_2':
            shapes[name] = (A['conv1_1'].shape, A['conv1_2'].shape)
        else:
            ys = A[name].shape
            shapes[name] = ((ys[0], ys[1] + 2, ys[2] + 2, layer.cin), ys)
    kind_of = {'seg_conv2d_fwd': 'fwd', 'seg_conv2d_dgrad': 'dgrad', 'seg_conv2d_wgrad': 'wgrad',
               'seg_deconv2d_fwd': 'fwd', 'seg_deconv2d_dgrad': 'dgrad',
               'seg_deconv2d_wgrad': 'wgrad'}
    total = sum(t for _, _, t in timeline)
    per = []
    for fn, tag, ms in timeline:
        if fn not in kind_of or tag not in model.layers:
            continue
        xs, ys = shapes[tag]
        fl, by = conv_work(model.layers[tag], xs, ys, kind_of[fn])
        if tag == 'conv1_2' and (kind_of[fn] != 'fwd' or getattr(ex, 'c12_crop', False)):
            # runs on the 72x72 skip crop only (exact: no consumer / zero gradient outside)
            y0, x0, h, w = ex.crop[4]
            lay = model.layers[tag]
            fl = 2.0 * ys[0] * h * w * lay.cout * lay.cin * 9
            by = ys[0] * ((h + 2) * (w + 2) * lay.cin + h * w * lay.cout) * 2.0
        per.append((tag, kind_of[fn], ms, fl, by))
    if not per:
        return None
    tag, kind, ms, fl, by = max(per, key=lambda p: p[2])
    ridge = pk['tf_burst'] * 1e12 / (pk['hbm_gbs'] * 1e9)
    conv_ms = sum(p[2] for p in per)
    tf = fl / (ms * 1e-3) / 1e12
    gbs = by / (ms * 1e-3) / 1e9
    if fl / by >= ridge:
        out = {'bound': 'tensor', 'achieved': tf, 'peak': pk['tf_burst'], 'unit': 'TFLOP/s',
               'frac': tf / pk['tf_burst']}
    else:
        out = {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
               'frac': gbs / pk['hbm_gbs']}
    out.update({'kernel': '%s %s' % (tag, kind),
                'traffic': NCU_TRAFFIC.get('%s %s' % (tag, kind)), 'traffic_unit': 'bytes/launch',
                'launch_ms': ms,
                'flops_per_launch': fl, 'bytes_per_launch': by, 'flop_per_byte': fl / by,
                'tflops': tf, 'peak_source': pk['src'] + ' burst',
                'share_of_step': ms / total if total > 0 else None,
                'conv_family_share_of_step': conv_ms / total if total > 0 else None,
                'conv_family_tflops': sum(p[3] for p in per) / (conv_ms * 1e-3) / 1e12,
                'serial_step_ms': total,
                'top5': [{'kernel': '%s %s' % (p[0], p[1]), 'ms': p[2],
                          'tflops': p[3] / (p[2] * 1e-3) / 1e12,
                          'gbs': p[4] / (p[2] * 1e-3) / 1e9}
                         for p in sorted(per, key=lambda q: -q[2])[:5]]})
    return out


def _watchdog(seconds):
    """A hung collective must not hold the GPU box: hard-exit after `seconds`."""
    import threading

    def _kill():
        sys.stderr.write('bench.py watchdog: exceeded %d s, aborting\n' % seconds)
        sys.stderr.flush()
        os._exit(3)

    t = threading.Timer(seconds, _kill)
    t.daemon = True
    t.start()


_REAL_STDOUT = None


def _emit(line):
    """The one JSON line: written to the process's original stdout (fd 1 is redirected to
    stderr for the whole run so that NCCL banners and library prints cannot interleave)."""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    _watchdog(int(os.environ.get('SEGB200_BENCH_WATCHDOG', '420')))
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--skip-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
