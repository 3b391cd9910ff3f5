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
  roofline     : the kernel FAMILY with the largest share of the step (tconv / hconv / igemm /
                 twgrad / wgrad / pool / adam ...), CUDA-event timed per launch in this run,
                 against the measured tensor peak or (arithmetic intensity below the ridge)
                 the measured HBM peak; `families` lists every family; `traffic` comes from
                 profiles/r02_traffic.json (tools/ncu_traffic.py) and is dropped when that
                 file was produced from different kernel sources
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
    """images fp32 [B,256,256,3] (reference utils/datasets.py:174-190 tensor contract):
    smooth random fields in [0,1) (a coarse 32x32 grid upsampled bilinearly + noise);
    masks uint8 [B,256,256,1] = (red channel > 0.5), i.e. a learnable per-pixel target, so
    that the loss the bench prints says something about the backward pass (with
    Bernoulli(.5) labels it sits at ln 2 whatever the gradients are).  A small pool of
    pinned host batches is cycled."""
    use_feed, has_masks = False, True

    def __init__(self, batch_size, seed, pool=4, pinned=True):
        import numpy as np
        import torch
        self.batch_size = batch_size
        g = np.random.default_rng(seed)
        self.pool = []          # fp32 images in [0,1] + {0,1} masks (the model's tensor contract)
        self.pool_u8 = []       # the same batches as raw uint8 images + 0/255 masks (what the
        #                         reference's files decode to, utils/datasets.py:160-179)
        self.serve_u8 = False
        for _ in range(pool):
            coarse = torch.from_numpy(g.random((batch_size, 3, S // 32, S // 32), dtype=np.float32))
            smooth = torch.nn.functional.interpolate(coarse, size=(S, S), mode='bilinear',
                                                     align_corners=False)
            noise = torch.from_numpy(g.random((batch_size, 3, S, S), dtype=np.float32))
            x8 = ((0.8 * smooth + 0.2 * noise).permute(0, 2, 3, 1) * 255.0).round().clamp(0, 255)
            x8 = x8.to(torch.uint8).contiguous()
            x = (x8.to(torch.float32) / 255.0).contiguous()      # == what the device computes
            y = (x[..., 0:1] > 0.5).to(torch.uint8).contiguous()
            y8 = (y * 255).contiguous()
            if pinned:
                x, y, x8, y8 = x.pin_memory(), y.pin_memory(), x8.pin_memory(), y8.pin_memory()
            self.pool.append((x, y))
            self.pool_u8.append((x8, y8))
        self.i = 0

    def set_tf_sess(self, sess):
        pass

    def next_batch(self):
        pool = self.pool_u8 if self.serve_u8 else self.pool
        b = pool[self.i % len(pool)]
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


def parity_check(dev):
    """Correctness signal of the run (untimed; part of the cpu_baseline leg, the one place
    besides tests/ and smoke() that may execute oracle/): BASELINE config 1 (batch 4) -
    the oracle's parameters and one synthetic batch go through ONE fwd + loss + bwd of the
    CUDA path and of the bf16-emulating oracle; loss and the flat parameter gradient are
    compared.  Tolerances are the ones tests/test_gpu_unet.py states."""
    import numpy as np
    import torch
    from oracle import nets, tf_ops as T
    from segmentation_b200.models.unet import UNetModel

    class DS(object):
        batch_size, use_feed, has_masks = 4, False, True

        def set_tf_sess(self, s):
            pass

    ds = SyntheticDataSet(4, seed=123, pool=1, pinned=False)
    x, y = ds.pool[0]
    model = UNetModel(dataset=DS(), n_classes=NCLS, input_dims=S, n_kernels=NK, learning_rate=1e-4,
                      load_snapshot=False, save_dir=None, seed=0)
    p = nets.unet_params(n_kernels=NK, n_classes=NCLS, seed=0)
    model.load_weights({k: v.numpy() for k, v in p.items()})
    ex = model._get_exec(4, True)
    ex.stage(x.to(dev), y.to(dev))
    ex.forward()
    ex.loss(True)
    ex.backward()
    torch.cuda.synchronize()
    loss = float(ex.loss_sum.item()) / ex.loss_pixels
    fwd = lambda q, xx: nets.unet_forward(q, xx, prec=T.BF16)
    loss_ref, logits_ref, grads = nets.loss_and_grads(fwd, p, x, y)
    names = nets.trainable_names(p)
    g_ref = torch.cat([grads[k].flatten() for k in names]).double()
    g_gpu = torch.cat([model.store.params[k].grad().flatten() for k in names]).double().cpu()
    cos = float((g_ref * g_gpu).sum() / (g_ref.norm() * g_gpu.norm()))
    rel = float((g_ref - g_gpu).norm() / g_ref.norm())
    lg = float((ex.logits.cpu().double() - logits_ref.double()).norm() / logits_ref.double().norm())
    ok = abs(loss - float(loss_ref)) < 2e-3 and lg < 1e-2 and cos > 0.995
    return {'config': 'U-Net 256x256 nk32 batch 4 (BASELINE configs[0]), one fwd+loss+bwd, '
                      'vs the bf16-emulating oracle', 'loss': loss, 'loss_oracle': float(loss_ref),
            'logits_rel_l2': lg, 'grad_cosine': cos, 'grad_rel_l2': rel,
            'grad_checksum': float(g_gpu.sum()), 'grad_checksum_oracle': float(g_ref.sum()),
            'pass': bool(ok)}


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
    loss_first = model.seg_loss_op
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
    # The dataset serves raw uint8 images + 0/255 masks (what the reference's image files
    # decode to): the /255 of utils/datasets.py:178 runs inside the step's staging launch, so a
    # step uploads 4.2 MB instead of 13.6 MB.  Same pixel values as the device-resident leg.
    ds.serve_u8 = os.environ.get('SEGB200_E2E_FP32', '0') != '1'
    ex._pf = None                                # drop any batch prefetched in the other format
    for i in range(3):
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
        for _ in range(3):
            torch.cuda.synchronize()
            # a spin kernel first: the host enqueues the whole pass behind it, so every event
            # pair brackets back-to-back GPU execution (no host launch latency inside)
            torch.cuda._sleep(int(4e-3 * 1.9e9))
            N.TIMELINE.clear()
            ex.forward()
            ex.loss(True)
            ex.backward()
            # the optimizer groups' Adam launches with lr_t = 0 (the timed regions are over;
            # the parameters stay where they are)
            for grp in model.opt_groups:
                N.set_tag('adam')
                model.store.adam_launch(0.0, chunk_range=grp['chunks'])
        torch.cuda.synchronize()
        tl = [(n, tag, a.elapsed_time(b), fam, fl, by) for (n, tag, a, b, fam, fl, by) in N.TIMELINE]
        N.TIMELINE = None
        out_dir = os.path.join(ROOT, 'gpurun_out')
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, 'timeline%s.json' % os.environ.get('SEGB200_TAG', '')), 'w') as f:
                json.dump(tl, f)
        ex.graph, ex.use_graph = saved, True
        ex.use_side = saved_side
        roof = family_roofline(tl)

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
    h2d = BATCH * S * S * 3 * (1 if ds.serve_u8 else 4) + BATCH * S * S
    if args.skip_cpu:
        cpu_rate, cores, sample, parity = None, 0, 'skipped (--skip-cpu)', None
    else:
        os.sched_setaffinity(0, _ALL_CORES)      # the CPU baseline gets every host core
        import torch as _t
        _t.set_num_threads(os.cpu_count() or 1)
        parity = parity_check(dev)
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
                   'labels': 'mask = (red channel > 0.5) of smooth random images',
                   'impl': 'umma' if model.impl == 0 else 'simt', 'cuda_graph': True,
                   'dead_code': 'conv1_2 is evaluated on the 72x72 window that feeds concat4 (its '
                                'only consumer): identical outputs, loss and gradients',
                   'cpu_affinity': affinity},
        'e2e': {'value': e2e, 'unit': 'img/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 8, 'ms_per_step': ms_e2e / K,
                'input': ('uint8 images + 0/255 masks from pinned host memory, /255 on the device '
                          '(utils/datasets.py:176-179)' if ds.serve_u8 else
                          'fp32 images + {0,1} masks from pinned host memory'),
                'loss_read': 'every step: the next step\'s staging launch stores {loss, step id} '
                             '(8 bytes) into pinned host memory; the host reads it one step '
                             'behind the launch, the last one inside the timed region',
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
        # labels are a function of the image (mask = red > 0.5): the loss starts at ~ln 2 and
        # must fall; parity_check compares loss + gradients with the oracle (config 1)
        'loss': loss, 'loss_first': loss_first, 'loss_last': loss,
        'train_steps_between': model.global_step - 1,
        'parity_check': parity,
    }
    _emit(line)
    if world > 1:
        sys.stderr.flush()
        os._exit(0)


TRAFFIC_FILE = os.path.join(ROOT, 'profiles', 'r02_traffic.json')


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu, one pass over one eager
    step: tools/ncu_traffic.py), keyed 'layer kind'.  Valid only for the kernel sources it
    was measured on: the file carries the digest of segmentation_b200/csrc + the build
    flags, and a mismatch drops the numbers (traffic: null) instead of reporting stale ones."""
    try:
        from segmentation_b200 import build as B
        t = json.load(open(TRAFFIC_FILE))
        if t.get('csrc_digest') != B._digest():
            return None, 'profiles/r02_traffic.json is stale (kernel sources changed since)'
        return t, 'profiles/r02_traffic.json (ncu, same kernel sources)'
    except Exception as e:                       # noqa: BLE001
        return None, 'no traffic profile (%s)' % type(e).__name__


KIND_OF = {'seg_conv2d_fwd': 'fwd', 'seg_conv2d_dgrad': 'dgrad', 'seg_conv2d_wgrad': 'wgrad',
           'seg_deconv2d_fwd': 'fwd', 'seg_deconv2d_dgrad': 'dgrad', 'seg_deconv2d_wgrad': 'wgrad',
           'seg_stem_fwd': 'fwd', 'seg_stem_wgrad': 'wgrad'}


def launch_key(fn, tag):
    return '%s %s' % (tag, KIND_OF.get(fn, fn.replace('seg_', '')))


def family_roofline(timeline):
    """Per kernel family of one serialised step (every launch CUDA-event timed in THIS run):
    time, share, algorithmic TFLOP/s and GB/s (engine.py notes the SURVEY-8d work of each
    call), and the fraction of the measured peak that bounds it - the tensor peak when the
    family's arithmetic intensity is above the ridge of the two measured peaks, else the HBM
    peak.  `kernel` / `achieved` / `frac` at the top level are those of the family with the
    largest share of the step."""
    pk = peaks()
    ridge = pk['tf_burst'] * 1e12 / (pk['hbm_gbs'] * 1e9)
    traffic, traffic_src = load_traffic()
    tmap = (traffic or {}).get('launches', {})
    total = sum(e[2] for e in timeline)
    fams, launches = {}, []
    for fn, tag, ms, fam, fl, by in timeline:
        f = fams.setdefault(fam, {'ms': 0.0, 'flops': 0.0, 'bytes': 0.0, 'launches': 0,
                                  'traffic': 0.0, 'traffic_known': True})
        f['ms'] += ms; f['flops'] += fl; f['bytes'] += by; f['launches'] += 1
        t = tmap.get(launch_key(fn, tag))
        if t is None:
            f['traffic_known'] = False
        else:
            f['traffic'] += t['dram_bytes']
        launches.append({'kernel': launch_key(fn, tag), 'family': fam, 'ms': ms,
                         'tflops': fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0,
                         'gbs': by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                         'flops': fl, 'bytes': by,
                         'traffic': t['dram_bytes'] if t is not None else None})

    def bound_of(fl, by, ms):
        tf = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        gbs = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        if by > 0 and fl / by >= ridge:
            return {'bound': 'tensor', 'achieved': tf, 'peak': pk['tf_burst'], 'unit': 'TFLOP/s',
                    'frac': tf / pk['tf_burst']}
        return {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                'frac': gbs / pk['hbm_gbs']}

    fam_list = []
    for name, f in fams.items():
        e = {'family': name, 'launches': f['launches'], 'us': f['ms'] * 1e3,
             'share': f['ms'] / total if total > 0 else None,
             'tflops': f['flops'] / (f['ms'] * 1e-3) / 1e12 if f['ms'] > 0 else 0.0,
             'gbs': f['bytes'] / (f['ms'] * 1e-3) / 1e9 if f['ms'] > 0 else 0.0,
             'flops': f['flops'], 'bytes': f['bytes'],
             'traffic': f['traffic'] if (traffic is not None and f['traffic_known']) else None}
        e.update(bound_of(f['flops'], f['bytes'], f['ms']))
        fam_list.append(e)
    fam_list.sort(key=lambda e: -e['us'])
    if not fam_list:
        return None
    top = fam_list[0]
    n_top = max(top['launches'], 1)
    out = {'bound': top['bound'], 'achieved': top['achieved'], 'peak': top['peak'],
           'unit': top['unit'], 'frac': top['frac'],
           # per launch, like `achieved` (family totals / its launch count)
           'traffic': top['traffic'] / n_top if top['traffic'] is not None else None,
           'traffic_unit': 'bytes/launch (family average)', 'traffic_source': traffic_src,
           'kernel': top['family'], 'kernel_launches': top['launches'],
           'launch_ms': top['us'] / 1e3 / n_top,
           'flops_per_launch': top['flops'] / n_top, 'bytes_per_launch': top['bytes'] / n_top,
           'share_of_step': top['share'], 'peak_source': pk['src'] + ' burst',
           'ridge_flop_per_byte': ridge, 'serial_step_ms': total,
           'timing': 'CUDA events around every launch of one eager, single-stream step (3rd of 3 '
                     'passes) in this run; event pairs add ~2 us per launch',
           'families': fam_list,
           'top5_launches': sorted(launches, key=lambda l: -l['ms'])[:5]}
    conv = [e for e in fam_list if e['family'] in ('tconv', 'hconv', 'igemm', 'twgrad', 'wgrad',
                                                   'stem', 'simt')]
    if conv:
        cms = sum(e['us'] for e in conv) / 1e3
        out['conv_family_share_of_step'] = cms / total
        out['conv_family_tflops'] = sum(e['flops'] for e in conv) / (cms * 1e-3) / 1e12
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
