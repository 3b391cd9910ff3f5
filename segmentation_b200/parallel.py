"""Data-parallel training over one 8xB200 box: one process per GPU, batch
partitioned across ranks, gradients summed with NCCL all-reduce over
NVLink/NVSwitch, bucketed in reverse layer order and overlapped with the rest
of the backward pass.  The reference has no distributed code at all
(SURVEY.md §2.1); images are independent in U-Net / FCN (no batch-coupled op),
and the loss is a mean over N*H*W (reference models/basemodel.py:360), so with
equal per-rank batches grad_global = (1/W) * sum_r grad_r: the 1/W is folded
into the Adam kernel (`grad_scale`).

The flat gradient buffer is laid out in TF variable order (forward order);
backward completes it from the back, so buckets are contiguous slices that
become ready one after the other; each is reduced and then updated by its own
Adam launch on the executor's optimizer stream while the backward pass goes on.
"""
import os

import torch
import torch.distributed as dist

from . import native as N


class GradBuckets(object):
    """Contiguous slices of a flat gradient buffer, all-reduced asynchronously
    as they become ready, joined before the optimizer."""

    def __init__(self, flat_grad, boundaries, group=None):
        """boundaries: ascending element offsets [0, b1, ..., numel]."""
        assert boundaries[0] == 0 and boundaries[-1] == flat_grad.numel()
        self.flat = flat_grad
        self.slices = [flat_grad[a:b] for a, b in zip(boundaries[:-1], boundaries[1:])]
        self.group = group
        self.pending = []

    def launch(self, idx):
        """All-reduce (sum) bucket idx; returns immediately."""
        w = dist.all_reduce(self.slices[idx], op=dist.ReduceOp.SUM, group=self.group,
                            async_op=True)
        self.pending.append(w)

    def join(self):
        for w in self.pending:
            w.wait()
        self.pending = []

    def allreduce_all(self):
        for i in reversed(range(len(self.slices))):
            self.launch(i)
        self.join()


def bucket_boundaries(store, first_layers):
    """Offsets that split the flat buffer at the first parameter of each named
    layer (in TF order)."""
    offs = [0]
    for name in first_layers:
        o = store.params[name + '/weights'].offset
        if o > offs[-1]:
            offs.append(o)
    offs.append(store.numel)
    return offs


class DataParallel(object):
    """Wraps a model: identical initial parameters on every rank, world_size for
    the Adam grad_scale, one gradient bucket per optimizer group of the model
    (U-Net: encoder | bottleneck conv5_* | decoder).  The executor calls
    `allreduce_group(i)` on its optimizer stream when group i's gradients are
    complete; the group's Adam launch follows on the same stream, so reduction and
    update of the back of the network overlap the backward pass of the front."""

    def __init__(self, model, group=None):
        from . import engine as E
        bn = [n for n, l in getattr(model, 'layers', {}).items() if isinstance(l, E.BatchNorm)]
        if bn:
            # batch statistics would be per-rank (not the statistics of the global batch) and
            # the moving averages would drift apart after the initial broadcast
            raise Exception('DataParallel: batch-coupled layers (%s) are not supported; images '
                            'must be independent in forward and backward (SURVEY 8e)' % ', '.join(bn))
        self.model = model
        if group is None and dist.get_backend() == 'nccl' and \
                os.environ.get('SEGB200_NCCL_PRIO', '1') != '0':
            # The all-reduce kernels must find SM slots beside persistent 148-CTA kernels that
            # hold every SM until they exit: on a high-priority stream NCCL's CTAs are placed
            # first at the next kernel boundary instead of queueing behind every pending CTA of
            # the backward pass (measured at 2 GPUs: bucket conv3-4 ready at ~900 us, started at
            # 961 us on a default-priority stream; profiles/r02_dp.md)
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            # Optional (SEGB200_NCCL_CTAS=n, default off): hold NCCL to n CTAs and size the tile
            # kernels' persistent grids for the remaining SMs, so that an SM held by an
            # all-reduce does not force a second wave on every overlapped launch.  Measured at
            # 2 GPUs: n = 4 / 8 / 16 -> 1.29 / 1.15 / 1.08 ms per step against 1.03 with NCCL's
            # own choice (32 CTAs, ring LL): few channels cannot carry 31 MB inside the backward
            # pass, the exposed reduction costs more than the second waves it avoids.
            ctas = int(os.environ.get('SEGB200_NCCL_CTAS', '0'))
            if ctas > 0:
                try:
                    opts.config.min_ctas = ctas
                    opts.config.max_ctas = ctas
                    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
                    N.set_option(N.OPT_SM_LIMIT, sms - ctas)
                except AttributeError:                   # older torch: no ncclConfig fields
                    pass
            group = dist.new_group(backend='nccl', pg_options=opts)
        self.group = group
        self.world = dist.get_world_size(group)
        # executors built (and graphs captured) before wrapping have no all-reduce and a
        # grad_scale of 1 baked in
        model._exec.clear()
        model._last_train_exec = None
        model.world_size = self.world
        st = model.store
        dist.broadcast(st.master, src=0, group=group)
        for t in st.state.values():
            dist.broadcast(t, src=0, group=group)
        st.refresh_shadow()
        bounds = [g['slice'][0] for g in model.opt_groups] + [st.numel]
        self.buckets = GradBuckets(st.grad, bounds, group)
        model._allreduce = self.allreduce_group

    def allreduce_group(self, i):
        """All-reduce (sum) bucket i; the current stream waits for it (no host wait)."""
        w = dist.all_reduce(self.buckets.slices[i], op=dist.ReduceOp.SUM, group=self.group,
                            async_op=True)
        w.wait()
