"""Host-side execution engine: parameter store + layer-level wrappers over the
C ABI (include/segb200.h).  The three model classes build explicit forward and
backward schedules out of these wrappers (no torch.autograd on the hot path);
a whole train step is a fixed kernel sequence that is captured into a CUDA graph.

What the pieces replace in the reference:
  * ParamStore  <- the slim variables + `tf.train.AdamOptimizer` slots
                   (/root/reference/models/basemodel.py:321,366)
  * ConvLayer   <- slim.convolution2d / slim.convolution2d_transpose call sites
"""
import ctypes
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import native as N

BF16 = torch.bfloat16


def pad16(c):
    return (c + 15) // 16 * 16


def same_pad(n, k, s):
    """TF SAME padding (extra goes bottom/right)."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def xavier_uniform(shape, gen):
    """tf.contrib.layers.xavier_initializer() (uniform), slim's default
    weights_initializer for convolution2d / convolution2d_transpose."""
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in, fan_out = rf * shape[-2], rf * shape[-1]
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (gen.random(shape, dtype=np.float32) * 2 - 1) * np.float32(limit)


class Param(object):
    __slots__ = ('name', 'shape', 'offset', 'numel', 'shadow_offset', 'shadow_shape',
                 'shadow_src', 'trainable', 'store')

    def value(self):
        return self.store.master[self.offset:self.offset + self.numel].view(self.shape)

    def grad(self):
        return self.store.grad[self.offset:self.offset + self.numel].view(self.shape)

    def shadow(self):
        n = int(np.prod(self.shadow_shape))
        return self.store.shadow[self.shadow_offset:self.shadow_offset + n].view(self.shadow_shape)


class ParamStore(object):
    """Flat fp32 master / grad / Adam-slot buffers in TF variable order, plus a
    flat bf16 shadow with channel-padded weight copies.  Non-trainable state
    (BN moving statistics) lives in a separate flat buffer."""

    def __init__(self, device):
        self.device = device
        self.params = OrderedDict()
        self.state = OrderedDict()      # name -> fp32 tensor (non-trainable)
        self._n = 0
        self._ns = 0
        self.finalized = False

    def add(self, name, shape, shadow_shape=None, shadow_src=None):
        """shadow_src: the shape the master is read as when it is copied into the padded
        bf16 shadow (default: `shape`; e.g. an HWIO tensor read as [1,1,H*W*I,O])."""
        assert not self.finalized
        p = Param()
        p.name, p.shape, p.store = name, tuple(shape), self
        p.shadow_src = tuple(shadow_src) if shadow_src is not None else tuple(shape)
        p.offset, p.numel = self._n, int(np.prod(shape))
        p.trainable = True
        self._n += p.numel
        if shadow_shape is not None:
            p.shadow_shape = tuple(shadow_shape)
            p.shadow_offset = self._ns
            self._ns += int(np.prod(shadow_shape))
            self._ns = (self._ns + 127) // 128 * 128      # keep 256-byte alignment
        else:
            p.shadow_shape, p.shadow_offset = None, -1
        self.params[name] = p
        return p

    def add_state(self, name, init):
        self.state[name] = init.to(self.device, torch.float32).clone()
        return self.state[name]

    def finalize(self):
        dev = self.device
        n = max(self._n, 1)
        self.master = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(max(self._ns, 1), dtype=BF16, device=dev)
        seg, soff = [], []
        for p in self.params.values():
            if p.shadow_shape is not None:
                inner, mid = p.shadow_src[-1], p.shadow_src[-2]
                inner_pad, mid_pad = p.shadow_shape[-1], p.shadow_shape[-2]
            else:
                inner = inner_pad = p.shape[-1]
                mid = mid_pad = 1
            seg += [p.offset, p.numel, inner, inner_pad, mid, mid_pad]
            soff.append(p.shadow_offset)
        # launch plan of the Adam kernel: one {segment, first element} pair per block
        ce = N.load().seg_adam_chunk_elems()
        chunks = []
        self.chunk_first = []           # first Adam chunk of each parameter (+ sentinel)
        for si, p in enumerate(self.params.values()):
            self.chunk_first.append(len(chunks) // 2)
            for b in range(0, p.numel, ce):
                chunks += [si, b]
        self.chunk_first.append(len(chunks) // 2)
        self.nchunks = len(chunks) // 2
        self.chunks = torch.tensor(chunks or [0, 0], dtype=torch.int32, device=dev)
        self.segments = torch.tensor(seg, dtype=torch.int32, device=dev)
        self.shadow_offsets = torch.tensor(soff, dtype=torch.int64, device=dev)
        self.numel = self._n
        self.step = 0
        self.lr_t_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self.finalized = True

    # -- host <-> device parameter exchange (TF names / TF layouts) ---------
    def load_state_dict(self, sd):
        for name, p in self.params.items():
            if name in sd:
                t = torch.as_tensor(np.asarray(sd[name]), dtype=torch.float32).reshape(p.shape)
                p.value().copy_(t.to(self.device))
        for name in self.state:
            if name in sd:
                self.state[name].copy_(torch.as_tensor(np.asarray(sd[name]),
                                                       dtype=torch.float32).to(self.device))
        self.refresh_shadow()

    def state_dict(self):
        out = OrderedDict()
        for name, p in self.params.items():
            out[name] = p.value().detach().cpu().numpy().copy()
        for name, t in self.state.items():
            out[name] = t.detach().cpu().numpy().copy()
        return out

    def refresh_shadow(self):
        """Re-derive every bf16 shadow from its fp32 master (init / load)."""
        for p in self.params.values():
            if p.shadow_shape is None:
                continue
            sh = p.shadow()
            sh.zero_()
            idx = tuple(slice(0, s) for s in p.shadow_src)
            sh[idx] = p.value().reshape(p.shadow_src).to(BF16)

    def adam_step(self, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
        """tf.train.AdamOptimizer update over the flat buffer (one launch); also
        rewrites the bf16 shadows and zeroes the gradient buffer."""
        self.adam_launch(self.next_lr_t(lr, beta1, beta2), beta1, beta2, eps, grad_scale)

    def next_lr_t(self, lr, beta1=0.9, beta2=0.999):
        """Advance the step counter; returns lr*sqrt(1-b2^t)/(1-b1^t)."""
        self.step += 1
        return lr * math.sqrt(1.0 - beta2 ** self.step) / (1.0 - beta1 ** self.step)

    def chunk_range(self, off_a, off_b):
        """Adam chunk index range of the parameters whose flat offsets lie in
        [off_a, off_b) (both must be parameter boundaries)."""
        offs = [p.offset for p in self.params.values()] + [self.numel]
        return self.chunk_first[offs.index(off_a)], self.chunk_first[offs.index(off_b)]

    def adam_launch(self, lr_t, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0,
                    from_device=False, chunk_range=None):
        """from_device=True reads lr_t from self.lr_t_dev (CUDA-graph replay).
        chunk_range=(a, b): update only the parameters covered by chunks [a, b)."""
        a, b = chunk_range if chunk_range is not None else (0, self.nchunks)
        if b <= a:
            return
        chunks = ctypes.c_void_p(self.chunks.data_ptr() + a * 8)
        if N.TIMELINE is not None:
            # 4 B grad + 3 x 4 B state read + 3 x 4 B written + 2 B shadow per parameter
            ce = N.load().seg_adam_chunk_elems()
            N.note_work(0, min((b - a) * ce, self.numel) * 30.0)
        N.call('seg_adam_multi', N.ptr(self.master), N.ptr(self.grad), N.ptr(self.m),
               N.ptr(self.v), N.ptr(self.shadow), N.ptr(self.segments),
               N.ptr(self.shadow_offsets), chunks, b - a, lr_t,
               N.ptr(self.lr_t_dev) if from_device else None, beta1, beta2,
               eps, grad_scale, N.stream_ptr())


class ConvLayer(object):
    """One slim.convolution2d (kind='conv', weights HWIO) or
    slim.convolution2d_transpose (kind='deconv', weights HWOI) call site."""

    def __init__(self, store, name, kind, k, stride, padding, cin, cout, relu=True, gen=None,
                 in_pad=None):
        self.name, self.kind, self.k, self.stride, self.padding = name, kind, k, stride, padding
        self.cin, self.cout, self.relu = cin, cout, relu
        self.cin_pad = in_pad if in_pad is not None else pad16(cin)
        self.cout_pad = pad16(cout)
        if kind == 'conv':
            shape, sshape = (k, k, cin, cout), (k, k, self.cin_pad, self.cout_pad)
        else:
            shape, sshape = (k, k, cout, cin), (k, k, self.cout_pad, self.cin_pad)
        self.w = store.add(name + '/weights', shape, sshape)
        self.b = store.add(name + '/biases', (cout,))
        self._init = xavier_uniform(shape, gen) if gen is not None else None

    def init_values(self):
        if self._init is not None:
            self.w.value().copy_(torch.from_numpy(self._init).to(self.w.store.device))
            self._init = None

    def out_hw(self, h, w):
        k, s = self.k, self.stride
        if self.kind == 'conv':
            if self.padding == 'SAME':
                return -(-h // s), -(-w // s)
            return (h - k) // s + 1, (w - k) // s + 1
        if self.padding == 'SAME':
            return h * s, w * s
        return h * s + max(k - s, 0), w * s + max(k - s, 0)

    def desc(self, in_h, in_w, flags, impl):
        """seg_conv_desc.  For 'conv', (in_h,in_w) is the conv input size; for
        'deconv' it is the size of the LARGE side (the deconv output), whose
        SAME crop plays the role of the padding."""
        k, s = self.k, self.stride
        if self.padding == 'SAME':
            if self.kind == 'conv':
                pt, pb = same_pad(in_h, k, s)
                pl, pr = same_pad(in_w, k, s)
            else:
                tot = max(k - s, 0)
                pt = pl = tot // 2
                pb = pr = tot - tot // 2
        else:
            pt = pb = pl = pr = 0
        return N.SegConvDesc(k, k, s, pt, pl, pb, pr, self.cin, self.cout, self.cin_pad,
                             self.cout_pad, flags, impl)

    def work(self, small, big, big_bytes=2):
        """(flops, bytes) of one fwd / dgrad / wgrad launch of this layer between the
        tensor on its input side (`small`: x or dx) and on its output side (`big`: y or dz):
        2*N*Ho*Wo*Cout*Cin*kh*kw (conv) / 2*N*Hi*Wi*Cin*Cout*kh*kw (transposed conv); both
        tensors once, at their real channel counts."""
        n = small.shape[0]
        px = big.shape[1] * big.shape[2] if self.kind == 'conv' else small.shape[1] * small.shape[2]
        flops = 2.0 * n * px * self.cout * self.cin * self.k * self.k
        in_b = n * small.shape[1] * small.shape[2] * self.cin * 2.0
        return flops, in_b + n * big.shape[1] * big.shape[2] * self.cout * float(big_bytes)

    def epi_flags(self, out_f32=False):
        f = N.EPI_BIAS
        if self.relu:
            f |= N.EPI_RELU
        if out_f32:
            f |= N.EPI_OUT_F32
        return f

    # ---- forward -----------------------------------------------------------
    def forward(self, x, y, x2=None, impl=N.IMPL_UMMA, out_f32=False):
        st = N.stream_ptr()
        N.set_tag(self.name)
        N.note_work(*self.work(x, y, 4 if out_f32 else 2))
        if y.shape[3] > self.cout:
            # the bias vector has `cout` entries and is indexed by output column: hand the
            # kernels the logical channels only (the padded ones stay at their zero fill)
            y = y[..., :self.cout]
        if self.kind == 'conv':
            d = self.desc(x.shape[1], x.shape[2], self.epi_flags(out_f32), impl)
            N.call('seg_conv2d_fwd', ctypes.byref(d), N.vref(x), N.vref(x2),
                   N.ptr(self.w.shadow()), N.ptr(self.b.value()), N.vref(y), st)
        else:
            d = self.desc(y.shape[1], y.shape[2], self.epi_flags(out_f32), impl)
            N.call('seg_deconv2d_fwd', ctypes.byref(d), N.vref(x), N.ptr(self.w.shadow()),
                   N.ptr(self.b.value()), N.vref(y), st)

    def forward_bn_infer(self, x, y, bn, impl=N.IMPL_UMMA):
        """Inference: this layer and the batch-norm `bn` that follows it (moving statistics)
        as one launch (seg_conv2d_fwd_affine / seg_deconv2d_fwd_affine) - the un-normalised
        activation is never stored.  Returns False, with nothing launched that matters, where
        the entry does not take the geometry: the caller then runs the unfused pair."""
        if impl != N.IMPL_UMMA or os.environ.get('SEGB200_FUSE_BN', '1') == '0':
            return False
        st = N.stream_ptr()
        scale, shift = bn.fold(self.cout_pad)
        if y.shape[3] > self.cout:
            y = y[..., :self.cout]
        N.set_tag(self.name)
        N.note_work(*self.work(x, y, 2))
        hw = (x.shape[1], x.shape[2]) if self.kind == 'conv' else (y.shape[1], y.shape[2])
        d = self.desc(hw[0], hw[1], self.epi_flags(False), impl)
        name = 'seg_conv2d_fwd_affine' if self.kind == 'conv' else 'seg_deconv2d_fwd_affine'
        try:
            N.call(name, ctypes.byref(d), N.vref(x), N.ptr(self.w.shadow()),
                   N.ptr(self.b.value()), N.ptr(scale), N.ptr(shift), N.vref(y), st)
        except N.SegError as e:
            if e.status != N.E_UNSUPPORTED:
                raise
            return False
        return True

    # ---- first layer fused with its 2x2 max-pool (seg_conv2d_pool_fwd / _wgrad) ----
    def pool_fusable(self, x, pooled, impl=N.IMPL_UMMA):
        """True if conv + 2x2/2 max-pool of this layer can run as one launch: the first-layer
        kernel's shape (x is the (R,G,B,1) input), 32 padded output channels, even output
        grid.  SEGB200_FUSE_POOL1=0 disables it."""
        if impl != N.IMPL_UMMA or self.kind != 'conv' or type(self) is not ConvLayer:
            return False
        if self.k != 3 or self.stride != 1 or self.cin != 3 or x.shape[3] != 4 or self.cout_pad != 32:
            return False
        oh, ow = self.out_hw(x.shape[1], x.shape[2])
        return (oh % 2 == 0 and ow % 2 == 0 and tuple(pooled.shape[1:]) == (oh // 2, ow // 2, 32)
                and os.environ.get('SEGB200_FUSE_POOL1', '1') != '0')

    def forward_pool(self, x, pooled, argmax, y_win=None, win_y0=0, win_x0=0):
        st = N.stream_ptr()
        N.set_tag(self.name)
        oh, ow = self.out_hw(x.shape[1], x.shape[2])
        px = x.shape[0] * oh * ow
        # algorithmic work: the convolution; bytes = input + pooled output + slots (+ window)
        N.note_work(2.0 * px * self.cout * self.cin * 9,
                    x.numel() * 2.0 + pooled.numel() * 3.0 + (y_win.numel() * 2.0 if y_win is not None else 0))
        d = self.desc(x.shape[1], x.shape[2], self.epi_flags(False), N.IMPL_UMMA)
        N.call('seg_conv2d_pool_fwd', ctypes.byref(d), N.vref(x), N.ptr(self.w.shadow()),
               N.ptr(self.b.value()), N.vref(y_win), int(win_y0), int(win_x0), N.vref(pooled),
               N.ptr(argmax), st)

    def wgrad_pool(self, x, dpool, argmax, pooled):
        st = N.stream_ptr()
        N.set_tag(self.name)
        oh, ow = self.out_hw(x.shape[1], x.shape[2])
        px = x.shape[0] * oh * ow
        N.note_work(2.0 * px * self.cout * self.cin * 9, x.numel() * 2.0 + pooled.numel() * 5.0)
        d = self.desc(x.shape[1], x.shape[2], 0, N.IMPL_UMMA)
        N.call('seg_conv2d_pool_wgrad', ctypes.byref(d), N.vref(x), N.vref(dpool), N.ptr(argmax),
               N.vref(pooled), N.ptr(self.w.grad()), N.ptr(self.b.grad()), st)

    # ---- backward ----------------------------------------------------------
    def backward(self, x, dz, dx=None, x2=None, dx2=None, mask=None, mask2=None,
                 impl=N.IMPL_UMMA, y_hw=None, dz_bias=None, side=None, after=None):
        """Accumulates dW, db into the grad buffer; writes dx (/dx2) if given.
        `dz` is the gradient w.r.t. this layer's pre-activation output (channels
        padded to cout_pad).  `mask`/`mask2` = forward tensors whose ReluGrad is
        applied to dx/dx2.  `dz_bias`: view of dz restricted to the real cout
        channels when cout_pad != cout.
        `side`: a SideStream — the weight/bias gradient kernels are enqueued there (they
        only feed the optimizer, so they run beside the input-gradient chain); `after()` is
        called once they are enqueued (in the side stream's context)."""
        N.set_tag(self.name)
        if side is not None:
            with side.fork():
                self._wgrad(x, dz, x2, impl, dz_bias)
                if after is not None:
                    after()
        else:
            self._wgrad(x, dz, x2, impl, dz_bias)
            if after is not None:
                after()
        if dx is None:
            return
        st = N.stream_ptr()
        N.note_work(*self.work(dx, dz))
        if self.kind == 'conv':
            d = self.desc(x.shape[1], x.shape[2], 0, impl)
            N.call('seg_conv2d_dgrad', ctypes.byref(d), N.vref(dz), N.ptr(self.w.shadow()),
                   N.vref(dx), N.vref(dx2), N.vref(mask), N.vref(mask2), st)
        else:
            d = self.desc(dz.shape[1], dz.shape[2], 0, impl)
            N.call('seg_deconv2d_dgrad', ctypes.byref(d), N.vref(dz),
                   N.ptr(self.w.shadow()), N.vref(dx), N.vref(mask), st)

    def dgrad_slice(self, dz, cin_lo, dx, mask=None, impl=N.IMPL_UMMA):
        """Input gradient of the channel slice [cin_lo, cin_lo + dx.c) only (tcgen05 path):
        the two halves of a virtual concat can be computed by separate launches, the skip
        half off the critical path (seg_conv2d_dgrad_slice)."""
        assert self.kind == 'conv'
        N.set_tag(self.name)
        d = self.desc(dx.shape[1], dx.shape[2], 0, impl)
        N.call('seg_conv2d_dgrad_slice', ctypes.byref(d), N.vref(dz), N.ptr(self.w.shadow()),
               int(cin_lo), N.vref(dx), N.vref(mask), N.stream_ptr())

    def _wgrad(self, x, dz, x2, impl, dz_bias):
        st = N.stream_ptr()
        if self.kind == 'conv':
            # BiasAddGrad is fused into the wgrad GEMM
            N.note_work(*self.work(x, dz))
            d = self.desc(x.shape[1], x.shape[2], 0, impl)
            N.call('seg_conv2d_wgrad', ctypes.byref(d), N.vref(x), N.vref(x2), N.vref(dz),
                   N.ptr(self.w.grad()), N.ptr(self.b.grad()), st)
        else:
            dzb = dz_bias if dz_bias is not None else (dz if dz.shape[3] == self.cout
                                                       else dz[..., :self.cout])
            N.note_work(0, dzb.shape[0] * dzb.shape[1] * dzb.shape[2] * self.cout * 2.0)
            N.call('seg_bias_grad', N.vref(dzb), N.ptr(self.b.grad()), st)
            N.note_work(*self.work(x, dz))
            d = self.desc(dz.shape[1], dz.shape[2], 0, impl)
            N.call('seg_deconv2d_wgrad', ctypes.byref(d), N.vref(x), N.vref(dz),
                   N.ptr(self.w.grad()), st)


class PatchConvLayer(ConvLayer):
    """The first convolution of a model, on the raw few-channel input: seg_pack_patches
    puts the k x k x cin patch of every output pixel into the channel axis, after which the
    layer IS a 1x1 convolution over pad16(k*k*cin) channels.  Same TF variables
    (`<name>/weights` [k,k,cin,cout] HWIO, `<name>/biases`): the HWIO tensor read as
    [k*k*cin][cout] is the 1x1 weight matrix, and the 1x1 weight gradient lands in the
    HWIO gradient unchanged.  (3 channels padded to 16 per tap waste 13/16 of every MMA and
    make 32-byte TMA rows; the packed layer has 9x fewer MMAs and one 64-byte row per pixel.)"""

    def __init__(self, store, name, k, stride, padding, cin, cout, relu=True, gen=None):
        self.name, self.kind, self.relu = name, 'conv', relu
        self.patch_k, self.patch_stride, self.patch_padding, self.patch_cin = k, stride, padding, cin
        self.k, self.stride, self.padding = 1, 1, 'VALID'
        self.cin, self.cout = k * k * cin, cout
        self.cin_pad, self.cout_pad = pad16(self.cin), pad16(cout)
        shape = (k, k, cin, cout)
        self.w = store.add(name + '/weights', shape, (1, 1, self.cin_pad, self.cout_pad),
                           shadow_src=(1, 1, self.cin, cout))
        self.b = store.add(name + '/biases', (cout,))
        self._init = xavier_uniform(shape, gen) if gen is not None else None

    def patch_out_hw(self, h, w):
        k, s = self.patch_k, self.patch_stride
        if self.patch_padding == 'SAME':
            return -(-h // s), -(-w // s)
        return (h - k) // s + 1, (w - k) // s + 1

    def pack(self, x_f32, y):
        """x_f32 [B,H,W,cin] fp32 -> y [B,Ho,Wo,cin_pad] bf16 patch tensor."""
        k, s = self.patch_k, self.patch_stride
        h, w = x_f32.shape[1], x_f32.shape[2]
        pt = same_pad(h, k, s)[0] if self.patch_padding == 'SAME' else 0
        pl = same_pad(w, k, s)[0] if self.patch_padding == 'SAME' else 0
        N.call('seg_pack_patches', N.ptr(x_f32), x_f32.shape[3], h, w, k, k, s, pt, pl,
               N.vref(y), N.stream_ptr())


class SideStream(object):
    """A second stream for work that is off the critical path of a schedule (the weight
    gradients of the backward pass).  fork(): the side stream first waits for everything
    enqueued so far on the current stream; join(): the current stream waits for the side
    stream.  Works inside CUDA-graph capture (fork/join become graph edges)."""

    def __init__(self, device, lanes=1, priority=0):
        """`lanes` > 1: successive fork()s rotate over that many streams, so that narrow
        kernels of different layers (weight-gradient grids of 30-128 CTAs) can overlap each
        other as well as the current stream.  `priority` < 0: a high-priority stream (its
        kernels' CTAs are placed first when SM slots free up; captured into graph nodes)."""
        self.streams = [torch.cuda.Stream(device=device, priority=priority)
                        for _ in range(max(1, lanes))]
        self.stream = self.streams[0]
        self._used = [False] * len(self.streams)
        self._next = 0

    @property
    def used(self):
        return any(self._used)

    def wait_into(self, stream):
        """`stream` waits for everything enqueued on this side stream's lanes."""
        for s, u in zip(self.streams, self._used):
            if u:
                ev = torch.cuda.Event()
                ev.record(s)
                stream.wait_event(ev)

    def fork(self, also=None):
        """`also`: another SideStream whose enqueued work must be waited for as well."""
        i = self._next
        self._next = (i + 1) % len(self.streams)
        self.stream = self.streams[i]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.stream.wait_event(ev)
        if also is not None:
            for o in (also if isinstance(also, (tuple, list)) else (also,)):
                o.wait_into(self.stream)
        self._used[i] = True
        return torch.cuda.stream(self.stream)

    def join(self):
        self.wait_into(torch.cuda.current_stream())
        self._used = [False] * len(self.streams)


# ---------------------------------------------------------------------------
# thin op wrappers
# ---------------------------------------------------------------------------
def pack_input(x_f32, y):
    N.note_work(0, x_f32.numel() * 4.0 + y.numel() * 2.0)
    N.call('seg_pack_input', N.ptr(x_f32), x_f32.shape[3], N.vref(y), N.stream_ptr())


def maxpool_fwd(x, y, argmax, k=2, s=2):
    N.note_work(0, x.numel() * 2.0 + y.numel() * 2.0 + argmax.numel())
    N.call('seg_maxpool_fwd', N.vref(x), k, s, N.vref(y), N.ptr(argmax), N.stream_ptr())


def maxpool_bwd(dy, argmax, dx, k=2, s=2, add=None, add_y0=0, add_x0=0, mask=None, pooled=None):
    """`pooled`: the forward pool output; lets the kernel skip reading `mask` (the pool
    input) wherever no `add` gradient arrives."""
    # dy + argmax slots (+ pool output for the mask) in, dx out, + the skip window's add / mask
    N.note_work(0, dy.numel() * 2.0 + argmax.numel() + dx.numel() * 2.0 +
                (pooled.numel() * 2.0 if pooled is not None else
                 (mask.numel() * 2.0 if mask is not None else 0.0)) +
                (add.numel() * 2.0 * (2 if (mask is not None and pooled is not None) else 1)
                 if add is not None else 0.0))
    if pooled is not None:
        N.call('seg_maxpool_bwd_y', N.vref(dy), N.ptr(argmax), k, s, N.vref(add), add_y0, add_x0,
               N.vref(mask), N.vref(pooled), N.vref(dx), N.stream_ptr())
        return
    N.call('seg_maxpool_bwd', N.vref(dy), N.ptr(argmax), k, s, N.vref(add), add_y0, add_x0,
           N.vref(mask), N.vref(dx), N.stream_ptr())


def stage_input(x, y4, mask_src=None, mask_dst=None, crop_yx=None, mask_kind=None, ctl=None):
    """seg_stage_input: x fp32 [B,Hs,Ws,C] in [0,1] or uint8 (divided by 255 on the device,
    reference utils/datasets.py:176-178) -> y4 bf16 [B,H,W,4] = (R,G,B,1), the first-layer
    kernel's input; optional per-image crop (device int32 [B,2], :184-185), the mask that
    goes with it, and the per-step scalars of `ctl` (a native.SegStageCtl)."""
    kind = 1 if x.dtype == torch.uint8 else 0
    assert x.is_contiguous() and (kind == 1 or x.dtype == torch.float32)
    if mask_kind is None:
        mask_kind = kind                     # raw 0/255 masks travel with raw uint8 images
    if mask_src is not None:
        assert mask_src.is_contiguous() and mask_src.dtype == torch.uint8
        assert mask_dst.is_contiguous() and mask_dst.dtype == torch.uint8
        assert mask_src.shape[1] == x.shape[1] and mask_src.shape[2] == x.shape[2]
    N.note_work(0, y4.shape[0] * y4.shape[1] * y4.shape[2] * (x.shape[3] * x.element_size() + 8.0))
    N.call('seg_stage_input', N.ptr(x), kind, x.shape[3], x.shape[1], x.shape[2], N.ptr(crop_yx),
           N.vref(y4), N.ptr(mask_src), int(mask_kind), N.ptr(mask_dst),
           ctypes.byref(ctl) if ctl is not None else None, N.stream_ptr())


def conv_bn_pool_infer(layer, bn, x4, pooled, argmax=None):
    """seg_conv2d_bn_pool_infer: first convolution (`layer`: a ConvLayer or PatchConvLayer of an
    RGB input, 3x3/s1 or 5x5/s2) + ReLU + batch-norm `bn` (moving statistics) + 2x2/2 max-pool in
    one launch on the (R,G,B,1) staged input (reference models/deconvolution.py:109-118)."""
    patch = isinstance(layer, PatchConvLayer)
    k = layer.patch_k if patch else layer.k
    s = layer.patch_stride if patch else layer.stride
    padding = layer.patch_padding if patch else layer.padding
    h, w = x4.shape[1], x4.shape[2]
    if padding == 'SAME':
        (pt, pb), (pl, pr) = same_pad(h, k, s), same_pad(w, k, s)
    else:
        pt = pb = pl = pr = 0
    cout = layer.cout
    d = N.SegConvDesc(k, k, s, pt, pl, pb, pr, 3, cout, 16, layer.cout_pad,
                      N.EPI_BIAS | (N.EPI_RELU if layer.relu else 0), N.IMPL_UMMA)
    N.set_tag(layer.name)
    px = pooled.shape[0] * pooled.shape[1] * pooled.shape[2] * 4
    N.note_work(2.0 * px * cout * 3 * k * k, x4.numel() * 2.0 + pooled.numel() * 2.0)
    N.call('seg_conv2d_bn_pool_infer', ctypes.byref(d), N.vref(x4), N.ptr(layer.w.shadow()),
           3 if patch else 0, N.ptr(layer.b.value()), N.ptr(bn.moving_mean), N.ptr(bn.moving_var),
           bn.eps, N.ptr(bn.beta.value()), N.vref(pooled), N.ptr(argmax), N.stream_ptr())


def classmap_tail_infer(x, rh, rw, up, bn, conv_out, logits, probs, labelmap):
    """seg_classmap_tail_infer: resize_bilinear -> 2x2/s2 transposed conv `up` (+ReLU) ->
    batch-norm `bn` (moving statistics) -> 3x3 SAME conv `conv_out` -> sigmoid / argmax, one
    launch (reference models/deconvolution.py:163-174, :79-82)."""
    n, h, w = x.shape[0], 2 * rh, 2 * rw
    N.set_tag('tail')
    N.note_work(2.0 * n * h * w * (up.cin * up.cout + 9 * conv_out.cin * conv_out.cout),
                x.numel() * 2.0 + n * h * w * (conv_out.cout * 8.0 + 4.0))
    N.call('seg_classmap_tail_infer', N.vref(x), rh, rw, N.ptr(up.w.shadow()), up.cout_pad,
           up.cin_pad, N.ptr(up.b.value()), N.ptr(bn.moving_mean), N.ptr(bn.moving_var), bn.eps,
           N.ptr(bn.beta.value()), N.ptr(conv_out.w.shadow()), conv_out.cin_pad, conv_out.cout_pad,
           N.ptr(conv_out.b.value()), conv_out.cout, N.ptr(logits), N.ptr(probs), N.ptr(labelmap),
           N.stream_ptr())


def softmax_xent(logits, labels, loss_sum, dlogits=None):
    N.call('seg_softmax_xent_fwd_bwd', N.vref(logits), N.vref(labels), N.ptr(loss_sum),
           N.vref(dlogits), N.stream_ptr())


def upscore8_xent(x, labels, loss_sum, dx, mask=None, logits=None):
    """seg_upscore8_xent_fwd_bwd: FCN-8s training head in one launch - bilinear x8 upscore of
    the class-score map `x`, softmax cross-entropy against `labels`, and the upscore's input
    gradient `dx`; neither the full-resolution logits (unless `logits` is given) nor their
    gradient touch HBM (reference models/fcn.py:207-220, models/basemodel.py:59-70)."""
    n, h, w, c = x.shape
    px = n * h * w * 64
    N.set_tag('upscore_xent')
    N.note_work(px * c * 16.0, x.numel() * 4.0 + px * 1.0 + (px * c * 4.0 if logits is not None else 0.0))
    N.call('seg_upscore8_xent_fwd_bwd', N.vref(x), N.vref(labels), N.ptr(loss_sum), N.vref(dx),
           N.vref(mask), N.ptr(logits), N.stream_ptr())


def head1x1_xent(x, layer, labels, logits, loss_sum, dx):
    """Fused training head: 1x1 conv `layer` (<= 4 classes) + softmax x-entropy + the
    layer's weight / bias gradient + the input gradient dx, one pass over x."""
    N.set_tag(layer.name)
    px = x.shape[0] * x.shape[1] * x.shape[2]
    N.note_work(3 * 2.0 * px * layer.cin * layer.cout,
                px * (layer.cin * 2.0 * 2 + layer.cout * 4.0 + 1))
    N.call('seg_head1x1_xent', N.vref(x), N.ptr(layer.w.shadow()), layer.cout_pad,
           N.ptr(layer.b.value()), N.vref(labels), layer.cout, N.vref(logits), N.ptr(loss_sum),
           N.vref(dx), N.ptr(layer.w.grad()), N.ptr(layer.b.grad()), N.stream_ptr())


def sigmoid_argmax(logits, probs, labelmap):
    N.call('seg_sigmoid_argmax', N.vref(logits), N.ptr(probs), N.ptr(labelmap), N.stream_ptr())


def fill_zero(t):
    N.call('seg_fill_zero', N.ptr(t), t.numel() * t.element_size(), N.stream_ptr())


def bilinear_upsample_fwd(x, factor, y, add=None):
    N.call('seg_bilinear_upsample_fwd', N.vref(x), factor, N.vref(add), N.vref(y),
           1 if y.dtype == torch.float32 else 0, N.stream_ptr())


def bilinear_upsample_bwd(dy, factor, dx, mask=None):
    N.call('seg_bilinear_upsample_bwd', N.vref(dy), 1 if dy.dtype == torch.float32 else 0,
           factor, N.vref(mask), N.vref(dx), N.stream_ptr())


def maxpool_bwd2(dy, dy2, argmax, dx, k=2, s=2, mask=None, pooled=None):
    """`pooled`: the forward pool output; the ReLU mask is then taken from it and the pool
    input `mask` is not read."""
    N.note_work(0, dy.numel() * 4.0 + argmax.numel() + dx.numel() * 2.0 +
                (pooled.numel() * 2.0 if pooled is not None else
                 (mask.numel() * 2.0 if mask is not None else 0.0)))
    if pooled is not None:
        N.call('seg_maxpool_bwd2_y', N.vref(dy), N.vref(dy2), N.ptr(argmax), k, s, N.vref(mask),
               N.vref(pooled), N.vref(dx), N.stream_ptr())
    else:
        N.call('seg_maxpool_bwd2', N.vref(dy), N.vref(dy2), N.ptr(argmax), k, s, N.vref(mask),
               N.vref(dx), N.stream_ptr())


def relu_grad(dy, y, dz):
    N.call('seg_relu_grad', N.vref(dy), N.vref(y), N.vref(dz), N.stream_ptr())


def resize_bilinear_fwd(x, y):
    N.call('seg_resize_bilinear_fwd', N.vref(x), N.vref(y), N.stream_ptr())


def resize_bilinear_bwd(dy, dx):
    N.call('seg_resize_bilinear_bwd', N.vref(dy), N.vref(dx), N.stream_ptr())


def dropout(x, y, seed, stream_id, keep_prob=0.5):
    N.call('seg_dropout', N.vref(x), seed, stream_id, keep_prob, N.vref(y), N.stream_ptr())


def dropout_ex(x, y, seed, stream_id0, per_image_step=0, step_dev=None, step_mul=0,
               keep_prob=0.5):
    """One launch for the whole batch: image n uses Philox stream
    stream_id0 + n*per_image_step + step_dev[0]*step_mul (step_dev: device int32/uint32
    scalar read at run time, so a captured CUDA graph gets fresh masks on every replay)."""
    N.call('seg_dropout_ex', N.vref(x), seed, stream_id0, per_image_step, N.ptr(step_dev),
           step_mul, keep_prob, N.vref(y), N.stream_ptr())


def mc_mean_var(probs, mean, var):
    t = probs.shape[0]
    N.call('seg_mc_mean_var', N.ptr(probs), t, probs[0].numel(), N.ptr(mean), N.ptr(var),
           N.stream_ptr())


class BatchNorm(object):
    """slim.batch_norm with slim defaults (decay .999, eps 1e-3, center only),
    applied AFTER the ReLU of the producing layer
    (/root/reference/models/deconvolution.py:116)."""

    def __init__(self, store, name, c, decay=0.999, eps=1e-3):
        self.name, self.c, self.decay, self.eps = name, c, decay, eps
        self.beta = store.add(name + '/beta', (c,))
        self.moving_mean = store.add_state(name + '/moving_mean', torch.zeros(c))
        self.moving_var = store.add_state(name + '/moving_variance', torch.ones(c))
        dev = store.device
        self.scratch = torch.zeros(6, c, dtype=torch.float32, device=dev)

    def forward(self, x, y, training=True):
        st = N.stream_ptr()
        xs = x[..., :self.c]
        if training:
            s = self.scratch
            fill_zero(s[0:2])
            N.call('seg_batchnorm_stats', N.vref(xs), N.ptr(s[0]), N.ptr(s[1]), st)
            count = x.shape[0] * x.shape[1] * x.shape[2]
            N.call('seg_batchnorm_finalize', N.ptr(s[0]), N.ptr(s[1]), count, self.c, self.eps,
                   self.decay, N.ptr(s[2]), N.ptr(s[3]), N.ptr(self.moving_mean),
                   N.ptr(self.moving_var), st)
            N.call('seg_batchnorm_apply', N.vref(xs), N.ptr(s[2]), N.ptr(s[3]),
                   N.ptr(self.beta.value()), N.vref(y[..., :self.c]), st)
        else:
            N.call('seg_batchnorm_infer', N.vref(xs), N.ptr(self.moving_mean),
                   N.ptr(self.moving_var), self.eps, N.ptr(self.beta.value()),
                   N.vref(y[..., :self.c]), st)

    def fold(self, c_pad):
        """(scale, shift) of the moving-statistics normalisation over `c_pad` columns (zero
        beyond this layer's channels), recomputed on the current stream: the epilogue form
        ConvLayer.forward_bn_infer hands to the convolution kernels."""
        buf = getattr(self, '_fold', None)
        if buf is None or buf.shape[1] != c_pad:
            buf = self._fold = torch.zeros(2, c_pad, dtype=torch.float32, device=self.scratch.device)
        N.call('seg_batchnorm_fold', N.ptr(self.moving_mean), N.ptr(self.moving_var), self.eps,
               N.ptr(self.beta.value()), self.c, c_pad, N.ptr(buf[0]), N.ptr(buf[1]),
               N.stream_ptr())
        return buf[0], buf[1]

    def pool_infer(self, x, y, k):
        """y = batch_norm(maxpool_k(x)) with the moving statistics: bit-identical to pooling
        the normalised tensor (seg_maxpool_bn_infer), which is then never materialised."""
        N.set_tag(self.name)
        N.note_work(0, x.numel() * 2.0 + y.numel() * 2.0)
        N.call('seg_maxpool_bn_infer', N.vref(x), k, N.ptr(self.moving_mean),
               N.ptr(self.moving_var), self.eps, N.ptr(self.beta.value()), self.c, N.vref(y),
               N.stream_ptr())

    def backward(self, dy, x, dx, relu_mask=True):
        """dx = BN-grad(dy) masked by the ReluGrad of the layer that produced x;
        dbeta accumulated into the grad buffer."""
        st = N.stream_ptr()
        s = self.scratch
        fill_zero(s[5])
        c = self.c
        dbeta = self.beta.grad()       # zero at step start (Adam zeroes the grad buffer)
        N.call('seg_batchnorm_bwd_reduce', N.vref(dy[..., :c]), N.vref(x[..., :c]), N.ptr(s[2]),
               N.ptr(s[3]), N.ptr(dbeta), N.ptr(s[5]), st)
        count = x.shape[0] * x.shape[1] * x.shape[2]
        N.call('seg_batchnorm_bwd_apply', N.vref(dy[..., :c]), N.vref(x[..., :c]), N.ptr(s[2]),
               N.ptr(s[3]), N.ptr(dbeta), N.ptr(s[5]), count, 1 if relu_mask else 0,
               N.vref(dx[..., :c]), st)
