"""UNetModel — mirror of /root/reference/models/unet.py (Ronneberger et al. 2015
as the reference wires it), executed by hand-written sm_100a kernels.

Graph (reference `models/unet.py:109-175`): 18x (3x3 VALID conv + bias + ReLU),
4x 2x2/2 max-pool, 4x learned 2x2/2 transposed conv (+ReLU, slim default),
4x centre-crop + concat [skip_crop, upconv], 1x1 head without activation.
Load-bearing quirks kept: pool1 consumes conv1_1 (not conv1_2, `:118-120`),
conv1_2 feeds only the last skip (`:161`).

Crops and concats cost no kernel: the consumer conv reads two TMA descriptors
(skip crop view | upconv output) and its dgrad writes two destinations.  The
gradient of conv1_2's output is zero outside the 72x72 skip crop, so conv1_2's
backward runs on that crop only (exact, not an approximation).
"""
import os

import numpy as np
import torch

from .. import engine as E
from .. import native as N
from .basemodel import BaseModel, ExecBase

BF16 = torch.bfloat16


class UNetModel(BaseModel):
    def __init__(self,
                 sess=None,
                 n_classes=2,
                 log_dir=None,
                 dataset=None,
                 save_dir=None,
                 bayesian=False,
                 input_dims=512,
                 mode='TRAINING',
                 input_channel=3,
                 test_dataset=None,
                 learning_rate=1e-4,
                 load_snapshot=None,
                 load_snapshot_from=None,
                 n_kernels=32,
                 adversarial_training=False,
                 seed=0):
        super(UNetModel, self).__init__(
            sess=sess, mode=mode, log_dir=log_dir, dataset=dataset, bayesian=bayesian,
            save_dir=save_dir, n_classes=n_classes, input_dims=input_dims,
            test_dataset=test_dataset, input_channel=input_channel, load_snapshot=load_snapshot,
            learning_rate=learning_rate, load_snapshot_from=load_snapshot_from,
            adversarial_training=adversarial_training)
        self.model_name = 'unet'
        self.IN_OUT_CROP = True
        self.n_kernels = n_kernels
        # conv1-2 | conv3-4 | bottleneck conv5_* (61 % of the bytes) | decoder: the group that
        # completes last (and whose all-reduce + Adam is exposed at the end of the step) is small
        self.opt_splits = tuple(os.environ.get('SEGB200_OPT_SPLITS', 'conv3_1,conv5_1,upconv1').split(','))
        self._finish_init(seed)
        self.y_hat = None          # logits of the most recent forward (device fp32 tensor)
        self.y_hat_sig = None
        self.output = None
        self.inference_ops = ['y_hat_sig', 'output']

    # ---------------------------------------------------------------- layers
    def _build_layers(self, gen):
        nk, st = self.n_kernels, self.store
        L = self.layers = {}

        def conv(name, cin, cout, k=3, relu=True):
            L[name] = E.ConvLayer(st, name, 'conv', k, 1, 'VALID', cin, cout, relu, gen)

        def up(name, cin, cout):
            L[name] = E.ConvLayer(st, name, 'deconv', 2, 2, 'VALID', cin, cout, True, gen)

        # first layer.  Optional (SEGB200_PATCH_L1=1): 3x3x3 patch packed into 27 (of 32)
        # channels + 1x1 conv (engine.PatchConvLayer).  Correct, but measured slower on the
        # current kernels (1.415 vs 1.248 ms/step): a K=32 GEMM leaves the im2col kernel's
        # LSU epilogue (~3k cycles per 128x32 tile) and 64-byte TMA rows exposed.
        if os.environ.get('SEGB200_PATCH_L1', '0') == '1':
            L['conv1_1'] = E.PatchConvLayer(st, 'conv1_1', 3, 1, 'VALID', self.input_channel, nk,
                                            True, gen)
        else:
            conv('conv1_1', self.input_channel, nk)
        conv('conv1_2', nk, nk)
        conv('conv2_1', nk, nk * 2); conv('conv2_2', nk * 2, nk * 2)
        conv('conv3_1', nk * 2, nk * 4); conv('conv3_2', nk * 4, nk * 4)
        conv('conv4_1', nk * 4, nk * 8); conv('conv4_2', nk * 8, nk * 8)
        conv('conv5_1', nk * 8, nk * 16); conv('conv5_2', nk * 16, nk * 16)
        up('upconv1', nk * 16, nk * 8)
        conv('conv6_1', nk * 16, nk * 8); conv('conv6_2', nk * 8, nk * 8)
        up('upconv2', nk * 8, nk * 4)
        conv('conv7_1', nk * 8, nk * 4); conv('conv7_2', nk * 4, nk * 4)
        up('upconv3', nk * 4, nk * 2)
        conv('conv8_1', nk * 4, nk * 2); conv('conv8_2', nk * 2, nk * 2)
        up('upconv4', nk * 2, nk)
        conv('conv9_1', nk * 2, nk); conv('conv9_2', nk, nk)
        conv('output', nk, self.n_classes, k=1, relu=False)

    def _make_exec(self, batch, training):
        return _UNetExec(self, batch, training)

    def model(self, input_op, reuse=False):
        """Forward graph on a fp32 NHWC batch (numpy or torch); returns the
        logits as a device fp32 tensor [B,H',W',n_classes]."""
        x = self._to_device(input_op, torch.float32)
        ex = self._get_exec(x.shape[0], False)
        ex.stage(x, None)
        ex.forward()
        return ex.logits

    def infer_mc(self, imgs, passes=16, seed=None, pass_offset=0, return_probs=True):
        """Bayesian mode (BASELINE config 5): `passes` stochastic forward passes of
        ONE tile executed as one batch, MC-dropout (keep .5) after conv2_2,
        conv4_2, conv6_2 (build-defined placement — the reference's UNetModel
        accepts `bayesian` but never reads it, `models/unet.py:31`), then mean
        and variance of the sigmoid probabilities over the passes.
        Returns [mean [H',W',C], var [H',W',C], probs [T,H',W',C]]."""
        x = self._to_device(imgs, torch.float32)
        assert x.shape[0] == 1, 'infer_mc takes one tile'
        ex = self._get_exec(passes, False)
        ex.stage(x.expand(passes, -1, -1, -1), None)
        ex.forward(dropout=(self.mc_seed if seed is None else seed, pass_offset))
        probs, _ = ex.head()
        mean = torch.empty_like(probs[0])
        var = torch.empty_like(probs[0])
        E.mc_mean_var(probs, mean, var)
        torch.cuda.current_stream().synchronize()
        # return_probs=False: mean and variance maps only (BASELINE config 5's output); the T
        # probability maps are T x larger than both and their device->host copy dominates
        out = [mean.cpu().numpy(), var.cpu().numpy()]
        if return_probs:
            out.append(probs.cpu().numpy())
        return out


class _UNetExec(ExecBase):
    """Buffers + kernel schedule for one (batch size, mode)."""

    MC_SITES = {'conv2_2': 0, 'conv4_2': 1, 'conv6_2': 2}

    def __init__(self, model, B, training):
        self.m, self.B, self.training = model, B, training
        dev, nk = model.device, model.n_kernels
        H, W = model.input_dims
        self.H, self.W = H, W
        L = model.layers
        self.act = {}
        self.amax = {}

        def buf(name, h, w, c, dtype=BF16):
            self.act[name] = torch.zeros(B, h, w, c, dtype=dtype, device=dev)
            return self.act[name]

        self.patch_l1 = isinstance(L['conv1_1'], E.PatchConvLayer)
        self.x4 = self.first_layer_x4(model, L['conv1_1'])
        if self.patch_l1:
            buf('x', H - 2, W - 2, L['conv1_1'].cin_pad)
        else:
            buf('x', H, W, 4 if self.x4 else L['conv1_1'].cin_pad)
        # ---- encoder geometry
        h, w = H - 2, W - 2
        buf('conv1_1', h, w, nk)
        buf('conv1_2', h - 2, w - 2, nk)
        ph, pw = h // 2, w // 2
        buf('pool1', ph, pw, nk)
        self.amax['pool1'] = torch.zeros(B, ph, pw, nk, dtype=torch.uint8, device=dev)
        for i in range(2, 6):
            c = nk * 2 ** (i - 1)
            h, w = ph - 2, pw - 2
            buf('conv%d_1' % i, h, w, c)
            buf('conv%d_2' % i, h - 2, w - 2, c)
            if i < 5:
                ph, pw = (h - 2) // 2, (w - 2) // 2
                buf('pool%d' % i, ph, pw, c)
                self.amax['pool%d' % i] = torch.zeros(B, ph, pw, c, dtype=torch.uint8, device=dev)
        # ---- decoder geometry
        self.skip_of = {1: 'conv4_2', 2: 'conv3_2', 3: 'conv2_2', 4: 'conv1_2'}
        self.crop = {}
        below = self.act['conv5_2']
        for j in range(1, 5):
            c = nk * 2 ** (4 - j)
            uh, uw = below.shape[1] * 2, below.shape[2] * 2
            buf('upconv%d' % j, uh, uw, c)
            sk = self.act[self.skip_of[j]]
            self.crop[j] = ((sk.shape[1] - uh) // 2, (sk.shape[2] - uw) // 2, uh, uw)
            buf('conv%d_1' % (5 + j), uh - 2, uw - 2, c)
            below = buf('conv%d_2' % (5 + j), uh - 4, uw - 4, c)
        oh, ow = below.shape[1], below.shape[2]
        self._init_io(model, B, H, W, oh, ow, model.n_classes, L['output'].cout_pad, training)
        if training:
            self._alloc_grads()
        # training: the 1x1 head, the loss and the head's backward are one kernel
        # optional backward schedules (both measured neutral-to-slower, 1.175 -> 1.18-1.19
        # ms/step: the step is bound by total SM time, not by the main stream's chain):
        # decoder stages whose concat input gradient is split into a decoder half (main
        # stream) and a skip half (skip stream); conv1_2's backward pass right behind
        # conv9_1's instead of at the end
        self.c12_crop = os.environ.get('SEGB200_C12_CROP', '1') != '0'
        self.tail_split = max(1, int(os.environ.get('SEGB200_TAIL_SPLIT', '1')))
        self.skip_split = set(int(v) for v in
                              os.environ.get('SEGB200_SKIP_SPLIT', '').split(',') if v.strip())
        self.early_conv1_2 = os.environ.get('SEGB200_EARLY_C12', '0') != '0'
        self.fused_head = (training and model.n_classes <= 4 and nk in (16, 32) and
                           model.impl == N.IMPL_UMMA and
                           os.environ.get('SEGB200_FUSED_HEAD', '1') != '0')

    def _pack_now(self):
        if self.patch_l1:
            self.m.layers['conv1_1'].pack(self.x_f32, self.act['x'])
        else:
            super(_UNetExec, self)._pack_now()

    # ------------------------------------------------------------- buffers
    def skip_view(self, j):
        y0, x0, h, w = self.crop[j]
        return self.act[self.skip_of[j]][:, y0:y0 + h, x0:x0 + w, :]

    def _alloc_grads(self):
        dev, B = self.m.device, self.B
        self.g = {}

        def gbuf(name, like=None, shape=None):
            shp = tuple(like.shape) if like is not None else shape
            self.g[name] = torch.zeros(shp, dtype=BF16, device=dev)
            return self.g[name]

        self.g['logits'] = self.dlogits
        for name, t in self.act.items():
            if name in ('x', 'conv1_2'):
                continue
            gbuf(name, like=t)
        for j in range(1, 5):
            y0, x0, h, w = self.crop[j]
            gbuf('skip%d' % j, shape=(B, h, w, self.act[self.skip_of[j]].shape[3]))
        y0, x0, h, w = self.crop[4]
        gbuf('conv1_1_part', shape=(B, h + 2, w + 2, self.act['conv1_1'].shape[3]))

    # ------------------------------------------------------------- forward
    def forward(self, dropout=None):
        m, L, A, impl = self.m, self.m.layers, self.act, self.m.impl
        self.pack()

        def conv(name, src, x2=None):
            L[name].forward(src, A[name], x2=x2, impl=impl)
            if dropout is not None and name in self.MC_SITES:
                seed, off = dropout            # one launch, one Philox stream per MC pass
                E.dropout_ex(A[name], A[name], seed, off * 8 + self.MC_SITES[name],
                             per_image_step=8)

        # conv1_1 + pool1 as ONE launch where the first-layer kernel applies: the full-resolution
        # conv1_1 activation is needed again only inside the window conv1_2 reads
        y0, x0, h, w = self.crop[4]
        self.fuse_pool1 = (self.x4 and not self.patch_l1 and
                           L['conv1_1'].pool_fusable(A['x'], A['pool1'], impl))
        if self.fuse_pool1:
            if self.c12_crop:
                win, wy, wx = A['conv1_1'][:, y0:y0 + h + 2, x0:x0 + w + 2, :], y0, x0
            else:
                win, wy, wx = A['conv1_1'], 0, 0
            L['conv1_1'].forward_pool(A['x'], A['pool1'], self.amax['pool1'], y_win=win,
                                      win_y0=wy, win_x0=wx)
        else:
            conv('conv1_1', A['x'])
        # conv1_2 feeds only the last skip connection (reference models/unet.py:118-120,161):
        # it runs on the side stream, filling the SMs the small deep layers leave idle
        # 0: off; i: fork before stage i+1.  With conv1_1 + pool1 fused nothing is left for
        # conv1_2 to overlap with at the start: inline is faster (0.942 vs 0.954 ms/step)
        fwd_at = int(os.environ.get('SEGB200_FWD_SIDE', '0' if self.fuse_pool1 else '1'))
        fwd_side = self.side if (self.use_side and dropout is None and fwd_at > 0) else None

        def conv1_2():
            if self.c12_crop:
                # only the centre crop of conv1_2's output has a consumer (concat4; pool1
                # takes conv1_1, reference models/unet.py:118-120,159-161): evaluate the
                # layer on that window - identical skip tensor, 1/12 of the work
                y0, x0, h, w = self.crop[4]
                src = A['conv1_1'][:, y0:y0 + h + 2, x0:x0 + w + 2, :]
                dst = A['conv1_2'][:, y0:y0 + h, x0:x0 + w, :]
                run = lambda: L['conv1_2'].forward(src, dst, impl=impl)
            else:
                run = lambda: conv('conv1_2', A['conv1_1'])
            if fwd_side is not None:
                with fwd_side.fork():
                    run()
            else:
                run()

        if fwd_at <= 1:
            conv1_2()
        if not self.fuse_pool1:
            E.maxpool_fwd(A['conv1_1'], A['pool1'], self.amax['pool1'])
        for i in range(2, 6):
            if fwd_at == i:
                conv1_2()
            conv('conv%d_1' % i, A['pool%d' % (i - 1)])
            conv('conv%d_2' % i, A['conv%d_1' % i])
            if i < 5:
                E.maxpool_fwd(A['conv%d_2' % i], A['pool%d' % i], self.amax['pool%d' % i])
        below = A['conv5_2']
        for j in range(1, 5):
            up = 'upconv%d' % j
            L[up].forward(below, A[up], impl=impl)
            if j == 4 and fwd_side is not None:
                fwd_side.join()                   # conv1_2 is needed from here on
            conv('conv%d_1' % (5 + j), self.skip_view(j), x2=A[up])
            conv('conv%d_2' % (5 + j), A['conv%d_1' % (5 + j)])
            below = A['conv%d_2' % (5 + j)]
        if not self.fused_head:                   # else loss() runs the head as well
            L['output'].forward(below, self.logits, impl=impl, out_f32=True)
        m.y_hat = self.logits

    def loss(self, with_grad):
        if not (self.fused_head and with_grad):
            if self.fused_head:                   # loss only, on a training executor
                self.m.layers['output'].forward(self.act['conv9_2'], self.logits, impl=self.m.impl,
                                                out_f32=True)
            return super(_UNetExec, self).loss(with_grad)
        self.zero_loss()
        E.head1x1_xent(self.act['conv9_2'], self.m.layers['output'], self.mask_view(), self.logits,
                       self.loss_sum, self.g['conv9_2'])

    # ------------------------------------------------------------ backward
    def backward(self):
        L, A, G, impl = self.m.layers, self.act, self.g, self.m.impl
        nc = self.m.n_classes
        side = self.side if self.use_side else None
        skipside = self.skipside if (self.use_side and impl == N.IMPL_UMMA) else None
        # conv1_2: output gradient lives only on the skip crop -> run on the crop
        y0, x0, h, w = self.crop[4]
        x_win = A['conv1_1'][:, y0:y0 + h + 2, x0:x0 + w + 2, :]

        def bw(name, *args, **kw):
            # weight gradients go to the side stream; an optimizer group may be complete
            # once this layer's kernels are enqueued
            L[name].backward(*args, impl=impl, side=side, **kw)
            self.layer_done(name)

        # head (no activation): dz = dlogits; already done by loss() when the head is fused
        if not self.fused_head:
            bw('output', A['conv9_2'], G['logits'], dx=G['conv9_2'], mask=A['conv9_2'],
               dz_bias=G['logits'][..., :nc])
        for j in range(4, 0, -1):
            c1, c2, up = 'conv%d_1' % (5 + j), 'conv%d_2' % (5 + j), 'upconv%d' % j
            below = 'conv%d_2' % (4 + j) if j > 1 else 'conv5_2'
            bw(c2, A[c1], G[c2], dx=G[c1], mask=A[c1])
            # conv over the virtual concat [skip_crop | upconv]: two dgrad destinations.
            # The skip gradient is ReLU-masked later by the pool backward that merges
            # it (j<4); for j==4 conv1_2 has no other consumer, so mask it here.
            sk = self.skip_view(j)
            if skipside is not None and j in self.skip_split:
                # the skip half of the input gradient is not needed before the encoder's
                # backward pass: only the decoder half stays on the critical path
                L[c1].backward(sk, G[c1], dx=None, x2=A[up], impl=impl, side=side)   # dW only
                with skipside.fork():
                    L[c1].dgrad_slice(G[c1], 0, G['skip%d' % j], mask=sk if j == 4 else None,
                                      impl=impl)
                L[c1].dgrad_slice(G[c1], sk.shape[3], G[up], mask=A[up], impl=impl)
                self.layer_done(c1)               # after every kernel that reads its weights
            else:
                bw(c1, sk, G[c1], dx=G['skip%d' % j], x2=A[up], dx2=G[up],
                   mask=sk if j == 4 else None, mask2=A[up])
            if j == 4 and skipside is not None and self.early_conv1_2:
                # conv1_2 feeds only this skip connection (models/unet.py:118-120,161): its
                # whole backward pass can run here, beside the rest of the decoder
                with skipside.fork():
                    L['conv1_2'].backward(x_win, G['skip4'], dx=G['conv1_1_part'], impl=impl)
            bw(up, A[below], G[up], dx=G[below], mask=A[below])
        # encoder
        bw('conv5_2', A['conv5_1'], G['conv5_2'], dx=G['conv5_1'], mask=A['conv5_1'])
        for i in range(5, 1, -1):
            c1, c2, pool = 'conv%d_1' % i, 'conv%d_2' % i, 'pool%d' % (i - 1)
            if i < 5:
                # dZ(conv i_2) = relu_mask(pool_i backward + skip gradient at the crop)
                j = 5 - i
                y0, x0, _, _ = self.crop[j]
                if skipside is not None:
                    skipside.join()               # G[skip j] may come from the skip stream
                E.maxpool_bwd(G['pool%d' % i], self.amax['pool%d' % i], G[c2],
                              add=G['skip%d' % j], add_y0=y0, add_x0=x0, mask=A[c2],
                              pooled=A['pool%d' % i])
                bw(c2, A[c1], G[c2], dx=G[c1], mask=A[c1])
            bw(c1, A[pool], G[c1], dx=G[pool])
        y0, x0, h, w = self.crop[4]
        if skipside is not None:
            skipside.join()
        fused = getattr(self, 'fuse_pool1', False)
        if not (skipside is not None and self.early_conv1_2):
            # fused first layer: conv1_2's input gradient is ReLU-masked here (not by pool1's
            # backward, which never materialises) ...
            bw('conv1_2', x_win, G['skip4'], dx=G['conv1_1_part'], mask=x_win if fused else None)
        else:
            assert not fused
            self.layer_done('conv1_2')
        if fused:
            # ... and enters conv1_1's weight gradient as its own linear term: the plain
            # first-layer weight gradient on the window (input = a crop view of the staged x)
            L['conv1_1'].backward(A['x'][:, y0:y0 + h + 4, x0:x0 + w + 4, :], G['conv1_1_part'],
                                  dx=None, impl=impl, side=side)
        # The step ends with pool1's backward -> conv1_1's weight gradient -> Adam of the
        # first group, a chain nothing else is left to overlap.  Optionally (tail_split
        # batch slices) the weight gradient of slice k runs on the side stream while the
        # pool backward of slice k+1 runs here; partial sums meet in the fp32 reductions.
        if getattr(self, 'fuse_pool1', False):
            # pool1's backward evaluated inside the operand producer of conv1_1's weight
            # gradient: one launch, the full-resolution gradient never exists
            N.set_tag('conv1_1')
            L['conv1_1'].wgrad_pool(A['x'], G['pool1'], self.amax['pool1'], A['pool1'])
            self.layer_done('conv1_1')
            if side is not None:
                side.join()
            return
        ts = self.tail_split if (side is not None and self.B % max(self.tail_split, 1) == 0) else 1
        nb = self.B // ts
        for k in range(ts):
            sl = slice(k * nb, (k + 1) * nb)
            E.maxpool_bwd(G['pool1'][sl], self.amax['pool1'][sl], G['conv1_1'][sl],
                          add=G['conv1_1_part'][sl], add_y0=y0, add_x0=x0, mask=A['conv1_1'][sl],
                          pooled=A['pool1'][sl])
            if k < ts - 1:
                L['conv1_1'].backward(A['x'][sl], G['conv1_1'][sl], dx=None, impl=impl, side=side)
            else:
                bw('conv1_1', A['x'][sl], G['conv1_1'][sl], dx=None)
        if side is not None:
            side.join()
