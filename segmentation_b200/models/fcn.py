"""FCNModel — mirror of /root/reference/models/fcn.py (FCN-32s/16s/8s, Long et al.
2015 as the reference wires it), executed by hand-written sm_100a kernels.

Encoder (reference `models/fcn.py:106-130`): 5x (3x3 SAME conv + ReLU -> 2x2/2
pool), 1x1 conv6 / conv7 (n_kernels*32 channels) and conv_fr -> n_classes.
Decoders (`:133-220`): 1x1 score convs on pool3 / pool4, x2 bilinear transposed
convs with skip ADDs, final x8 / x16 / x32 bilinear transposed conv.
Quirks kept: conv_fr and pool*_score keep slim's default ReLU (`:128,159,192,195`);
the bilinear filters are constants (`:139`) so they have no gradient.

The `[k,k,C,C]` bilinear filter bank is channel-diagonal, so the transposed conv
is executed as a depthwise separable bilinear kernel fused with the skip add
(`seg_bilinear_upsample_fwd`), not as a dense C x C GEMM.
Input sizes must be multiples of 32 (then every crop_or_pad in the reference graph
is the identity, as at the BASELINE configuration).
"""
import os

import torch

from .. import engine as E
from .. import native as N
from .basemodel import BaseModel, ExecBase

BF16 = torch.bfloat16


class FCNModel(BaseModel):
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
                 fcn_type='32s',
                 seed=0):
        super(FCNModel, self).__init__(
            sess=sess, mode=mode, log_dir=log_dir, dataset=dataset, bayesian=bayesian,
            save_dir=save_dir, n_classes=n_classes, input_dims=input_dims,
            test_dataset=test_dataset, input_channel=input_channel, load_snapshot=load_snapshot,
            learning_rate=learning_rate, load_snapshot_from=load_snapshot_from,
            adversarial_training=adversarial_training)
        self.model_name = 'FCN'
        self.n_kernels = n_kernels
        if fcn_type not in ('32s', '16s', '8s'):
            raise Exception('MODE ERROR')        # reference prints 'MODE ERROR' (fcn.py:103)
        self.fcn_type = fcn_type
        if self.input_dims[0] % 32 or self.input_dims[1] % 32:
            raise Exception('FCNModel: input_dims must be multiples of 32')
        # optimizer groups in variable order: conv1-3 | conv4-5 | conv6, conv7, conv_fr, score
        # convs (58 % of the parameters).  The backward pass completes them back to front, so
        # the (all-reduce +) Adam of the 1x1 head runs beside the encoder's backward and only
        # the small conv1-3 update is exposed at the end of the step.
        self.opt_splits = tuple(s for s in os.environ.get('SEGB200_FCN_OPT_SPLITS',
                                                          'conv4,conv6').split(',') if s)
        self._finish_init(seed)
        self.y_hat = self.y_hat_sig = self.output = None
        self.inference_ops = ['y_hat_sig', 'output']

    def _build_layers(self, gen):
        nk, st, nc = self.n_kernels, self.store, self.n_classes
        L = self.layers = {}

        def conv(name, cin, cout, k):
            L[name] = E.ConvLayer(st, name, 'conv', k, 1, 'SAME', cin, cout, True, gen)

        conv('conv1', self.input_channel, nk, 3)
        conv('conv2', nk, nk * 2, 3)
        conv('conv3', nk * 2, nk * 4, 3)
        conv('conv4', nk * 4, nk * 8, 3)
        conv('conv5', nk * 8, nk * 8, 3)
        conv('conv6', nk * 8, nk * 32, 1)
        conv('conv7', nk * 32, nk * 32, 1)
        conv('conv_fr', nk * 32, nc, 1)
        if self.fcn_type == '8s':
            conv('fcn8s/pool3_score', nk * 4, nc, 1)
            conv('fcn8s/pool4_score', nk * 8, nc, 1)
        elif self.fcn_type == '16s':
            conv('fcn16s/pool4_score', nk * 8, nc, 1)

    def _make_exec(self, batch, training):
        return _FCNExec(self, batch, training)

    def model(self, input_op, reuse=False):
        x = self._to_device(input_op, torch.float32)
        ex = self._get_exec(x.shape[0], False)
        ex.stage(x, None)
        ex.forward()
        return ex.logits


class _FCNExec(ExecBase):
    def __init__(self, model, B, training):
        dev, nk, nc = model.device, model.n_kernels, model.n_classes
        H, W = model.input_dims
        L = model.layers
        ncp = L['conv_fr'].cout_pad
        self.ncp = ncp
        self.act, self.amax = {}, {}

        def buf(name, h, w, c):
            self.act[name] = torch.zeros(B, h, w, c, dtype=BF16, device=dev)
            return self.act[name]

        self.x4 = self.first_layer_x4(model, L['conv1'])
        buf('x', H, W, 4 if self.x4 else L['conv1'].cin_pad)
        chans = [nk, nk * 2, nk * 4, nk * 8, nk * 8]
        h, w = H, W
        for i, c in enumerate(chans, 1):
            buf('conv%d' % i, h, w, c)
            h, w = h // 2, w // 2
            buf('pool%d' % i, h, w, c)
            self.amax['pool%d' % i] = torch.zeros(B, h, w, c, dtype=torch.uint8, device=dev)
        buf('conv6', h, w, nk * 32)
        buf('conv7', h, w, nk * 32)
        buf('conv_fr', h, w, ncp)
        t = model.fcn_type
        if t in ('8s', '16s'):
            buf('pool4_score', H // 16, W // 16, ncp)
            buf('fuse4', H // 16, W // 16, ncp)
        if t == '8s':
            buf('pool3_score', H // 8, W // 8, ncp)
            buf('fuse3', H // 8, W // 8, ncp)
        self._init_io(model, B, H, W, H, W, nc, ncp, training)
        if training:
            self.g = {name: torch.zeros_like(tt) for name, tt in self.act.items() if name != 'x'}
            self.g['logits'] = self.dlogits
            if t in ('8s', '16s'):
                self.g['pool4_b'] = torch.zeros_like(self.act['pool4'])
            if t == '8s':
                self.g['pool3_b'] = torch.zeros_like(self.act['pool3'])

    def v(self, t):
        """logical n_classes channels of a class-score tensor."""
        return t[..., :self.m.n_classes]

    def forward_for_step(self):
        # train step: FCN-8s ends with one launch for upscore x8 + loss + their gradient
        self.forward(fused_loss=os.environ.get('SEGB200_FCN_FUSED_LOSS', '1') != '0')

    def forward(self, fused_loss=False):
        """`fused_loss` (train steps of FCN-8s): stop at the fused score map; loss() then runs
        seg_upscore8_xent_fwd_bwd, which also produces the gradient backward() starts from -
        the full-resolution logits (m.y_hat) are NOT refreshed by such a step."""
        m, L, A, impl, v = self.m, self.m.layers, self.act, self.m.impl, self.v
        self._fused_loss = bool(fused_loss and self.training and m.fcn_type == '8s' and
                                m.n_classes <= 32)
        self.pack()
        src = A['x']
        # conv1 + pool1 as one launch (first-layer kernel): conv1's full-resolution
        # activation has no other consumer (models/fcn.py:110-117)
        self.fuse_pool1 = self.x4 and L['conv1'].pool_fusable(A['x'], A['pool1'], impl)
        for i in range(1, 6):
            if i == 1 and self.fuse_pool1:
                L['conv1'].forward_pool(A['x'], A['pool1'], self.amax['pool1'])
            else:
                L['conv%d' % i].forward(src, A['conv%d' % i], impl=impl)
                E.maxpool_fwd(A['conv%d' % i], A['pool%d' % i], self.amax['pool%d' % i])
            src = A['pool%d' % i]
        L['conv6'].forward(A['pool5'], A['conv6'], impl=impl)
        L['conv7'].forward(A['conv6'], A['conv7'], impl=impl)
        L['conv_fr'].forward(A['conv7'], v(A['conv_fr']), impl=impl)
        t = m.fcn_type
        if t == '32s':
            E.bilinear_upsample_fwd(v(A['conv_fr']), 32, self.logits)
        elif t == '16s':
            L['fcn16s/pool4_score'].forward(A['pool4'], v(A['pool4_score']), impl=impl)
            E.bilinear_upsample_fwd(v(A['conv_fr']), 2, v(A['fuse4']), add=v(A['pool4_score']))
            E.bilinear_upsample_fwd(v(A['fuse4']), 16, self.logits)
        else:
            L['fcn8s/pool3_score'].forward(A['pool3'], v(A['pool3_score']), impl=impl)
            L['fcn8s/pool4_score'].forward(A['pool4'], v(A['pool4_score']), impl=impl)
            E.bilinear_upsample_fwd(v(A['conv_fr']), 2, v(A['fuse4']), add=v(A['pool4_score']))
            E.bilinear_upsample_fwd(v(A['fuse4']), 2, v(A['fuse3']), add=v(A['pool3_score']))
            if not self._fused_loss:
                E.bilinear_upsample_fwd(v(A['fuse3']), 8, self.logits)
        m.y_hat = self.logits

    def loss(self, with_grad):
        if not (getattr(self, '_fused_loss', False) and with_grad):
            return super(_FCNExec, self).loss(with_grad)
        self.zero_loss()
        E.upscore8_xent(self.v(self.act['fuse3']), self.mask_view(), self.loss_sum,
                        self.v(self.g['fuse3']))

    def backward(self):
        m, L, A, G, impl, v = self.m, self.m.layers, self.act, self.g, self.m.impl, self.v
        side = self.side if self.use_side else None

        def bw(name, *args, **kw):
            # weight gradients on the side stream (they only feed the optimizer)
            L[name].backward(*args, impl=impl, side=side, **kw)
            self.layer_done(name)

        t = m.fcn_type
        dl = v(G['logits'])
        if t == '32s':
            E.bilinear_upsample_bwd(dl, 32, v(G['conv_fr']), mask=v(A['conv_fr']))
        elif t == '16s':
            E.bilinear_upsample_bwd(dl, 16, v(G['fuse4']))
            E.relu_grad(v(G['fuse4']), v(A['pool4_score']), v(G['pool4_score']))
            E.bilinear_upsample_bwd(v(G['fuse4']), 2, v(G['conv_fr']), mask=v(A['conv_fr']))
            bw('fcn16s/pool4_score', A['pool4'], G['pool4_score'], dx=G['pool4_b'])
        else:
            if not getattr(self, '_fused_loss', False):
                E.bilinear_upsample_bwd(dl, 8, v(G['fuse3']))
            E.relu_grad(v(G['fuse3']), v(A['pool3_score']), v(G['pool3_score']))
            E.bilinear_upsample_bwd(v(G['fuse3']), 2, v(G['fuse4']))
            E.relu_grad(v(G['fuse4']), v(A['pool4_score']), v(G['pool4_score']))
            E.bilinear_upsample_bwd(v(G['fuse4']), 2, v(G['conv_fr']), mask=v(A['conv_fr']))
            bw('fcn8s/pool3_score', A['pool3'], G['pool3_score'], dx=G['pool3_b'])
            bw('fcn8s/pool4_score', A['pool4'], G['pool4_score'], dx=G['pool4_b'])
        bw('conv_fr', A['conv7'], G['conv_fr'], dx=G['conv7'], mask=A['conv7'])
        bw('conv7', A['conv6'], G['conv7'], dx=G['conv6'], mask=A['conv6'])
        bw('conv6', A['pool5'], G['conv6'], dx=G['pool5'])
        for i in range(5, 0, -1):
            conv, pool = 'conv%d' % i, 'pool%d' % i
            second = G.get(pool + '_b')
            if i == 1 and getattr(self, 'fuse_pool1', False):
                # pool1 backward + conv1 weight gradient in one launch
                N.set_tag(conv)
                L[conv].wgrad_pool(A['x'], G[pool], self.amax[pool], A[pool])
                self.layer_done(conv)
                continue
            if second is not None:
                E.maxpool_bwd2(G[pool], second, self.amax[pool], G[conv], mask=A[conv],
                               pooled=A[pool])
            else:
                E.maxpool_bwd(G[pool], self.amax[pool], G[conv], mask=A[conv], pooled=A[pool])
            if i > 1:
                bw(conv, A['pool%d' % (i - 1)], G[conv], dx=G['pool%d' % (i - 1)])
            else:
                bw(conv, A['x'], G[conv], dx=None)
        if side is not None:
            side.join()
