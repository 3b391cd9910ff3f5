"""DeconvModel — mirror of /root/reference/models/deconvolution.py (generic
conv / batch-norm / pool encoder + transposed-conv / batch-norm decoder with a
bilinear resize in the middle), executed by hand-written sm_100a kernels.

Graph (reference `models/deconvolution.py:101-178`):
  conv5x5/s2 SAME -> bn1 -> pool2/2 -> conv3 VALID -> bn2 -> [drop] -> pool3/3 ->
  conv3 -> bn3 -> pool3/3 -> conv3 -> bn4 -> [drop] -> deconv5/s2 -> bn5 -> [drop] ->
  deconv5/s2 -> bn6 -> deconv5/s2 -> bn7 -> resize_bilinear(H/2) -> deconv2/s2 ->
  bn8 -> crop_or_pad(H) -> conv3 SAME (no activation).
Quirks kept: every conv / transposed conv keeps slim's default ReLU (incl.
deconv3_0, `:166`), BN comes AFTER the ReLU and has no gamma (`:116`), dropout
(`bayesian=True`) is active in training AND inference (`:128-129,143-144,153-154`:
slim.dropout without is_training).

This is the only reference model whose MC-dropout placement is reference-defined;
`infer_mc` runs T stochastic passes of one tile as one batch and returns the
mean / variance maps (BASELINE config 5).
"""
import os

import torch

from .. import engine as E
from .. import native as N
from .basemodel import BaseModel, ExecBase

BF16 = torch.bfloat16


class DeconvModel(BaseModel):
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
                 autoencoder=False,
                 adversarial_training=False,
                 seed=0):
        super(DeconvModel, self).__init__(
            sess=sess, mode=mode, log_dir=log_dir, dataset=dataset, bayesian=bayesian,
            save_dir=save_dir, n_classes=n_classes, input_dims=input_dims,
            autoencoder=autoencoder, test_dataset=test_dataset, input_channel=input_channel,
            load_snapshot=load_snapshot, learning_rate=learning_rate,
            load_snapshot_from=load_snapshot_from, adversarial_training=adversarial_training)
        if autoencoder:
            raise Exception('autoencoder mode is outside the segmentation hot path (SURVEY §2 #7)')
        self.model_name = 'deconvolution'
        self.n_kernels = n_kernels
        if self.input_dims[0] % 2 or self.input_dims[1] % 2:
            raise Exception('DeconvModel: input_dims must be even')
        self._finish_init(seed)
        self.y_hat = self.y_hat_sig = self.output = None
        self.inference_ops = ['y_hat_sig', 'output']

    def _build_layers(self, gen):
        nk, st, nc = self.n_kernels, self.store, self.n_classes
        L = self.layers = {}

        def conv(name, k, s, padding, cin, cout, relu=True):
            L[name] = E.ConvLayer(st, name, 'conv', k, s, padding, cin, cout, relu, gen)

        def deconv(name, k, cin, cout):
            L[name] = E.ConvLayer(st, name, 'deconv', k, 2, 'VALID', cin, cout, True, gen)

        def bn(name, c):
            L[name] = E.BatchNorm(st, name, c)

        # conv1_0: 5x5 / stride 2 / SAME on the raw RGB input: the 5x5x3 patch is packed into 75
        # (of 80) channels and the layer runs as a 1x1 conv (engine.PatchConvLayer): 25 taps of
        # 16 zero-padded channels through TMA-im2col cost 2.96 ms at 32 x 1024^2, this ~1 ms
        if os.environ.get('SEGB200_PATCH_DECONV', '1') == '1':
            L['conv1_0'] = E.PatchConvLayer(st, 'conv1_0', 5, 2, 'SAME', self.input_channel, nk,
                                            True, gen)
        else:
            conv('conv1_0', 5, 2, 'SAME', self.input_channel, nk)
        bn('bn1', nk)
        conv('conv2_0', 3, 1, 'VALID', nk, nk * 2); bn('bn2', nk * 2)
        conv('conv3_0', 3, 1, 'VALID', nk * 2, nk * 4); bn('bn3', nk * 4)
        conv('conv4_0', 3, 1, 'VALID', nk * 4, nk * 8); bn('bn4', nk * 8)
        deconv('deconv1_0', 5, nk * 8, nk * 2); bn('bn5', nk * 2)
        deconv('deconv2_0', 5, nk * 2, nk); bn('bn6', nk)
        deconv('deconv2_1', 5, nk, nk); bn('bn7', nk)
        deconv('deconv3_0', 2, nk, nc); bn('bn8', nc)
        conv('conv_out', 3, 1, 'SAME', nc, nc, relu=False)

    def _make_exec(self, batch, training):
        return _DeconvExec(self, batch, training)

    def model(self, input_op, reuse=False, training=True):
        """Forward graph; `training` selects batch statistics vs moving statistics
        in the batch-norm layers (reference `models/deconvolution.py:101,116`)."""
        x = self._to_device(input_op, torch.float32)
        if training:
            # batch statistics: an executor with the unfused first layer (the inference
            # executor evaluates conv1_0 + bn1 + pool1 with the moving statistics in one launch)
            key = (x.shape[0], 'forward-batchstats')
            if key not in self._exec:
                self._exec[key] = _DeconvExec(self, x.shape[0], False, fuse_head=False)
            ex = self._exec[key]
        else:
            ex = self._get_exec(x.shape[0], False)
        ex.stage(x, None)
        ex.forward(bn_training=training)
        return ex.logits

    def infer_mc(self, imgs, passes=16, seed=None, pass_offset=0, return_probs=True):
        """T stochastic passes of one tile as one batch (dropout sites of the
        reference graph, Philox stream (pass_offset+t)*8+site), mean / variance of
        the sigmoid probabilities.  Returns [mean, var, probs[T]]."""
        x = self._to_device(imgs, torch.float32)
        assert x.shape[0] == 1, 'infer_mc takes one tile'
        ex = self._get_exec(passes, False)
        ex.stage(x.expand(passes, -1, -1, -1), None)
        ex.forward(dropout=(self.mc_seed if seed is None else seed, pass_offset), per_image=True,
                   fused_tail=True)
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


class _DeconvExec(ExecBase):
    SITES = {'bn2': 0, 'bn4': 1, 'bn5': 2}

    def __init__(self, model, B, training, fuse_head=True):
        dev, nk, nc = model.device, model.n_kernels, model.n_classes
        H, W = model.input_dims
        L = model.layers
        self.act, self.amax = {}, {}
        self.ncp = L['deconv3_0'].cout_pad

        def buf(name, h, w, c):
            self.act[name] = torch.zeros(B, h, w, c, dtype=BF16, device=dev)
            return self.act[name]

        def pair(conv, bn, h, w, c):
            buf(conv, h, w, c)
            buf(bn, h, w, c)
            return h, w

        def pool(name, src, k):
            t = self.act[src]
            h, w = (t.shape[1] - k) // k + 1, (t.shape[2] - k) // k + 1
            buf(name, h, w, t.shape[3])
            self.amax[name] = torch.zeros(B, h, w, t.shape[3], dtype=torch.uint8, device=dev)
            return h, w

        self.patch_l1 = isinstance(L['conv1_0'], E.PatchConvLayer)
        # inference: conv1_0 + bn1 + pool1 as ONE launch of the first-layer kernel on the
        # (R,G,B,1) staged input (no patch tensor, no full-resolution conv1_0 / bn1 tensors)
        self.head_fused = (fuse_head and not training and model.impl == N.IMPL_UMMA and
                           model.input_channel == 3
                           and L['conv1_0'].cout_pad == 32 and (H // 2) % 2 == 0 and (W // 2) % 2 == 0
                           and os.environ.get('SEGB200_FUSED_HEAD1', '1') != '0')
        self.x4 = self.head_fused
        if self.head_fused:
            buf('x', H, W, 4)
            h, w = -(-H // 2), -(-W // 2)
        elif self.patch_l1:
            h, w = L['conv1_0'].patch_out_hw(H, W)
            buf('x', h, w, L['conv1_0'].cin_pad)
        else:
            buf('x', H, W, L['conv1_0'].cin_pad)
            h, w = L['conv1_0'].out_hw(H, W)
        if self.head_fused:
            h, w = h // 2, w // 2
            buf('pool1', h, w, nk)
        else:
            pair('conv1_0', 'bn1', h, w, nk)
            h, w = pool('pool1', 'bn1', 2)
        h, w = pair('conv2_0', 'bn2', h - 2, w - 2, nk * 2)
        h, w = pool('pool2', 'bn2', 3)
        h, w = pair('conv3_0', 'bn3', h - 2, w - 2, nk * 4)
        h, w = pool('pool3', 'bn3', 3)
        h, w = pair('conv4_0', 'bn4', h - 2, w - 2, nk * 8)
        h, w = pair('deconv1_0', 'bn5', *L['deconv1_0'].out_hw(h, w), nk * 2)
        h, w = pair('deconv2_0', 'bn6', *L['deconv2_0'].out_hw(h, w), nk)
        h, w = pair('deconv2_1', 'bn7', *L['deconv2_1'].out_hw(h, w), nk)
        buf('resize', H // 2, W // 2, nk)
        pair('deconv3_0', 'bn8', H, W, self.ncp)
        self._init_io(model, B, H, W, H, W, nc, L['conv_out'].cout_pad, training)
        if training:
            self.g = {name: torch.zeros_like(t) for name, t in self.act.items() if name != 'x'}
            self.g['logits'] = self.dlogits
        self.step_seed = 0

    def _pack_now(self):
        if self.head_fused:
            E.stage_input(self.x_f32, self.act['x'])
        elif self.patch_l1:
            self.m.layers['conv1_0'].pack(self.x_f32, self.act['x'])
        else:
            E.pack_input(self.x_f32, self.act['x'])

    def forward(self, bn_training=None, dropout=None, per_image=False, fused_tail=False):
        m, L, A, impl = self.m, self.m.layers, self.act, self.m.impl
        if bn_training is None:
            bn_training = self.training
        # bayesian=True without explicit streams: fresh masks every step, stream offset =
        # global_step read from the device scalar m.step_dev at run time, so the captured
        # CUDA graph of a train step does not bake one step's masks in
        self._from_dev = dropout is None and m.bayesian
        if self._from_dev:
            dropout = (m.mc_seed, 0)
        self._dropout = dropout
        self._per_image = per_image

        def bn(name, src):
            L[name].forward(A[src], A[name], training=bn_training)
            if dropout is not None and name in self.SITES:
                self._drop(A[name], self.SITES[name])

        self.pack()
        if self.head_fused and not bn_training:
            E.conv_bn_pool_infer(L['conv1_0'], L['bn1'], A['x'], A['pool1'])
        else:
            assert not self.head_fused, 'inference executor: batch statistics not available'
            L['conv1_0'].forward(A['x'], A['conv1_0'], impl=impl); bn('bn1', 'conv1_0')
            E.maxpool_fwd(A['bn1'], A['pool1'], self.amax['pool1'], 2, 2)
        # Inference executors never read the un-normalised activations again, so the moving-
        # statistics batch-norms ride along with a neighbour (SEGB200_FUSE_BN=0: unfused):
        # in front of a pool they are applied to the pooled tensor (bit-identical, see
        # seg_maxpool_bn_infer), elsewhere they are the producing layer's epilogue.
        fuse_bn = (not bn_training and not self.training and impl == N.IMPL_UMMA and
                   os.environ.get('SEGB200_FUSE_BN', '1') != '0')

        def layer_bn(layer, bn_name, src, limpl):
            """layer -> ReLU -> batch-norm (-> dropout at the model's sites)"""
            if fuse_bn and L[layer].forward_bn_infer(A[src], A[bn_name], L[bn_name], impl=limpl):
                if dropout is not None and bn_name in self.SITES:
                    self._drop(A[bn_name], self.SITES[bn_name])
                return
            L[layer].forward(A[src], A[layer], impl=limpl); bn(bn_name, layer)

        def layer_bn_pool(layer, bn_name, src, pool_name, k):
            """layer -> ReLU -> batch-norm (-> dropout) -> k x k max-pool"""
            if fuse_bn and not (dropout is not None and bn_name in self.SITES):
                L[layer].forward(A[src], A[layer], impl=impl)
                L[bn_name].pool_infer(A[layer], A[pool_name], k)
                return
            L[layer].forward(A[src], A[layer], impl=impl); bn(bn_name, layer)
            E.maxpool_fwd(A[bn_name], A[pool_name], self.amax[pool_name], k, k)

        layer_bn_pool('conv2_0', 'bn2', 'pool1', 'pool2', 3)
        layer_bn_pool('conv3_0', 'bn3', 'pool2', 'pool3', 3)
        layer_bn('conv4_0', 'bn4', 'pool3', impl)
        # 5x5 stride-2 transposed convs: four output-parity classes, each a stride-1
        # correlation with a 3x3 / 3x2 / 2x3 / 2x2 sub-kernel on the halo-tile tcgen05 kernel
        dimpl = N.IMPL_SIMT if os.environ.get('SEGB200_DECONV5', 'umma') == 'simt' else impl
        layer_bn('deconv1_0', 'bn5', 'bn4', dimpl)
        layer_bn('deconv2_0', 'bn6', 'bn5', dimpl)
        layer_bn('deconv2_1', 'bn7', 'bn6', dimpl)
        self._tail_done = False
        if (fused_tail and not bn_training and impl == N.IMPL_UMMA and 2 <= m.n_classes <= 4 and
                A['bn7'].shape[3] == 32 and os.environ.get('SEGB200_FUSED_TAIL', '1') != '0'):
            # inference: resize -> deconv3_0 -> bn8 -> conv_out -> sigmoid/argmax in ONE launch;
            # the class maps never travel through HBM as 16-channel padded tensors
            E.classmap_tail_infer(A['bn7'], self.H // 2, self.W // 2, L['deconv3_0'], L['bn8'],
                                  L['conv_out'], self.logits, self.probs, self.labelmap)
            m.y_hat, m.y_hat_sig, m.output = self.logits, self.probs, self.labelmap
            self._tail_done = True
            return
        E.resize_bilinear_fwd(A['bn7'], A['resize'])
        nc = m.n_classes
        L['deconv3_0'].forward(A['resize'], A['deconv3_0'][..., :nc], impl=impl)
        bn('bn8', 'deconv3_0')
        L['conv_out'].forward(A['bn8'], self.logits, impl=impl, out_f32=True)
        m.y_hat = self.logits

    def head(self):
        if getattr(self, '_tail_done', False):          # the fused tail produced both already
            return self.probs, self.labelmap
        return super(_DeconvExec, self).head()

    def infer(self, x):
        self.stage(x, None)
        self.forward(fused_tail=True)
        return self.head()

    def _drop(self, t, site):
        """One launch per site: stream (off [+ image]) * 8 + site (+ 8 * global_step, read
        from the device, when the streams follow the training step)."""
        seed, off = self._dropout
        E.dropout_ex(t, t, seed, off * 8 + site, per_image_step=8 if self._per_image else 0,
                     step_dev=self.m.step_dev if self._from_dev else None,
                     step_mul=8 if self._from_dev else 0)

    def backward(self):
        m, L, A, G, impl = self.m, self.m.layers, self.act, self.g, self.m.impl
        nc = m.n_classes
        S = N.IMPL_SIMT
        # 5x5 / stride-2 transposed convs: dgrad is a strided correlation over dz and wgrad the
        # matching strided im2col GEMM - both on the tcgen05 im2col kernels
        D5 = S if os.environ.get('SEGB200_DECONV5', 'umma') == 'simt' else impl

        def bn_bwd(name, src):
            if self._dropout is not None and name in self.SITES:
                self._drop(G[name], self.SITES[name])         # same mask on the gradient
            L[name].backward(G[name], A[src], G[src], relu_mask=True)

        L['conv_out'].backward(A['bn8'], G['logits'], dx=G['bn8'], impl=impl)
        bn_bwd('bn8', 'deconv3_0')
        L['deconv3_0'].backward(A['resize'], G['deconv3_0'], dx=G['resize'], impl=impl,
                                dz_bias=G['deconv3_0'][..., :nc])
        E.resize_bilinear_bwd(G['resize'], G['bn7'])
        bn_bwd('bn7', 'deconv2_1')
        L['deconv2_1'].backward(A['bn6'], G['deconv2_1'], dx=G['bn6'], impl=D5)
        bn_bwd('bn6', 'deconv2_0')
        L['deconv2_0'].backward(A['bn5'], G['deconv2_0'], dx=G['bn5'], impl=D5)
        bn_bwd('bn5', 'deconv1_0')
        L['deconv1_0'].backward(A['bn4'], G['deconv1_0'], dx=G['bn4'], impl=D5)
        bn_bwd('bn4', 'conv4_0')
        L['conv4_0'].backward(A['pool3'], G['conv4_0'], dx=G['pool3'], impl=impl)
        E.maxpool_bwd(G['pool3'], self.amax['pool3'], G['bn3'], 3, 3)
        bn_bwd('bn3', 'conv3_0')
        L['conv3_0'].backward(A['pool2'], G['conv3_0'], dx=G['pool2'], impl=impl)
        E.maxpool_bwd(G['pool2'], self.amax['pool2'], G['bn2'], 3, 3)
        bn_bwd('bn2', 'conv2_0')
        L['conv2_0'].backward(A['pool1'], G['conv2_0'], dx=G['pool1'], impl=impl)
        E.maxpool_bwd(G['pool1'], self.amax['pool1'], G['bn1'], 2, 2)
        bn_bwd('bn1', 'conv1_0')
        # strided 5x5 wgrad: CUDA-core correlation kernel (the tcgen05 wgrad covers
        # stride 1 and k == stride)
        L['conv1_0'].backward(A['x'], G['conv1_0'], dx=None, impl=S)
