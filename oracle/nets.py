"""Reference graph restatement (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Restates the three `model()` graphs of the reference on top of oracle.tf_ops:

  * `unet_forward`    <- /root/reference/models/unet.py:109-175
  * `fcn_forward`     <- /root/reference/models/fcn.py:93-220
  * `deconv_forward`  <- /root/reference/models/deconvolution.py:101-178

plus the intended loss / optimizer step (`models/basemodel.py:59-70,185-196,
357-369`) and inference head (`models/unet.py:76-79`).  Parameters live in an
ordered dict keyed by the slim variable names the reference would create
(`conv1_1/weights` [kh,kw,Cin,Cout], `upconv1/weights` [kh,kw,Cout,Cin], ...).

Load-bearing quirks kept on purpose: pool1 consumes conv1_1
not conv1_2 (`models/unet.py:118-120`); U-Net upconvs, FCN conv_fr and
pool*_score keep slim's default ReLU; BN after ReLU, no gamma; concat order
[skip_crop, upconv].
"""
from collections import OrderedDict

import numpy as np
import torch

from . import tf_ops as T


# --------------------------------------------------------------------------
# parameter construction
# --------------------------------------------------------------------------
def _add_conv(p, gen, name, k, cin, cout):
    p[name + '/weights'] = torch.from_numpy(T.xavier_uniform((k, k, cin, cout), gen))
    p[name + '/biases'] = torch.zeros(cout)


def _add_deconv(p, gen, name, k, cin, cout):
    # slim.convolution2d_transpose weights are [kh,kw,Cout,Cin]
    p[name + '/weights'] = torch.from_numpy(T.xavier_uniform((k, k, cout, cin), gen))
    p[name + '/biases'] = torch.zeros(cout)


def unet_params(n_kernels=32, n_classes=2, input_channel=3, seed=0):
    gen = np.random.default_rng(seed)
    nk, p = n_kernels, OrderedDict()
    _add_conv(p, gen, 'conv1_1', 3, input_channel, nk)
    _add_conv(p, gen, 'conv1_2', 3, nk, nk)
    _add_conv(p, gen, 'conv2_1', 3, nk, nk * 2)
    _add_conv(p, gen, 'conv2_2', 3, nk * 2, nk * 2)
    _add_conv(p, gen, 'conv3_1', 3, nk * 2, nk * 4)
    _add_conv(p, gen, 'conv3_2', 3, nk * 4, nk * 4)
    _add_conv(p, gen, 'conv4_1', 3, nk * 4, nk * 8)
    _add_conv(p, gen, 'conv4_2', 3, nk * 8, nk * 8)
    _add_conv(p, gen, 'conv5_1', 3, nk * 8, nk * 16)
    _add_conv(p, gen, 'conv5_2', 3, nk * 16, nk * 16)
    _add_deconv(p, gen, 'upconv1', 2, nk * 16, nk * 8)
    _add_conv(p, gen, 'conv6_1', 3, nk * 16, nk * 8)
    _add_conv(p, gen, 'conv6_2', 3, nk * 8, nk * 8)
    _add_deconv(p, gen, 'upconv2', 2, nk * 8, nk * 4)
    _add_conv(p, gen, 'conv7_1', 3, nk * 8, nk * 4)
    _add_conv(p, gen, 'conv7_2', 3, nk * 4, nk * 4)
    _add_deconv(p, gen, 'upconv3', 2, nk * 4, nk * 2)
    _add_conv(p, gen, 'conv8_1', 3, nk * 4, nk * 2)
    _add_conv(p, gen, 'conv8_2', 3, nk * 2, nk * 2)
    _add_deconv(p, gen, 'upconv4', 2, nk * 2, nk)
    _add_conv(p, gen, 'conv9_1', 3, nk * 2, nk)
    _add_conv(p, gen, 'conv9_2', 3, nk, nk)
    _add_conv(p, gen, 'output', 1, nk, n_classes)
    return p


def fcn_params(n_kernels=32, n_classes=21, input_channel=3, fcn_type='8s', seed=0):
    gen = np.random.default_rng(seed)
    nk, p = n_kernels, OrderedDict()
    _add_conv(p, gen, 'conv1', 3, input_channel, nk)
    _add_conv(p, gen, 'conv2', 3, nk, nk * 2)
    _add_conv(p, gen, 'conv3', 3, nk * 2, nk * 4)
    _add_conv(p, gen, 'conv4', 3, nk * 4, nk * 8)
    _add_conv(p, gen, 'conv5', 3, nk * 8, nk * 8)
    _add_conv(p, gen, 'conv6', 1, nk * 8, nk * 32)
    _add_conv(p, gen, 'conv7', 1, nk * 32, nk * 32)
    _add_conv(p, gen, 'conv_fr', 1, nk * 32, n_classes)
    if fcn_type == '8s':
        _add_conv(p, gen, 'fcn8s/pool3_score', 1, nk * 4, n_classes)
        _add_conv(p, gen, 'fcn8s/pool4_score', 1, nk * 8, n_classes)
    elif fcn_type == '16s':
        _add_conv(p, gen, 'fcn16s/pool4_score', 1, nk * 8, n_classes)
    return p


def deconv_params(n_kernels=32, n_classes=2, input_channel=3, seed=0):
    gen = np.random.default_rng(seed)
    nk, p = n_kernels, OrderedDict()

    def bn(name, c):
        p[name + '/beta'] = torch.zeros(c)
        p[name + '/moving_mean'] = torch.zeros(c)
        p[name + '/moving_variance'] = torch.ones(c)

    _add_conv(p, gen, 'conv1_0', 5, input_channel, nk); bn('bn1', nk)
    _add_conv(p, gen, 'conv2_0', 3, nk, nk * 2); bn('bn2', nk * 2)
    _add_conv(p, gen, 'conv3_0', 3, nk * 2, nk * 4); bn('bn3', nk * 4)
    _add_conv(p, gen, 'conv4_0', 3, nk * 4, nk * 8); bn('bn4', nk * 8)
    _add_deconv(p, gen, 'deconv1_0', 5, nk * 8, nk * 2); bn('bn5', nk * 2)
    _add_deconv(p, gen, 'deconv2_0', 5, nk * 2, nk); bn('bn6', nk)
    _add_deconv(p, gen, 'deconv2_1', 5, nk, nk); bn('bn7', nk)
    _add_deconv(p, gen, 'deconv3_0', 2, nk, n_classes); bn('bn8', n_classes)
    _add_conv(p, gen, 'conv_out', 3, n_classes, n_classes)
    return p


TRAINABLE_SUFFIXES = ('/weights', '/biases', '/beta')


def trainable_names(p):
    return [k for k in p if k.endswith(TRAINABLE_SUFFIXES)]


# --------------------------------------------------------------------------
# layer helpers (slim defaults: ReLU unless activation_fn=None)
# --------------------------------------------------------------------------
def _conv(p, prec, name, x, stride=1, padding='VALID', relu=True):
    y = T.conv2d(x, prec.wt(p[name + '/weights']), p[name + '/biases'], stride, padding)
    if relu:
        y = torch.relu(y)
    return prec.act(y)


def _deconv(p, prec, name, x, stride, padding='VALID', relu=True):
    y = T.conv2d_transpose(x, prec.wt(p[name + '/weights']), p[name + '/biases'], stride, padding)
    if relu:
        y = torch.relu(y)
    return prec.act(y)


# --------------------------------------------------------------------------
# U-Net   (/root/reference/models/unet.py:109-175)
# --------------------------------------------------------------------------
def unet_forward(p, x, prec=T.FP32, taps=None, dropout=None):
    """x: fp32 NHWC.  Returns logits [N,H',W',n_classes] (fp32).

    `dropout=(seed, pass_offset)` enables the build-defined U-Net MC-dropout
    placement (the reference's UNetModel accepts `bayesian` but never reads
    it): keep 0.5 after conv2_2, conv4_2, conv6_2 with Philox
    streams pass_offset*8 + {0,1,2}."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    def drop(t, site):
        if dropout is None:
            return t
        seed, off = dropout
        return prec.act(T.dropout(t, seed, off * 8 + site))

    x = prec.act(x)
    net = tap('conv1_1', _conv(p, prec, 'conv1_1', x))
    net1_2 = tap('conv1_2', _conv(p, prec, 'conv1_2', net))
    net2_0 = tap('pool1', T.max_pool(net, 2, 2))            # quirk: pools conv1_1
    net2_1 = tap('conv2_1', _conv(p, prec, 'conv2_1', net2_0))
    net2_2 = tap('conv2_2', drop(_conv(p, prec, 'conv2_2', net2_1), 0))
    net3_0 = tap('pool2', T.max_pool(net2_2, 2, 2))
    net3_1 = tap('conv3_1', _conv(p, prec, 'conv3_1', net3_0))
    net3_2 = tap('conv3_2', _conv(p, prec, 'conv3_2', net3_1))
    net4_0 = tap('pool3', T.max_pool(net3_2, 2, 2))
    net4_1 = tap('conv4_1', _conv(p, prec, 'conv4_1', net4_0))
    net4_2 = tap('conv4_2', drop(_conv(p, prec, 'conv4_2', net4_1), 1))
    net5_0 = tap('pool4', T.max_pool(net4_2, 2, 2))
    net5_1 = tap('conv5_1', _conv(p, prec, 'conv5_1', net5_0))
    net5_2 = tap('conv5_2', _conv(p, prec, 'conv5_2', net5_1))

    def up(name, below, skip):
        u = tap(name, _deconv(p, prec, name, below, 2, 'VALID'))
        t = u.shape[1]
        return torch.cat([T.crop_or_pad(skip, t, t), u], dim=-1)

    net6_0 = up('upconv1', net5_2, net4_2)
    net6_1 = tap('conv6_1', _conv(p, prec, 'conv6_1', net6_0))
    net6_2 = tap('conv6_2', drop(_conv(p, prec, 'conv6_2', net6_1), 2))
    net7_0 = up('upconv2', net6_2, net3_2)
    net7_1 = tap('conv7_1', _conv(p, prec, 'conv7_1', net7_0))
    net7_2 = tap('conv7_2', _conv(p, prec, 'conv7_2', net7_1))
    net8_0 = up('upconv3', net7_2, net2_2)
    net8_1 = tap('conv8_1', _conv(p, prec, 'conv8_1', net8_0))
    net8_2 = tap('conv8_2', _conv(p, prec, 'conv8_2', net8_1))
    net9_0 = up('upconv4', net8_2, net1_2)
    net9_1 = tap('conv9_1', _conv(p, prec, 'conv9_1', net9_0))
    net9_2 = tap('conv9_2', _conv(p, prec, 'conv9_2', net9_1))
    # 1x1 head, activation_fn=None; logits stay fp32 on the CUDA path too
    out = T.conv2d(net9_2, prec.wt(p['output/weights']), p['output/biases'], 1, 'VALID')
    return tap('output', out)


def unet_out_size(n):
    """256 -> 68, 512 -> 324 (printed by the reference at `models/unet.py:168`)."""
    n = n - 4
    for _ in range(4):
        n = n // 2 - 4
    for _ in range(4):
        n = n * 2 - 4
    return n


# --------------------------------------------------------------------------
# FCN   (/root/reference/models/fcn.py:93-220)
# --------------------------------------------------------------------------
def fcn_forward(p, x, fcn_type='8s', prec=T.FP32, taps=None):
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    H, W = x.shape[1], x.shape[2]
    x = prec.act(x)
    net = tap('conv1', _conv(p, prec, 'conv1', x, 1, 'SAME'))
    net = tap('pool1', T.max_pool(net))
    net = tap('conv2', _conv(p, prec, 'conv2', net, 1, 'SAME'))
    net = tap('pool2', T.max_pool(net))
    net = tap('conv3', _conv(p, prec, 'conv3', net, 1, 'SAME'))
    pool3 = tap('pool3', T.max_pool(net))
    net = tap('conv4', _conv(p, prec, 'conv4', pool3, 1, 'SAME'))
    pool4 = tap('pool4', T.max_pool(net))
    net = tap('conv5', _conv(p, prec, 'conv5', pool4, 1, 'SAME'))
    pool5 = tap('pool5', T.max_pool(net))
    net = tap('conv6', _conv(p, prec, 'conv6', pool5, 1, 'SAME'))
    net = tap('conv7', _conv(p, prec, 'conv7', net, 1, 'SAME'))
    net = tap('conv_fr', _conv(p, prec, 'conv_fr', net, 1, 'SAME'))   # ReLU kept (quirk)

    def up(t, f):
        return prec.act(T.bilinear_upsample(t, f))

    if fcn_type == '32s':
        out = up(net, 32)
    elif fcn_type == '16s':
        s4 = tap('pool4_score', _conv(p, prec, 'fcn16s/pool4_score', pool4, 1, 'SAME'))
        u = T.crop_or_pad(up(net, 2), s4.shape[1], s4.shape[1])     # (pool4_h, pool4_h) quirk
        out = up(prec.act(s4 + u), 16)
    elif fcn_type == '8s':
        s3 = tap('pool3_score', _conv(p, prec, 'fcn8s/pool3_score', pool3, 1, 'SAME'))
        s4 = tap('pool4_score', _conv(p, prec, 'fcn8s/pool4_score', pool4, 1, 'SAME'))
        u = T.crop_or_pad(up(net, 2), s4.shape[1], s4.shape[2])
        u = tap('fuse4', prec.act(s4 + u))
        u = T.crop_or_pad(up(u, 2), s3.shape[1], s3.shape[2])
        u = tap('fuse3', prec.act(s3 + u))
        out = T.bilinear_upsample(u, 8)                              # logits: fp32
    else:
        raise Exception('MODE ERROR')
    return tap('output', T.crop_or_pad(out, H, W))


# --------------------------------------------------------------------------
# DeconvModel   (/root/reference/models/deconvolution.py:101-178)
# --------------------------------------------------------------------------
def deconv_forward(p, x, training=True, bayesian=False, prec=T.FP32, taps=None,
                   dropout=(0, 0), new_stats=None):
    """`dropout=(seed, pass_offset)`; dropout sites are the reference's
    (`:128-129,143-144,153-154`), Philox streams pass_offset*8 + {0,1,2}."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    def bn(name, t):
        y, m, v = T.batch_norm(t, p[name + '/beta'], p[name + '/moving_mean'],
                               p[name + '/moving_variance'], training)
        if new_stats is not None:
            new_stats[name + '/moving_mean'] = m
            new_stats[name + '/moving_variance'] = v
        return tap(name, prec.act(y))

    def drop(t, site):
        if not bayesian:
            return t
        seed, off = dropout
        return prec.act(T.dropout(t, seed, off * 8 + site))

    H, W = x.shape[1], x.shape[2]
    x = prec.act(x)
    net = bn('bn1', tap('conv1_0', _conv(p, prec, 'conv1_0', x, 2, 'SAME')))
    net = tap('pool1', T.max_pool(net, 2, 2))
    net = bn('bn2', tap('conv2_0', _conv(p, prec, 'conv2_0', net)))
    net = drop(net, 0)
    net = tap('pool2', T.max_pool(net, 3, 3))
    net = bn('bn3', tap('conv3_0', _conv(p, prec, 'conv3_0', net)))
    net = tap('pool3', T.max_pool(net, 3, 3))
    net = bn('bn4', tap('conv4_0', _conv(p, prec, 'conv4_0', net)))
    net = drop(net, 1)
    net = bn('bn5', tap('deconv1_0', _deconv(p, prec, 'deconv1_0', net, 2)))
    net = drop(net, 2)
    net = bn('bn6', tap('deconv2_0', _deconv(p, prec, 'deconv2_0', net, 2)))
    net = bn('bn7', tap('deconv2_1', _deconv(p, prec, 'deconv2_1', net, 2)))
    net = tap('resize', prec.act(T.resize_bilinear(net, H // 2, W // 2)))
    net = bn('bn8', tap('deconv3_0', _deconv(p, prec, 'deconv3_0', net, 2)))
    net = T.crop_or_pad(net, H, W)
    out = T.conv2d(net, prec.wt(p['conv_out/weights']), p['conv_out/biases'], 1, 'SAME')
    return tap('output', out)


# --------------------------------------------------------------------------
# training step / inference (`models/basemodel.py:357-369`, `:527-531`)
# --------------------------------------------------------------------------
def loss_and_grads(forward, p, x, mask_u8, crop_mask=True):
    """One forward+backward.  Returns (loss, logits, {name: grad})."""
    names = trainable_names(p)
    leaves = OrderedDict((k, p[k].clone().requires_grad_(True)) for k in names)
    q = OrderedDict(p)
    q.update(leaves)
    logits = forward(q, x)
    y = mask_u8
    if crop_mask:   # `models/unet.py:71-72`: mask centre-cropped to the logits' H
        y = T.crop_or_pad(mask_u8, logits.shape[1], logits.shape[2])
    loss = T.softmax_xent_mean(logits, y)
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return loss.detach(), logits.detach(), OrderedDict(zip(names, grads))


class AdamState(object):
    def __init__(self, p):
        self.step = 0
        self.m = OrderedDict((k, torch.zeros_like(p[k])) for k in trainable_names(p))
        self.v = OrderedDict((k, torch.zeros_like(p[k])) for k in trainable_names(p))


def train_step(forward, p, state, x, mask_u8, lr=1e-4, crop_mask=True):
    """forward + backward + Adam; mutates p/state in place; returns loss."""
    loss, _, grads = loss_and_grads(forward, p, x, mask_u8, crop_mask)
    state.step += 1
    for k, g in grads.items():
        p[k], state.m[k], state.v[k] = T.adam_update(p[k], g, state.m[k], state.v[k],
                                                     state.step, lr)
    return float(loss)


def infer(forward, p, x):
    """[sigmoid(y_hat), float32(argmax(sigmoid(y_hat),3)[...,None])]."""
    with torch.no_grad():
        sig, lab = T.sigmoid_argmax(forward(p, x))
    return [sig.numpy(), lab.numpy()]
