"""CPU tests of the host-side logic (no kernels run): parameter store layout,
padding arithmetic, layer geometry, dataset duck-type handling."""
import numpy as np
import torch

from segmentation_b200 import engine as E


def test_same_pad_and_out_sizes():
    assert E.same_pad(1024, 5, 2) == (1, 2)
    assert E.same_pad(512, 3, 1) == (1, 1)
    st = E.ParamStore(torch.device('cpu'))
    gen = np.random.default_rng(0)
    c = E.ConvLayer(st, 'c', 'conv', 5, 2, 'SAME', 3, 32, True, gen)
    assert c.out_hw(1024, 1024) == (512, 512) and c.cin_pad == 16 and c.cout_pad == 32
    v = E.ConvLayer(st, 'v', 'conv', 3, 1, 'VALID', 32, 64, True, gen)
    assert v.out_hw(127, 127) == (125, 125)
    d = E.ConvLayer(st, 'd', 'deconv', 5, 2, 'VALID', 256, 64, True, gen)
    assert d.out_hw(25, 25) == (53, 53)
    u = E.ConvLayer(st, 'u', 'deconv', 2, 2, 'VALID', 512, 256, True, gen)
    assert u.out_hw(8, 8) == (16, 16)
    ds = c.desc(1024, 1024, 3, 0)
    assert (ds.pad_t, ds.pad_l, ds.pad_b, ds.pad_r) == (1, 1, 2, 2)


def test_param_store_layout_tf_names_and_shadow_padding():
    st = E.ParamStore(torch.device('cpu'))
    gen = np.random.default_rng(0)
    a = E.ConvLayer(st, 'conv1_1', 'conv', 3, 1, 'VALID', 3, 32, True, gen)
    b = E.ConvLayer(st, 'upconv1', 'deconv', 2, 2, 'VALID', 64, 20, True, gen)
    st.finalize()
    a.init_values(); b.init_values()
    st.refresh_shadow()
    assert list(st.params) == ['conv1_1/weights', 'conv1_1/biases', 'upconv1/weights',
                               'upconv1/biases']
    assert a.w.value().shape == (3, 3, 3, 32) and a.w.shadow().shape == (3, 3, 16, 32)
    assert b.w.value().shape == (2, 2, 20, 64) and b.w.shadow().shape == (2, 2, 32, 64)
    assert st.numel == 3 * 3 * 3 * 32 + 32 + 2 * 2 * 20 * 64 + 20
    # xavier-uniform limits [TF-sem 12]
    lim = np.sqrt(6.0 / (9 * 3 + 9 * 32))
    assert float(a.w.value().abs().max()) <= lim and float(a.w.value().abs().max()) > 0.8 * lim
    assert float(a.b.value().abs().max()) == 0.0
    sh = a.w.shadow().float()
    assert torch.equal(sh[:, :, :3], a.w.value().to(torch.bfloat16).float())
    assert float(sh[:, :, 3:].abs().max()) == 0.0
    seg = st.segments.view(-1, 6)
    assert seg[0].tolist() == [0, 864, 32, 32, 3, 16]
    assert seg[2].tolist() == [896, 5120, 64, 64, 20, 32]
    assert st.shadow_offsets.tolist()[1] == -1 and st.shadow_offsets.tolist()[0] % 128 == 0
    sd = st.state_dict()
    st.master.zero_()
    st.load_state_dict(sd)
    assert torch.equal(a.w.value(), torch.from_numpy(sd['conv1_1/weights']))


def test_adam_lr_t_schedule():
    st = E.ParamStore(torch.device('cpu'))
    st.add('x', (4,))
    st.finalize()
    l1 = st.next_lr_t(1e-4)
    assert abs(l1 - 1e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)) < 1e-12 and st.step == 1
    l2 = st.next_lr_t(1e-4)
    assert abs(l2 - 1e-4 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)) < 1e-12
