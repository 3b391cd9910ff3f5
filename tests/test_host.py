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


def test_optimizer_groups_partition_the_parameters():
    """Optimizer groups (engine.ParamStore.chunk_range + BaseModel._init_opt_groups logic):
    contiguous parameter slices in TF variable order whose Adam chunk ranges partition the
    chunk table; boundaries coincide with the data-parallel bucket boundaries."""
    from segmentation_b200 import parallel as P
    st = E.ParamStore(torch.device('cpu'))
    gen = np.random.default_rng(0)
    names = ['conv1_1', 'conv3_1', 'conv5_1', 'conv5_2', 'upconv1', 'conv6_1', 'output']
    for n in names:
        E.ConvLayer(st, n, 'deconv' if n.startswith('up') else 'conv', 2 if n.startswith('up') else 3,
                    1, 'VALID', 48, 80, True, gen)
    st.finalize()
    bounds = P.bucket_boundaries(st, ('conv3_1', 'conv5_1', 'upconv1'))
    assert bounds[0] == 0 and bounds[-1] == st.numel and len(bounds) == 5
    ranges = [st.chunk_range(a, b) for a, b in zip(bounds[:-1], bounds[1:])]
    assert ranges[0][0] == 0 and ranges[-1][1] == st.nchunks
    for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
        assert a1 == b0 and a0 < a1
    # every chunk of a group addresses a parameter inside the group's slice
    ch = st.chunks.view(-1, 2).tolist()
    seg = st.segments.view(-1, 6).tolist()
    for (c0, c1), (lo, hi) in zip(ranges, zip(bounds[:-1], bounds[1:])):
        for si, first in ch[c0:c1]:
            assert lo <= seg[si][0] + first < hi


def test_patch_conv_layer_shadow_is_the_hwio_tensor_read_as_a_matrix():
    """PatchConvLayer: same TF variable [k,k,cin,cout]; the bf16 shadow is that tensor read as
    [k*k*cin][cout] with rows padded to a multiple of 16 (a 1x1 conv over packed patches)."""
    st = E.ParamStore(torch.device('cpu'))
    gen = np.random.default_rng(1)
    lay = E.PatchConvLayer(st, 'conv1_1', 3, 1, 'VALID', 3, 32, True, gen)
    st.finalize()
    lay.init_values()
    st.refresh_shadow()
    assert lay.w.value().shape == (3, 3, 3, 32) and lay.w.shadow().shape == (1, 1, 32, 32)
    assert (lay.k, lay.cin, lay.cin_pad, lay.cout_pad) == (1, 27, 32, 32)
    assert lay.patch_out_hw(256, 256) == (254, 254)
    sh = lay.w.shadow().float()[0, 0]
    assert torch.equal(sh[:27], lay.w.value().reshape(27, 32).to(torch.bfloat16).float())
    assert float(sh[27:].abs().max()) == 0.0
    assert st.segments.view(-1, 6)[0].tolist() == [0, 864, 32, 32, 27, 32]
    big = E.PatchConvLayer(E.ParamStore(torch.device('cpu')), 'conv1_0', 5, 2, 'SAME', 3, 32)
    assert (big.cin, big.cin_pad) == (75, 80) and big.patch_out_hw(1024, 1024) == (512, 512)


def test_array_image_mask_dataset_matches_reference_preprocessing():
    """utils/datasets.py (device-side input path): the batches it hands the model
    and its host restatement of reference utils/datasets.py:176-190 (x/255, uint8(mask/255),
    joint crop)."""
    import numpy as np
    from segmentation_b200.utils.datasets import ArrayImageMaskDataSet, load_images
    g = np.random.default_rng(0)
    imgs = g.integers(0, 256, (5, 40, 48, 3), dtype=np.uint8)
    msk = g.choice(np.array([0, 255], dtype=np.uint8), (5, 40, 48, 1))
    ds = ArrayImageMaskDataSet(imgs, msk, batch_size=3, crop_size=32, seed=1, pinned=False)
    x, y, c = ds.next_batch()
    assert tuple(x.shape) == (3, 40, 48, 3) and x.dtype.__str__() == 'torch.uint8'
    assert tuple(y.shape) == (3, 40, 48, 1) and tuple(c.shape) == (3, 2)
    assert int(c[:, 0].max()) <= 8 and int(c[:, 1].max()) <= 16 and int(c.min()) >= 0
    idx, crop = np.array([0, 4, 2]), np.array([[0, 0], [8, 16], [3, 5]], dtype=np.int32)
    xr, yr = ds.reference_batch(idx, crop)
    assert xr.dtype == np.float32 and xr.shape == (3, 32, 32, 3) and yr.shape == (3, 32, 32, 1)
    assert np.array_equal(xr[1], imgs[4, 8:40, 16:48].astype(np.float32) / np.float32(255))
    assert set(np.unique(yr)) <= {0, 1}
    assert np.array_equal(yr[2][..., 0] == 1, msk[2, 3:35, 5:37, 0] == 255)
    sel, cr = load_images(imgs, 4, 32, rng=np.random.default_rng(2))
    assert sel.shape == (4, 40, 48, 3) and cr.shape == (4, 2)


def test_fast_division_constants_of_the_first_layer_kernel():
    """csrc/fconv.cu computes (mul, shr) so that x // d == (x * mul) >> (32 + shr) for every
    x < 2^31 (pixel / tile index -> image, row, column on the device): the same construction
    in Python, checked at the divisors the models use and at the range ends."""
    def magic(d):
        if d == 1:
            return 0, 0
        l = 0
        while (1 << l) < d:
            l += 1
        p = 31 + l
        return ((1 << p) + d - 1) // d, p - 32

    import random
    rnd = random.Random(0)
    for d in [2, 3, 4, 7, 8, 62, 127, 254, 255, 256, 508, 512, 2048, 64516, 262144, 16129 * 4, 999983]:
        mul, shr = magic(d)
        assert mul < (1 << 32)
        xs = [0, 1, d - 1, d, d + 1, 2 * d - 1, (1 << 30) - 1, (1 << 30), (1 << 31) - 1]
        xs += [rnd.randrange(1 << 31) for _ in range(2000)]
        xs += [k * d + e for k in (1, 5, 1000, ((1 << 31) - 1) // d) for e in (-1, 0, 1) if 0 <= k * d + e < (1 << 31)]
        for x in xs:
            assert ((x * mul) >> 32) >> shr == x // d, (d, x)
