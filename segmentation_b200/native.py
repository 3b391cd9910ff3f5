"""ctypes binding of libsegb200.so (the C ABI declared in include/segb200.h).

PyTorch is used only for device memory (`tensor.data_ptr()`) and streams
(`torch.cuda.current_stream().cuda_stream`); no torch types cross the ABI.
There is NO fallback: if the library is missing or a call fails, we raise.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsegb200.so')

IMPL_UMMA, IMPL_SIMT = 0, 1
EPI_BIAS, EPI_RELU, EPI_OUT_F32, EPI_RELU_MASK = 1, 2, 4, 8


class SegError(RuntimeError):
    status = 0


E_UNSUPPORTED = -6    # SEG_E_UNSUPPORTED: the entry does not take this geometry (nothing ran)


class SegView(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p),
                ('n', ctypes.c_int32), ('h', ctypes.c_int32), ('w', ctypes.c_int32),
                ('c', ctypes.c_int32),
                ('sn', ctypes.c_int64), ('sh', ctypes.c_int64), ('sw', ctypes.c_int64)]


class SegConvDesc(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in
                ('kh', 'kw', 'stride', 'pad_t', 'pad_l', 'pad_b', 'pad_r', 'cin', 'cout',
                 'cin_pad', 'cout_pad', 'flags', 'impl')]


class SegStageCtl(ctypes.Structure):
    """seg_stage_ctl: per-step scalars written by the staging launch (include/segb200.h)."""
    _fields_ = [('loss_sum', ctypes.c_void_p), ('host_ring', ctypes.c_void_p),
                ('publish_step', ctypes.c_int32), ('lr_t', ctypes.c_float),
                ('lr_t_dev', ctypes.c_void_p), ('step', ctypes.c_int32),
                ('step_dev', ctypes.c_void_p)]


_VP = ctypes.POINTER(SegView)
_DP = ctypes.POINTER(SegConvDesc)
_P = ctypes.c_void_p
_I32, _I64, _U32, _U64, _F = (ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64,
                              ctypes.c_float)

# name -> argtypes; every function returns int32 status unless noted
SIGNATURES = {
    'seg_version': [],
    'seg_device_check': [],
    'seg_set_option': [_I32, _I32],
    'seg_conv2d_fwd': [_DP, _VP, _VP, _P, _P, _VP, _P],
    'seg_conv2d_dgrad': [_DP, _VP, _P, _VP, _VP, _VP, _VP, _P],
    'seg_conv2d_dgrad_slice': [_DP, _VP, _P, ctypes.c_int32, _VP, _VP, _P],
    'seg_conv2d_wgrad': [_DP, _VP, _VP, _VP, _P, _P, _P],
    'seg_conv2d_pool_fwd': [_DP, _VP, _P, _P, _VP, _I32, _I32, _VP, _P, _P],
    'seg_conv2d_pool_wgrad': [_DP, _VP, _VP, _P, _VP, _P, _P, _P],
    'seg_conv2d_bn_pool_infer': [_DP, _VP, _P, _I32, _P, _P, _P, _F, _P, _VP, _P, _P],
    'seg_deconv2d_fwd': [_DP, _VP, _P, _P, _VP, _P],
    'seg_conv2d_fwd_affine': [_DP, _VP, _P, _P, _P, _P, _VP, _P],
    'seg_deconv2d_fwd_affine': [_DP, _VP, _P, _P, _P, _P, _VP, _P],
    'seg_deconv2d_dgrad': [_DP, _VP, _P, _VP, _VP, _P],
    'seg_deconv2d_wgrad': [_DP, _VP, _VP, _P, _P],
    'seg_bias_grad': [_VP, _P, _P],
    'seg_maxpool_fwd': [_VP, _I32, _I32, _VP, _P, _P],
    'seg_maxpool_bn_infer': [_VP, _I32, _P, _P, _F, _P, _I32, _VP, _P],
    'seg_maxpool_bwd': [_VP, _P, _I32, _I32, _VP, _I32, _I32, _VP, _VP, _P],
    'seg_bilinear_upsample_fwd': [_VP, _I32, _VP, _VP, _I32, _P],
    'seg_bilinear_upsample_bwd': [_VP, _I32, _I32, _VP, _VP, _P],
    'seg_maxpool_bwd_y': [_VP, _P, _I32, _I32, _VP, _I32, _I32, _VP, _VP, _VP, _P],
    'seg_maxpool_bwd2': [_VP, _VP, _P, _I32, _I32, _VP, _VP, _P],
    'seg_maxpool_bwd2_y': [_VP, _VP, _P, _I32, _I32, _VP, _VP, _VP, _P],
    'seg_relu_grad': [_VP, _VP, _VP, _P],
    'seg_resize_bilinear_fwd': [_VP, _VP, _P],
    'seg_resize_bilinear_bwd': [_VP, _VP, _P],
    'seg_batchnorm_stats': [_VP, _P, _P, _P],
    'seg_batchnorm_finalize': [_P, _P, _I64, _I32, _F, _F, _P, _P, _P, _P, _P],
    'seg_batchnorm_apply': [_VP, _P, _P, _P, _VP, _P],
    'seg_batchnorm_infer': [_VP, _P, _P, _F, _P, _VP, _P],
    'seg_batchnorm_fold': [_P, _P, _F, _P, _I32, _I32, _P, _P, _P],
    'seg_batchnorm_bwd_reduce': [_VP, _VP, _P, _P, _P, _P, _P],
    'seg_batchnorm_bwd_apply': [_VP, _VP, _P, _P, _P, _P, _I64, _I32, _VP, _P],
    'seg_dropout': [_VP, _U64, _U32, _F, _VP, _P],
    'seg_dropout_ex': [_VP, _U64, _U32, _U32, _P, _U32, _F, _VP, _P],
    'seg_upscore8_xent_fwd_bwd': [_VP, _VP, _P, _VP, _VP, _P, _P],
    'seg_softmax_xent_fwd_bwd': [_VP, _VP, _P, _VP, _P],
    'seg_sigmoid_argmax': [_VP, _P, _P, _P],
    'seg_mc_mean_var': [_P, _I32, _I64, _P, _P, _P],
    'seg_classmap_tail_infer': [_VP, _I32, _I32, _P, _I32, _I32, _P, _P, _P, _F, _P, _P, _I32, _I32,
                                _P, _I32, _P, _P, _P, _P],
    'seg_adam_chunk_elems': [],
    'seg_adam_multi': [_P, _P, _P, _P, _P, _P, _P, _P, _I32, _F, _P, _F, _F, _F, _F, _P],
    'seg_head1x1_xent': [_VP, _P, _I32, _P, _VP, _I32, _VP, _P, _VP, _P, _P, _P],
    'seg_pack_input': [_P, _I32, _VP, _P],
    'seg_stage_input': [_P, _I32, _I32, _I32, _I32, _P, _VP, _P, _I32, _P,
                        ctypes.POINTER(SegStageCtl), _P],
    'seg_pack_patches': [_P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP, _P],
    'seg_fill_zero': [_P, _I64, _P],
}

# self-test / micro-benchmark hooks: include/segb200_probes.h, libsegb200_probes.so
PROBE_SIGNATURES = {
    'seg_probe_umma': [_I32, _I32, _I32, _I32, _P, _P, _P, _P],
    'seg_probe_mma_rate': [_I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P],
    'seg_probe_red_rate': [_I32, _I32, _I32, _I32, _I32, _P, _P, _P],
}
PROBES_LIB_PATH = os.path.join(_HERE, 'libsegb200_probes.so')

_lib = None
_probes = None


def load_probes():
    """The probes library (same kernels + the seg_probe_* hooks); used by tests/ and tools/."""
    global _probes
    if _probes is None:
        if not os.path.exists(PROBES_LIB_PATH):
            raise SegError('%s not found: build it with `python -m segmentation_b200.build`'
                           % PROBES_LIB_PATH)
        lib = ctypes.CDLL(PROBES_LIB_PATH)
        for name, args in PROBE_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int32
        lib.seg_last_error_string.argtypes = []
        lib.seg_last_error_string.restype = ctypes.c_char_p
        _probes = lib
    return _probes


def call_probe(name, *args):
    lib = load_probes()
    status = getattr(lib, name)(*args)
    if status != 0:
        msg = lib.seg_last_error_string()
        raise SegError('%s failed (status %d): %s' % (name, status, msg.decode() if msg else ''))


def load():
    """Load the C-ABI library.  Fails loudly — there is no CPU/eager fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SegError('%s not found: build it with `python -m segmentation_b200.build` '
                       '(or __graft_entry__.build()); there is no fallback path' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ctypes.c_int32
    lib.seg_last_error_string.argtypes = []
    lib.seg_last_error_string.restype = ctypes.c_char_p
    lib.seg_last_kernel_name.argtypes = []
    lib.seg_last_kernel_name.restype = ctypes.c_char_p
    _lib = lib
    if os.environ.get('SEGB200_PDL', '1') == '0':
        lib.seg_set_option(OPT_PDL, 0)
    for env, key in (('SEGB200_TCONV_MIN_EFF', OPT_TILE_CONV_MIN_EFF),
                     ('SEGB200_TWGRAD_MIN_EFF', OPT_TILE_WGRAD_MIN_EFF),
                     ('SEGB200_POOL_ROWS', OPT_POOL_ROWS),
                     ('SEGB200_DEEP_B', OPT_DEEP_B_RING),
                     ('SEGB200_HCONV_ROWSTAGE', OPT_HALO_ROWSTAGE),
                     ('SEGB200_TWGRAD_TRED', OPT_WGRAD_TENSOR_RED),
                     ('SEGB200_FIRST_LAYER', OPT_FIRST_LAYER),
                     ('SEGB200_TAIL_MMA', OPT_TAIL_MMA),
                     ('SEGB200_WGRAD_MIN_TILES', OPT_WGRAD_MIN_TILES)):
        if env in os.environ:
            lib.seg_set_option(key, int(os.environ[env]))
    return lib


def check(status, what):
    if status != 0:
        msg = load().seg_last_error_string()
        err = SegError('%s failed (status %d): %s' % (what, status,
                                                     msg.decode() if msg else ''))
        err.status = status
        raise err


LAUNCHES = 0          # number of kernel-enqueueing C-ABI calls made so far
TIMELINE = None       # set to a list to record (name, tag, start_event, end_event) per call
_TAG = ['']


def set_tag(tag):
    _TAG[0] = tag


_WORK = [0.0, 0.0]    # algorithmic (flops, bytes) of the next call, noted by the engine wrappers


def note_work(flops, nbytes):
    """Algorithmic work of the next C-ABI call (conv 2*N*Ho*Wo*Cout*Cin*kh*kw,
    every tensor read or written once at its real channel count); recorded with the call's
    timeline entry when a timeline is being taken, otherwise dropped."""
    if TIMELINE is not None:
        _WORK[0], _WORK[1] = float(flops), float(nbytes)


# seg_set_option keys: test / A-B switches over the plan selection (include/segb200.h)
OPT_HALO_CONV, OPT_HALO_ROW_ALIGN, OPT_TILE_CONV, OPT_TILE_CONV_MIN_EFF = 1, 2, 3, 4
OPT_TILE_WGRAD, OPT_TILE_WGRAD_MIN_EFF = 5, 6
OPT_PDL = 7           # programmatic dependent launch of the hot-path kernels (default on)
OPT_WGRAD_MIN_TILES = 9  # pixel tiles per CTA below which the weight-gradient grid is narrowed
OPT_POOL_ROWS = 11     # row-mapped max-pool kernels (default on)
OPT_DEEP_B_RING = 12   # halo / spatial-tile conv: streamed-weight ring as deep as shared memory allows
OPT_HALO_ROWSTAGE = 14 # halo kernel: one filter row (3 taps) per streamed weight stage
OPT_WGRAD_TENSOR_RED = 15  # spatial-tile weight gradient: TMA tensor reduce-add epilogue (default on)
OPT_FIRST_LAYER = 16   # first-layer kernel on the (R,G,B,1) staged input (default on)
OPT_SM_LIMIT = 17      # SMs the persistent tile kernels size their grids for (0 = all)
OPT_TAIL_MMA = 18      # class-map tail: transposed conv on warp-level MMAs (default on)


def set_option(key, value):
    check(load().seg_set_option(key, value), 'seg_set_option')


def call(name, *args):
    global LAUNCHES
    LAUNCHES += 1
    if TIMELINE is None:
        check(getattr(load(), name)(*args), name)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(load(), name)(*args), name)
    e1.record()
    TIMELINE.append((name, _TAG[0], e0, e1, kernel_family(name), _WORK[0], _WORK[1]))
    _WORK[0] = _WORK[1] = 0.0


_FAMILY_OF_CALL = (('seg_maxpool', 'pool'), ('seg_adam', 'adam'), ('seg_head1x1', 'head'),
                   ('seg_softmax', 'loss'), ('seg_pack', 'pack'), ('seg_stage', 'pack'), ('seg_bias_grad', 'bias_grad'),
                   ('seg_batchnorm', 'batchnorm'), ('seg_dropout', 'dropout'),
                   ('seg_bilinear', 'bilinear'), ('seg_resize', 'resize'))
_CONV_CALLS = ('seg_conv2d_', 'seg_deconv2d_')


def kernel_family(call_name):
    """Kernel family a C-ABI call ran on: for the convolution family the tile kernel the
    plan picked (fconv / tconv / hconv / igemm / twgrad / wgrad, from seg_last_kernel_name;
    'simt' for the CUDA-core kernels), else a name derived from the entry point."""
    if call_name.startswith(_CONV_CALLS):
        import re
        k = load().seg_last_kernel_name().decode()
        m = re.search(r'(fconv|tconv|hconv|igemm|twgrad|wgrad)_kernel', k)
        return m.group(1) if m else 'simt'
    for prefix, fam in _FAMILY_OF_CALL:
        if call_name.startswith(prefix):
            return fam
    return 'other'


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def view(t):
    """SegView of a 4-D NHWC torch tensor (any strides, channel stride 1).
    Slicing the tensor in torch (crops, channel slices) yields the crop / concat
    views the C ABI understands."""
    if t is None:
        return None
    assert t.dim() == 4, 'NHWC tensor expected'
    assert t.stride(3) == 1 or t.shape[3] == 1, 'channel stride must be 1'
    return SegView(t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.shape[3],
                   t.stride(0), t.stride(1), t.stride(2))


def vref(t):
    return ctypes.byref(view(t)) if t is not None else None
