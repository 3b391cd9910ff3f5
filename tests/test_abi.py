"""CPU tests: the C-ABI library loads without a GPU and exports every symbol that
include/segb200.h declares; the ctypes binding covers exactly that set."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'segb200.h')


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r'SEG_API\s+[\w\s\*]+?\b(seg_\w+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    from segmentation_b200 import build, native
    build.build()                       # no-op when the in-tree .so is current
    return ctypes.CDLL(native.LIB_PATH)


def test_header_declares_the_survey_export_set():
    syms = declared_symbols()
    for need in ('seg_conv2d_fwd', 'seg_conv2d_dgrad', 'seg_conv2d_wgrad', 'seg_deconv2d_fwd',
                 'seg_deconv2d_dgrad', 'seg_deconv2d_wgrad', 'seg_bilinear_upsample_fwd',
                 'seg_bilinear_upsample_bwd', 'seg_resize_bilinear_fwd', 'seg_resize_bilinear_bwd',
                 'seg_maxpool_fwd', 'seg_maxpool_bwd', 'seg_batchnorm_stats', 'seg_batchnorm_apply',
                 'seg_batchnorm_bwd_apply', 'seg_dropout', 'seg_softmax_xent_fwd_bwd',
                 'seg_sigmoid_argmax', 'seg_mc_mean_var', 'seg_adam_multi', 'seg_version',
                 'seg_device_check'):
        assert need in syms, need


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_header(lib):
    from segmentation_b200 import native
    bound = set(native.SIGNATURES) | {'seg_last_error_string', 'seg_last_kernel_name'}
    assert bound == set(declared_symbols())


def test_version_and_error_string_without_gpu(lib):
    lib.seg_version.restype = ctypes.c_int32
    assert lib.seg_version() >= 100
    lib.seg_last_error_string.restype = ctypes.c_char_p
    assert isinstance(lib.seg_last_error_string(), bytes)


def test_no_oracle_import_in_product():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, 'segmentation_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f


def test_sass_contains_tcgen05_and_tma():
    """Evidence that the hot kernels are Blackwell-native (B200_PROFILING.md):
    UTCHMMA (tcgen05.mma), UTMALDG (TMA, incl. im2col), LDTM (tcgen05.ld)."""
    import shutil
    import subprocess
    from segmentation_b200 import native
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not on PATH')
    sass = subprocess.run(['cuobjdump', '-sass', native.LIB_PATH], capture_output=True,
                          text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'IM2COL', 'LDTM'):
        assert mnemonic in sass, mnemonic


def test_option_keys_are_documented_and_accepted_without_gpu(lib):
    """Every OPT_* key of the ctypes binding is documented in the header ("key N") and
    accepted by seg_set_option (a host-side switch: no device needed); an unknown key is an
    error, not a silent no-op."""
    from segmentation_b200 import native
    keys = {k: v for k, v in vars(native).items() if k.startswith('OPT_') and isinstance(v, int)}
    assert len(set(keys.values())) == len(keys), keys            # no two names share a key
    doc = set(int(n) for n in re.findall(r'\bkeys?\s+(\d+)', open(HEADER).read()))
    doc |= set(int(n) for n in re.findall(r'key \d+ / key (\d+)', open(HEADER).read()))
    for name, key in keys.items():
        assert key in doc, (name, key)
    defaults = {native.OPT_HALO_CONV: 1, native.OPT_HALO_ROW_ALIGN: 0, native.OPT_TILE_CONV: 1,
                native.OPT_TILE_CONV_MIN_EFF: 65, native.OPT_TILE_WGRAD: 1,
                native.OPT_TILE_WGRAD_MIN_EFF: 40, native.OPT_PDL: 1,
                native.OPT_WGRAD_MIN_TILES: 8, native.OPT_POOL_ROWS: 1,
                native.OPT_DEEP_B_RING: 1, native.OPT_HALO_ROWSTAGE: 1,
                native.OPT_WGRAD_TENSOR_RED: 1, native.OPT_FIRST_LAYER: 1,
                native.OPT_SM_LIMIT: 0, native.OPT_TAIL_MMA: 1}
    assert set(defaults) == set(keys.values())
    lib.seg_set_option.restype = ctypes.c_int32
    lib.seg_set_option.argtypes = [ctypes.c_int32, ctypes.c_int32]
    for key, value in defaults.items():                          # re-apply the defaults
        assert lib.seg_set_option(key, value) == 0, key
    assert lib.seg_set_option(99, 1) != 0
    for gone in (8, 10, 13):                                     # round-1 options measured slower
        assert lib.seg_set_option(gone, 1) != 0


def test_probes_live_in_their_own_library(lib):
    """The product library exports no self-test / micro-benchmark hook: those are declared in
    include/segb200_probes.h and linked into libsegb200_probes.so only."""
    from segmentation_b200 import native
    src = open(os.path.join(ROOT, 'include', 'segb200_probes.h')).read()
    probes = sorted(set(re.findall(r'SEG_API\s+[\w\s\*]+?\b(seg_\w+)\s*\(', src)))
    assert probes and set(probes) == set(native.PROBE_SIGNATURES)
    plib = ctypes.CDLL(native.PROBES_LIB_PATH)
    for s in probes:
        assert hasattr(plib, s), s
        assert not hasattr(lib, s), s
    assert not hasattr(lib, 'seg_debug_prof_buffer')            # profiling builds only
