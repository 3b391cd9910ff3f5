"""Build libsegb200.so (the C-ABI CUDA library) and libsegb200_probes.so (the same kernel
objects + the self-test / micro-benchmark hooks of include/segb200_probes.h) in-tree with nvcc
for sm_100a.

    python -m segmentation_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libsegb200.so')
PROBES_LIB = os.path.join(HERE, 'libsegb200_probes.so')
SOURCES = ['api.cu', 'simt_conv.cu', 'pointwise.cu', 'umma_conv.cu', 'fconv.cu']
PROBE_SOURCES = ['probe.cu', 'probe_api.cu']      # linked into libsegb200_probes.so only
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
              '--expt-relaxed-constexpr', '-Xptxas', '-v']
if os.environ.get('SEGB200_KERNEL_PROF') == '1':      # in-kernel timeline marks (tools/layer_prof.py)
    NVCC_FLAGS.append('-DSEGB200_KERNEL_PROF=1')


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join('..', '..', 'include', 'segb200.h'),
                                        os.path.join('..', '..', 'include', 'segb200_probes.h')]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, 'stamp')
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(PROBES_LIB) and \
            os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + '.log', 'w') as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, r.stderr[-8000:]))
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES) + len(PROBE_SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES + PROBE_SOURCES))
    n = len(SOURCES)
    for lib, members in ((LIB, objs[:n]), (PROBES_LIB, objs)):
        cmd = [nvcc, '-shared', '-o', lib] + members + ['-gencode', 'arch=compute_100a,code=sm_100a']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stderr[-4000:])
    with open(stamp, 'w') as f:
        f.write(dig)
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
