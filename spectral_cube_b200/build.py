"""
Build libsc_b200.so (hand-written CUDA for sm_100a + the C ABI of include/sc_b200.h) in-tree.

    python -m spectral_cube_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to ``csrc/build/``; the shared library is
``csrc/libsc_b200.so`` (git-ignored, shipped to the GPU box with the snapshot).
"""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, 'libsc_b200.so')
OBJDIR = os.path.join(CSRC, 'build')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_FLAGS = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
              '-Xcompiler', '-O3', '-Xptxas', '-v', '--expt-relaxed-constexpr']


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsc_b200.so cannot be built (there is no CPU fallback)")


def _deps():
    return glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(ROOT, 'include', '*.h'))


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(nvcc, src, obj, verbose):
    cmd = [nvcc] + ARCH + NVCC_FLAGS + ['-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s" % src)
    # keep the ptxas resource report beside the object (registers / spills / smem)
    with open(obj + '.ptxas.txt', 'w') as f:
        f.write(r.stderr)
    return obj


def build_library(force=False, verbose=False):
    nvcc = find_nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    deps = _deps()
    jobs, objs = [], []
    for src in sources:
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or _stale(obj, [src] + deps):
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs))
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ['-shared', '-o', LIB] + objs + ['-lcudart_static', '-lpthread', '-ldl', '-lrt']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("link of libsc_b200.so failed")
    return LIB


if __name__ == '__main__':
    path = build_library(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
