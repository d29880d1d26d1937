"""Builds the in-tree CUDA library (sm_100a) with nvcc. No torch headers: the boundary is a C ABI."""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvitpose_b200.so')
NVCC_FLAGS = ['-shared', '-Xcompiler', '-fPIC', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-O3', '-std=c++17', '-cudart', 'static']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; cannot build libvitpose_b200.so')
    return nvcc


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = glob.glob(os.path.join(CSRC, '*')) + [os.path.join(HERE, '..', 'include', 'vitpose_b200.h')]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    cmd = [_nvcc()] + NVCC_FLAGS + ['-o', LIB] + srcs
    if verbose:
        print(' '.join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
