"""Builds the in-tree CUDA library (sm_100a) with nvcc. No torch headers: the boundary is a C ABI.

Every ``csrc/*.cu`` is compiled to an object under ``vitpose_b200/build/`` (git-ignored) — in parallel, and only
when it or a header is newer than its object — then linked into ``libvitpose_b200.so``.
"""
import glob
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libvitpose_b200.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-Xcompiler', '-fPIC', '-lineinfo', '-O3', '-std=c++17'] + ARCH
LDFLAGS = ['-shared', '-Xcompiler', '-fPIC', '-cudart', 'static'] + ARCH


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; cannot build libvitpose_b200.so')
    return nvcc


def _headers():
    return (glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) +
            [os.path.join(HERE, '..', 'include', 'vitpose_b200.h'), os.path.abspath(__file__)])


def _stale(src, obj, hdr_time):
    return (not os.path.exists(obj)) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time)


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = glob.glob(os.path.join(CSRC, '*')) + [os.path.join(HERE, '..', 'include', 'vitpose_b200.h')]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False, extra_flags=()):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr_time = max(os.path.getmtime(h) for h in _headers())
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + '.o') for s in srcs]

    def compile_one(pair):
        src, obj = pair
        cmd = [nvcc, '-c'] + CFLAGS + list(extra_flags) + ['-o', obj, src]
        if verbose:
            print(' '.join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {os.path.basename(src)}:\n' + r.stdout + r.stderr)
        return r.stdout + r.stderr

    todo = [(s, o) for s, o in zip(srcs, objs) if force or _stale(s, o, hdr_time)]
    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4) or 1) as ex:
        logs = list(ex.map(compile_one, todo))
    if verbose:
        for (s, _), log in zip(todo, logs):
            if log.strip():
                print(os.path.basename(s) + ':\n' + log)
    r = subprocess.run([nvcc] + LDFLAGS + ['-o', LIB] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose=True,
                extra_flags=['-Xptxas', '-v'] if '--ptxas' in sys.argv else ()))
