#!/usr/bin/env python
"""Build libskge_b200.so in-tree with nvcc for sm_100a (no CMake, no JIT cache).

    python scikit-kge_b200/build.py [--force] [--verbose]

The .so lands in scikit-kge_b200/lib/ (git-ignored, but shipped to the GPU box
with the repo snapshot).  Objects are rebuilt only when a source or header is
newer than the object.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
INC = os.path.join(ROOT, 'include')
OBJ = os.path.join(HERE, 'build')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libskge_b200.so')

NVCC = os.environ.get('NVCC') or shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr',
         '-I', INC, '-I', CSRC] + os.environ.get('SKGE_NVCC_EXTRA', '').split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers_mtime():
    ts = [os.path.getmtime(os.path.join(INC, f)) for f in os.listdir(INC)]
    ts += [os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    ts.append(os.path.getmtime(os.path.abspath(__file__)))
    return max(ts)


def compile_one(src, force, verbose):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    spath = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(spath), headers_mtime())):
        return obj, ''
    cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', spath, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: compile_one(s, force, verbose), srcs))
    objs = [o for o, _ in res]
    if verbose:
        for _, log in res:
            if log:
                print(log)
    if (force or not os.path.exists(LIB)
            or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs)):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                      '-Xcompiler', '-fPIC', '-cudart', 'static']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
