"""Build libsdeb.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m sdepy_b200._build [--force]

nvcc cross-compiles without a GPU.  The built library lives next to its
sources (sdepy_b200/csrc/libsdeb.so): git-ignored, but shipped to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'libsdeb.so')
SOURCES = ['sdeb.cu']
DEPS = ['sdeb.cu', 'sde_engine.cuh', os.path.join('..', '..', 'include', 'sdeb.h')]


def nvcc_path():
    for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: cannot build libsdeb.so')


def stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > built for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [nvcc_path(), '-shared', '-Xcompiler', '-fPIC', '-O3', '-std=c++17',
           '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
           '-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ['-ldl']
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
