"""Build libsdeb.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python sdepy_b200/_build.py [--force]      (or __graft_entry__.build())

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
SOURCES = ['sdeb.cu', 'sdeb_models_heston.cu', 'sdeb_models_linear.cu',
           'sdeb_models_meanrev.cu']
DEPS = SOURCES + ['sde_engine.cuh', 'sdeb_internal.h',
                  os.path.join('..', '..', 'include', 'sdeb.h')]
NVCC_FLAGS = ['-Xcompiler', '-fPIC', '-O3', '-std=c++17',
              '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo']


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
    nvcc = nvcc_path()
    extra = os.environ.get('SDEB_NVCC_FLAGS', '').split()
    objdir = os.path.join(CSRC, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:     # one nvcc per translation unit, in parallel
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc, '-c'] + NVCC_FLAGS + extra + (['-Xptxas=-v'] if verbose else []) + [
            '-o', obj, os.path.join(CSRC, src)]
        if verbose:
            print(' '.join(cmd))
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                 stderr=subprocess.STDOUT, text=True)))
    objs = []
    for cmd, obj, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), out))
        if verbose:
            print(out)
        objs.append(obj)
    link = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a',
            '-o', LIB] + objs + ['-ldl']
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
