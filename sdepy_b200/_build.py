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
N_MODEL_UNITS = 11          # = SDEB_N_UNITS in sdeb_internal.h
# (source, object stem, extra flags): sdeb_models.cu is compiled once per unit
UNITS = [('sdeb.cu', 'sdeb', [])] + [
    ('sdeb_models.cu', 'sdeb_models_%d' % k, ['-DSDEB_UNIT=%d' % k])
    for k in range(N_MODEL_UNITS)]
SOURCES = ['sdeb.cu', 'sdeb_models.cu']
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
    jobs = []
    for src, stem, flags in UNITS:
        obj = os.path.join(objdir, stem + '.o')
        cmd = [nvcc, '-c'] + NVCC_FLAGS + extra + flags + (
            ['-Xptxas=-v'] if verbose else []) + ['-o', obj, os.path.join(CSRC, src)]
        jobs.append((cmd, obj))
    # one nvcc per unit, as many at a time as there are cores
    width = max(1, os.cpu_count() or 1)
    objs, running = [], []
    pending = list(jobs)
    while pending or running:
        while pending and len(running) < width:
            cmd, obj = pending.pop(0)
            if verbose:
                print(' '.join(cmd))
            running.append((cmd, obj, subprocess.Popen(
                cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        cmd, obj, proc = running.pop(0)
        out, _ = proc.communicate()
        if proc.returncode != 0:
            for _, _, other in running:
                other.kill()
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
