#!/usr/bin/env python
"""Link a VARIANT of libsdeb.so for timing experiments: unit 0 of sdeb_models.cu
(the Heston kernels) is recompiled with extra nvcc flags and linked with the
objects of the regular build.

    python tools/build_variant.py <name> [nvcc flags...]
      -> gpurun_variants/libsdeb_<name>.so      (git-ignored, shipped by gpurun)
    SDEB_LIB=gpurun_variants/libsdeb_<name>.so python bench.py --no-cpu-baseline

Variants are measurement tools only (e.g. -DSDEB_PHILOX_ROUNDS=7 changes the
stream); the product is the regular build.
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'sdepy_b200'))
import _build  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    unit = 0
    for f in list(flags):            # --unit=K: recompile unit K of sdeb_models.cu instead of unit 0
        if f.startswith('--unit='):
            unit = int(f.split('=')[1])
            flags.remove(f)
    if '--no-rebuild' in flags:      # link against the objects as they are (stale or not)
        flags.remove('--no-rebuild')
    else:
        _build.build()
    out = os.path.join(ROOT, 'gpurun_variants')
    os.makedirs(out, exist_ok=True)
    obj = os.path.join(out, 'unit%d_%s.o' % (unit, name))
    nvcc = _build.nvcc_path()
    subprocess.check_call([nvcc, '-c'] + _build.NVCC_FLAGS + ['-DSDEB_UNIT=%d' % unit] + flags +
                          ['-o', obj, os.path.join(_build.CSRC, 'sdeb_models.cu')])
    objs = [o for o in glob.glob(os.path.join(_build.CSRC, 'build', '*.o'))
            if not o.endswith('sdeb_models_%d.o' % unit)] + [obj]
    lib = os.path.join(out, 'libsdeb_%s.so' % name)
    subprocess.check_call([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a',
                           '-o', lib] + objs + ['-ldl'])
    print(lib)


if __name__ == '__main__':
    main()
