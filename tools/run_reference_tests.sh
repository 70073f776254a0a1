#!/bin/bash
# The reference's own test files against this package (the drop-in check of
# profiles/rNN_reference_tests.md).  Run in the authoring container:
#
#     bash tools/run_reference_tests.sh prepare     # scratch copy under _scratch_ref/ (git-ignored)
#     gpurun --timeout 1500 -- 'bash tools/run_reference_tests.sh run'
#     bash tools/run_reference_tests.sh clean
#
# The scratch copy holds the UNMODIFIED reference package and its tests with ONE edit
# to tests/shared.py (`import sdepy` -> a module exposing every name of sdepy_b200 and,
# for names this package does not define, the reference's) and the SciPy spelling
# fix `scipy.stats.trapz` -> `trapezoid` (the reference itself fails on those lines
# under SciPy 1.18).  Nothing of it is ever committed.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
S=$ROOT/_scratch_ref
case "$1" in
prepare)
    rm -rf "$S"; mkdir -p "$S/t"
    cp -r /root/reference/sdepy "$S/sdepy"
    cp /root/reference/sdepy/tests/*.py "$S/t/"
    cp -r /root/reference/sdepy/tests/cfr "$S/t/cfr" 2>/dev/null || true
    printf '[pytest]\nmarkers =\n    slow\n    quant\n' > "$S/pytest.ini"
    cat > "$S/sdepy_shim.py" <<'PY'
import os, sys, types
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)                      # the reference copy: `import sdepy`
sys.path.insert(0, os.path.dirname(HERE))     # this repository: `import sdepy_b200`
import sdepy as _ref
import sdepy_b200 as _new
mod = types.ModuleType('sdepy_under_test')
for k in dir(_ref):
    if not k.startswith('__'):
        setattr(mod, k, getattr(_ref, k))
for k in dir(_new):
    if not k.startswith('__'):
        setattr(mod, k, getattr(_new, k))
mod.__version__ = _ref.__version__
mod.__file__ = _ref.__file__
PY
    python - "$S" <<'PY'
import re, sys, glob, os
S = sys.argv[1]
p = os.path.join(S, 't', 'shared.py')
s = open(p).read()
s = s.replace('import sdepy\nimport sdepy as sp', 'from sdepy_shim import mod as sdepy\nsp = sdepy')
open(p, 'w').write(s)
for f in glob.glob(os.path.join(S, 't', 'test_*.py')):
    s = open(f).read()
    s2 = s.replace('scipy.stats.trapz(', 'scipy.stats.trapezoid(').replace(
        'scipy.integrate.trapz(', 'scipy.integrate.trapezoid(')
    if s2 != s:
        open(f, 'w').write(s2)
PY
    echo prepared "$S"
    ;;
run)
    cd "$S"
    mkdir -p "$ROOT/gpurun_out"
    PYTHONPATH="$S:$ROOT" python -m pytest t -q -m "not slow and not quant" -p no:cacheprovider \
        --deselect t/test_process.py::test_no_override --deselect t/test_process.py::test_import \
        > "$ROOT/gpurun_out/reference_tests.log" 2>&1 || true
    tail -40 "$ROOT/gpurun_out/reference_tests.log"
    ;;
clean)
    rm -rf "$S"
    ;;
esac
