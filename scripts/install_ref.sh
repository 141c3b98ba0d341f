#!/bin/sh
# Installs the UNMODIFIED reference (quantscious/finmlkit) into baseline/_ref -- git-ignored, NOT gpurun-ignored, so it
# travels to the GPU box where /root/reference does not exist.  bench.py --impl reference and the cpu_baseline legs import
# it from there (the Numba path is the timed baseline; the C port under oracle/ is only the fallback when this is absent).
# Also copies the reference's own test files next to it (baseline/_ref/ref_tests): tests/test_gpu_ref_suite.py runs them
# against finmlkit_b200 with `finmlkit` aliased to the mirror.  Nothing from the reference enters the git history.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
if [ ! -d "$REF/finmlkit" ]; then
    echo "install_ref: no reference tree at $REF (nothing to do)"; exit 0
fi
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"                      # /root/reference is read-only and the build writes egg-info into the tree
python -m pip install --quiet --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --upgrade \
    --target "$ROOT/baseline/_ref" "$TMP/ref"
rm -rf "$ROOT/baseline/_ref/ref_tests"
mkdir -p "$ROOT/baseline/_ref/ref_tests"
cp -r "$REF/tests/." "$ROOT/baseline/_ref/ref_tests/"
rm -rf "$TMP"
echo "install_ref: finmlkit $(ls "$ROOT/baseline/_ref" | grep dist-info) -> $ROOT/baseline/_ref"
