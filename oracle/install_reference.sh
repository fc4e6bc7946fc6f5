#!/bin/bash
# TEST INFRASTRUCTURE (build container only): put an UNMODIFIED copy of the reference where the GPU box can
# see it -- baseline/_ref/ is git-ignored (never in history) but travels with gpurun:
#   baseline/_ref/pyslam/  the package, installed with pip from a scratch copy (the tree itself is read-only)
#   baseline/_ref/tests/   the reference's own test files, run UNMODIFIED against the product by
#                          tests/test_reference_suite.py (module aliases pyslam -> pyslam_b200, liegroups -> pyslam_b200.lie)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${PYSLAM_REFERENCE_ROOT:-/root/reference}
[ -f "$REF/pyslam/problem.py" ] || { echo "no reference tree at $REF"; exit 0; }
rm -rf /tmp/pyslam_ref_src && cp -r "$REF" /tmp/pyslam_ref_src
mkdir -p "$ROOT/baseline/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" --upgrade /tmp/pyslam_ref_src || echo "pip install of the reference failed (recorded in DESIGN.md)"
rm -rf "$ROOT/baseline/_ref/tests" && cp -r "$REF/tests" "$ROOT/baseline/_ref/tests"
ls "$ROOT/baseline/_ref"
