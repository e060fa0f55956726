#!/usr/bin/env bash
# TEST INFRASTRUCTURE (checker / CPU baseline only; the product never imports it).
#
# Recipe for oracle/_ref/: an unmodified copy of the reference's Python package (pure Python + Numba, nothing to
# compile) plus the h5py / matplotlib import stand-ins of oracle/refstubs, so that the REAL reference can run on the
# GPU box, where /root/reference does not exist.  oracle/_ref/ is git-ignored (no reference sources in history) but
# not gpurun-ignored, so it travels with the snapshot like the built .so files.
#
# Used by: tests/test_gpu_dropin_real.py (hiten_b200.install() + the reference's public API on the real CUDA
# library), bench.py --impl reference and bench.py's cpu_baseline leg (kind "reference").
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${HITEN_REFERENCE_SRC:-/root/reference/src}"
DST="$HERE/_ref"
if [ ! -d "$SRC/hiten" ]; then
    echo "build_ref.sh: reference sources not found at $SRC/hiten (nothing to do)" >&2
    exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/hiten" "$DST/hiten"
find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} +
cp -r "$HERE/refstubs/." "$DST/"
find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} +
( cd "$SRC/.." && git rev-parse HEAD 2>/dev/null || echo "unknown" ) > "$DST/REFERENCE_COMMIT"
echo "build_ref.sh: $(find "$DST/hiten" -name '*.py' | wc -l) reference modules -> $DST"
