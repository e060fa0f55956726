"""Make the read-only reference importable (generation-time only).

The reference (/root/reference, HITEN v0.5.4) needs h5py and matplotlib at import;
neither is in this image and neither is on the propagation path, so tiny stand-ins
from oracle/refstubs are put ahead of it on sys.path.  The golden-vector generators use the
read-only checkout; on the GPU box, where /root/reference does not exist, the same unmodified package is
importable from oracle/_ref (copied there by oracle/build_ref.sh, git-ignored, shipped with the snapshot):
tests/test_gpu_dropin_real.py and bench.py's reference arm use that.
"""
import os
import sys

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF_SRC = os.environ.get("HITEN_REFERENCE_SRC", "/root/reference/src")
SHIPPED = os.path.join(REPO, "oracle", "_ref")       # oracle/build_ref.sh: unmodified copy that travels to the GPU box


def available():
    """Where the reference package can be imported from: the read-only checkout (build container) or the copy
    oracle/build_ref.sh made of it (GPU box); None if neither exists."""
    if os.path.isdir(os.path.join(REF_SRC, "hiten")):
        return REF_SRC
    if os.path.isdir(os.path.join(SHIPPED, "hiten")):
        return SHIPPED
    return None


def enable():
    src = available()
    if src is None:
        raise RuntimeError(f"reference sources not found at {REF_SRC} or {SHIPPED}")
    stubs = os.path.join(REPO, "oracle", "refstubs")
    for p in (src, stubs):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, src)
    sys.path.insert(0, stubs)
    os.environ.setdefault("HITEN_LOG_LEVEL", "WARNING")
    return src
