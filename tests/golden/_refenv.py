"""Make the read-only reference importable (generation-time only).

The reference (/root/reference, HITEN v0.5.4) needs h5py and matplotlib at import;
neither is in this image and neither is on the propagation path, so tiny stand-ins
from oracle/refstubs are put ahead of it on sys.path.  Only the golden-vector
generators import this module; tests, smoke() and bench.py never do
(/root/reference does not exist on the GPU box).
"""
import os
import sys

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF_SRC = os.environ.get("HITEN_REFERENCE_SRC", "/root/reference/src")


def enable():
    if not os.path.isdir(REF_SRC):
        raise RuntimeError(f"reference sources not found at {REF_SRC}")
    stubs = os.path.join(REPO, "oracle", "refstubs")
    for p in (REF_SRC, stubs):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_SRC)
    sys.path.insert(0, stubs)
    os.environ.setdefault("HITEN_LOG_LEVEL", "WARNING")
