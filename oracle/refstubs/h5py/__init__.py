"""Import-time stand-in for h5py (absent from this image).

Test tooling only: lets the read-only reference at /root/reference import so it
can be run as the parity oracle.  Not product code; nothing here is on the hot path.
"""


class File:  # pragma: no cover - never instantiated by the golden scripts
    def __init__(self, *a, **k):
        raise RuntimeError("h5py stub: HDF5 I/O is not available in this image")


class Group:  # pragma: no cover
    pass


class Dataset:  # pragma: no cover
    pass
