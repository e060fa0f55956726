"""CPU: the C-ABI shared library loads and exports every symbol include/hiten_b200.h declares."""
import ctypes
import os
import re

import pytest

import hiten_b200
from hiten_b200 import _build, _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "hiten_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    _build.build()
    lib = ctypes.CDLL(_build.LIB_PATH)
    names = _declared_symbols()
    assert "hb_cr3bp_propagate" in names and "hb_workspace_bytes" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hiten_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes SIGNATURES out of sync with the header"


def test_workspace_size_needs_no_gpu():
    assert _lib.load().hb_workspace_bytes() >= 256


def test_compute_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(hiten_b200.HitenB200Error):
        hiten_b200.cr3bp_propagate([[0.8, 0, 0, 0, 0.1, 0]], 0.0121, 1.0)
