"""hiten_b200 -- B200-native batched trajectory propagation behind HITEN's Python API.

Hand-written sm_100a CUDA kernels (fp64) reached through the C ABI in include/hiten_b200.h.
No CPU fallback: compute entry points raise without the built library and a CUDA device.
"""
from . import _lib
from ._lib import HitenB200Error
from .dropin import install, uninstall
from .propagate import (BatchResult, cr3bp_dense, cr3bp_event, cr3bp_propagate, cr3bp_stm, cr3bp_stm_dense,
                        cost_order, dfma_peak, make_integ, with_order)

__all__ = ["install", "uninstall", "BatchResult", "HitenB200Error", "cr3bp_dense", "cr3bp_event", "cr3bp_propagate", "cr3bp_stm", "cr3bp_stm_dense", "dfma_peak",
           "make_integ", "with_order", "cost_order", "_lib"]
