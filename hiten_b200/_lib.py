"""ctypes binding of libhiten_b200.so (C ABI declared in include/hiten_b200.h).

There is no CPU fallback: if the shared library is missing, or CUDA is unavailable when a compute
entry point is called, this module raises.  PyTorch is used only for device buffers and streams.
"""
import ctypes as C
import os

from . import _build

HB_OK = 0
HB_RK4, HB_RK6, HB_RK8, HB_RK45, HB_DOP853 = 4, 6, 8, 45, 853
HB_ARITH_PARITY, HB_ARITH_FAST = 0, 1
HB_RECORDS_ALL, HB_RECORDS_NEAR_SECTION = 0, 1
HB_TRAJ_OK, HB_TRAJ_HIT, HB_TRAJ_MAXSTEPS, HB_TRAJ_NONFINITE, HB_TRAJ_RECORD_OVERFLOW = 0, 1, 2, 3, 4

ERRORS = {-1: "HB_ERR_BADARG", -2: "HB_ERR_UNSUPPORTED", -3: "HB_ERR_NODEVICE"}


class HbCr3bp(C.Structure):
    _fields_ = [("mu", C.c_double), ("fwd", C.c_int32), ("flip_lo", C.c_int32),
                ("flip_hi", C.c_int32), ("_pad", C.c_int32)]


class HbInteg(C.Structure):
    _fields_ = [("method", C.c_int32), ("arith", C.c_int32), ("rtol", C.c_double), ("atol", C.c_double),
                ("max_step", C.c_double), ("min_step", C.c_double), ("max_attempts", C.c_int64),
                ("n_fixed_steps", C.c_int32), ("max_ctas", C.c_int32), ("order", C.c_void_p)]


class HbEvent(C.Structure):
    _fields_ = [("idx", C.c_int32), ("direction", C.c_int32), ("offset", C.c_double),
                ("xtol", C.c_double), ("gtol", C.c_double)]


class HbSection(C.Structure):
    _fields_ = [("idx", C.c_int32), ("direction", C.c_int32), ("offset", C.c_double),
                ("proj_i", C.c_int32), ("proj_j", C.c_int32), ("segment_refine", C.c_int32),
                ("max_hits_per_traj", C.c_int32), ("tol_on_surface", C.c_double),
                ("dedup_time_tol", C.c_double), ("dedup_point_tol", C.c_double)]


class HbPolyHam(C.Structure):
    _fields_ = [("n_dof", C.c_int32), ("max_deg", C.c_int32), ("ptr", C.c_int64 * 7), ("terms", C.c_void_p)]


HB_MAX_TAO_SUBSTEPS = 27
HB_SYMPLECTIC = 2


class HbCmOpts(C.Structure):
    _fields_ = [("dt", C.c_double), ("max_steps", C.c_int32), ("method", C.c_int32), ("order", C.c_int32),
                ("section", C.c_int32), ("arith", C.c_int32), ("n_sub", C.c_int32),
                ("sub_ts", C.c_double * HB_MAX_TAO_SUBSTEPS), ("sub_cos", C.c_double * HB_MAX_TAO_SUBSTEPS),
                ("sub_sin", C.c_double * HB_MAX_TAO_SUBSTEPS)]


class HbSympOpts(C.Structure):
    _fields_ = [("order", C.c_int32), ("arith", C.c_int32), ("m", C.c_int32), ("n_sub", C.c_int32)]


class HbCmLiftOpts(C.Structure):
    _fields_ = [("h0", C.c_double), ("initial_guess", C.c_double), ("expand_factor", C.c_double), ("xtol", C.c_double),
                ("max_expand", C.c_int32), ("symmetric", C.c_int32), ("section", C.c_int32), ("max_iter", C.c_int32)]


class HbTubeFilterOpts(C.Structure):
    _fields_ = [("mu", C.c_double), ("safe_r1", C.c_double), ("safe_r2", C.c_double), ("energy_tol", C.c_double)]


class HbCorrectOpts(C.Structure):
    _fields_ = [("ctrl", C.c_int32 * 2), ("res", C.c_int32 * 2), ("target", C.c_double * 2), ("event_idx", C.c_int32),
                ("halo_quadratic", C.c_int32), ("event_offset", C.c_double), ("finite_difference", C.c_int32),
                ("line_search", C.c_int32), ("tol", C.c_double), ("max_delta", C.c_double), ("fd_step", C.c_double),
                ("alpha_reduction", C.c_double), ("min_alpha", C.c_double), ("armijo_c", C.c_double),
                ("max_attempts", C.c_int32), ("_pad", C.c_int32)]


class HitenB200Error(RuntimeError):
    pass


_lib = None
vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/hiten_b200.h
SIGNATURES = {
    "hb_workspace_bytes": (C.c_int64, []),
    "hb_device_info": (C.c_int, [C.POINTER(C.c_int32)] * 3),
    "hb_cr3bp_propagate": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.c_int64, vp, C.c_double, C.c_double,
                                     vp, C.c_int32, vp, vp, vp, vp, vp, vp]),
    "hb_cr3bp_dense": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.c_int64, vp, vp, C.c_int32, vp, vp, vp,
                                 vp, vp, vp]),
    "hb_cr3bp_section": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.POINTER(HbSection), C.c_int64, vp, vp,
                                   C.c_int32, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp]),
    "hb_section2_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "hb_cr3bp_section2": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.POINTER(HbSection), C.c_int64, vp, vp,
                                    C.c_int32, vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int64, vp, vp, vp, C.c_int32]),
    "hb_section3_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "hb_cr3bp_section3": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.POINTER(HbSection), C.c_int64, vp, vp,
                                    C.c_int32, vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int64, vp, vp, vp]),
    "hb_cr3bp_event": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.POINTER(HbEvent), C.c_int64, vp,
                                 C.c_double, C.c_double, vp, vp, vp, vp, vp, vp, vp, vp]),
    "hb_dfma_peak": (C.c_int, [C.c_double, C.POINTER(C.c_double), vp]),
    "hb_cr3bp_stm": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.c_int64, vp, C.c_double, C.c_double, vp, vp,
                               vp, vp, vp, vp, vp]),
    "hb_cr3bp_stm_dense": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.c_int64, vp, vp, C.c_int32, C.c_int32,
                                     vp, vp, vp, vp, vp, vp]),
    "hb_cm_prepare": (C.c_int, [C.POINTER(HbCmOpts), C.c_double]),
    "hb_connections_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "hb_connections": (C.c_int, [vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_double, C.c_double, C.c_double, vp, C.c_int64,
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp, C.c_int64, vp]),
    "hb_cm_lift": (C.c_int, [C.POINTER(HbPolyHam), C.POINTER(HbCmLiftOpts), C.c_int64, vp, vp, vp, vp]),
    "hb_tao_grid_prepare": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_double, C.POINTER(C.c_int32), vp, C.c_int64]),
    "hb_ham_symplectic_dense": (C.c_int, [C.POINTER(HbPolyHam), C.POINTER(HbSympOpts), C.c_int64, vp, vp, vp, vp, vp]),
    "hb_ham_symplectic_event": (C.c_int, [C.POINTER(HbPolyHam), C.POINTER(HbSympOpts), C.POINTER(HbEvent), C.c_int64, vp,
                                          vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "hb_ham_symplectic_jit": (C.c_int, [C.POINTER(HbPolyHam), C.POINTER(HbSympOpts), C.POINTER(HbEvent), C.c_int64, vp,
                                        vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "hb_symp_jit_compile_host": (C.c_int, [vp, C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int64)]),
    "hb_ham_rk_dense": (C.c_int, [C.POINTER(HbPolyHam), C.c_int32, C.c_int32, C.c_int64, vp, vp, C.c_int32, vp, vp, vp, vp]),
    "hb_ham_rk_event": (C.c_int, [C.POINTER(HbPolyHam), C.c_int32, C.c_int32, C.POINTER(HbEvent), C.c_int64, vp, vp,
                                  C.c_int32, vp, vp, vp, vp, vp, vp, vp]),
    "hb_ham_adaptive_dense": (C.c_int, [C.POINTER(HbPolyHam), C.POINTER(HbInteg), C.c_int64, vp, vp, C.c_int32, vp, vp, vp,
                                        vp, vp, vp, vp]),
    "hb_ham_adaptive_event": (C.c_int, [C.POINTER(HbPolyHam), C.POINTER(HbInteg), C.POINTER(HbEvent), C.c_int64, vp,
                                        C.c_double, C.c_double, vp, vp, vp, vp, vp, vp, vp]),
    "hb_cm_poincare_map": (C.c_int, [C.POINTER(HbPolyHam), C.POINTER(HbCmOpts), C.c_int64, vp, vp, vp, vp, vp, vp]),
    "hb_cm_poincare_map_jit": (C.c_int, [C.POINTER(HbPolyHam), C.POINTER(HbCmOpts), C.c_int64, vp, vp, vp, vp, vp, vp]),
    "hb_cm_jit_compile_host": (C.c_int, [vp, C.POINTER(C.c_int64), C.c_int32, C.POINTER(HbCmOpts), C.POINTER(C.c_int64),
                                         C.c_char_p, C.c_int64]),
    "hb_manifold_ics": (C.c_int, [vp, vp, C.c_int32, C.c_double, vp, C.c_int32, vp, C.c_int64, vp, C.c_int64, vp, vp,
                                  vp]),
    "hb_tube_filter": (C.c_int, [C.POINTER(HbTubeFilterOpts), C.c_int64, vp, C.c_int32, vp, vp, vp]),
    "hb_correct_scratch_bytes": (C.c_int64, [C.c_int64]),
    "hb_correct_orbits": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.POINTER(HbCorrectOpts), C.c_int64, vp, vp,
                                    vp, vp, vp, vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp, C.c_int64, vp, vp]),
    "hb_section2_filter": (C.c_int, [C.POINTER(HbCr3bp), C.POINTER(HbInteg), C.POINTER(HbTubeFilterOpts), C.c_int64, vp,
                                     C.c_int32, vp, vp, vp, C.c_int64, vp, vp, vp]),
    "hb_synodic_detect": (C.c_int, [C.POINTER(HbSection), C.c_int64, vp, vp, vp, C.c_int32, C.c_int32, vp, C.c_int64,
                                    vp, vp, vp]),
    "hb_synodic_detect_cubic": (C.c_int, [C.POINTER(HbSection), C.c_int32, C.c_int64, vp, vp, vp, C.c_int32, C.c_int32, vp,
                                          C.c_int64, vp, vp, vp]),
    "hb_read_hit_count": (C.c_int, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp]),
    "hb_read_record_overflow": (C.c_int, [vp, C.POINTER(C.c_int64), vp]),
    "hb_peer_put": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, vp, C.c_int64, vp, C.c_int64, vp, vp]),
    "hb_selftest_arith": (C.c_int, [vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp]),
}


def lib_path():
    return _build.LIB_PATH


def load():
    """Load the shared library (building it first if sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("HITEN_B200_LIB")                    # override: kernel-variant experiments
    if path is None:
        path = _build.LIB_PATH
        missing = not os.path.exists(path)
        if missing or (_build.needs_build() and _build.have_nvcc()):
            try:
                _build.build()
            except Exception as exc:  # pragma: no cover
                if missing:
                    raise HitenB200Error(
                        f"libhiten_b200.so is missing and could not be built ({exc}); "
                        "run `python -c 'import __graft_entry__ as g; g.build()'`") from exc
                raise HitenB200Error(f"libhiten_b200.so is older than its sources and the rebuild failed: {exc}") from exc
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc == HB_OK:
        return
    if rc < 0:
        raise HitenB200Error(f"{what}: {ERRORS.get(rc, rc)}")
    raise HitenB200Error(f"{what}: CUDA error {rc}")
