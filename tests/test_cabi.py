"""CPU: the C-ABI shared library loads and exports every symbol include/hiten_b200.h declares."""
import ctypes
import os
import re

import numpy as np
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


def test_specialised_cm_kernel_compiles_offline():
    """The run-time specialised centre-manifold kernel (hb_cm_jit.cu) is generated and compiled for sm_100a with
    NVRTC here, without a GPU: one straight-line MADD per non-zero gradient term, both arithmetic variants."""
    import numpy as np
    from hiten_b200 import centermanifold as cm
    g = np.load(os.path.join(REPO, "tests", "golden", "cm_map.npz"))
    tab = cm.PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
    for method, order in (("symplectic", 4), ("fixed", 4)):
        for arith in ("parity", "fast"):
            nbytes, src = cm.jit_compile_host(tab, cm.make_opts(0.01, 2000, method, order, "p3", 20.0, arith), True)
            assert nbytes > 10_000
            # 119 terms in grad() and the same 119 dealt out over grad_part0..3 (cm_map_split: four warps per 32 seeds)
            assert src.count("acc = MADD(") == 2 * 119 and "extern \"C\" __global__" in src
            parts = [src.split(f"DEV void grad_part{w}(")[1].split("DEV void grad_part" if w < 3 else "const char")[0]
                     for w in range(4)]
            counts = [p_.split("\n}\n")[0].count("acc = MADD(") for p_ in parts]
            assert sum(counts) == 119 and max(counts) <= 40, counts


def test_tao_schedule_matches_reference_recursion():
    """hb_cm_prepare flattens _recursive_update_poly (symplectic.py:543-560): 3^(order/2 - 1) order-2 kernels whose
    time steps sum to dt, with omega = (c*dt)^-order."""
    from hiten_b200 import centermanifold as cm
    for order, n_sub in ((2, 1), (4, 3), (6, 9), (8, 27)):
        o = cm.make_opts(0.01, 10, "symplectic", order, "q3", 20.0)
        assert o.n_sub == n_sub
        ts = [o.sub_ts[i] for i in range(n_sub)]
        assert abs(sum(ts) - 0.01) < 1e-15
        omega = (20.0 * 0.01) ** (-float(order))
        import math
        assert o.sub_cos[0] == math.cos(2 * omega * ts[0]) and o.sub_sin[0] == math.sin(2 * omega * ts[0])


def test_argument_validation_needs_no_gpu():
    """Usage errors are rejected with HB_ERR_BADARG (-1) / HB_ERR_UNSUPPORTED before anything touches the device, and
    the size helpers are pure host arithmetic."""
    lib = _lib.load()
    C = ctypes
    i64 = C.c_int64
    # scratch sizes: monotone, 512 B per step per trajectory plus the fixed part
    a, b = lib.hb_section2_scratch_bytes(1000, 64), lib.hb_section2_scratch_bytes(1000, 128)
    assert b - a == 1000 * 64 * 512 and a > 1000 * 64 * 512
    assert lib.hb_section2_scratch_bytes(10, 0) < 0 and lib.hb_section2_scratch_bytes(-1, 64) < 0
    assert lib.hb_connections_scratch_bytes(1000, 2000) > 0 and lib.hb_connections_scratch_bytes(-1, 5) < 0
    # null / inconsistent arguments
    n_out = i64(7)
    assert lib.hb_connections(None, 5, None, 5, None, None, 1e-3, 1.0, 1e-3, None, 0, C.byref(n_out), None, None, None, 0,
                              None) < 0
    assert lib.hb_connections(None, 0, None, 5, None, None, 1e-3, 1.0, 1e-3, None, 0, C.byref(n_out), None, None, None, 0,
                              None) == 0 and n_out.value == 0                     # an empty set: no connections
    assert lib.hb_connections(None, 5, None, 5, None, None, -1.0, 1.0, 1e-3, None, 0, C.byref(n_out), None, None, None, 0,
                              None) < 0                                           # eps must be positive
    ham = _lib.HbPolyHam(3, 6, (C.c_int64 * 7)(0, 0, 0, 0, 0, 0, 0), None)
    o = _lib.HbCmLiftOpts(0.7, 1e-3, 2.0, 1e-12, 40, 0, 7, 200)                   # section 7 does not exist
    assert lib.hb_cm_lift(C.byref(ham), C.byref(o), 4, None, None, None, None) < 0
    o.section = 3
    assert lib.hb_cm_lift(C.byref(ham), C.byref(o), 0, None, None, None, None) == 0   # n = 0 is a no-op
    assert lib.hb_cm_lift(C.byref(ham), C.byref(o), 4, None, None, None, None) < 0    # null arrays
    sysd, integ, sec = _lib.HbCr3bp(0.0121, 1, -1, -1, 0), hiten_b200.make_integ(), hiten_b200.synodic.make_section("y")
    assert lib.hb_cr3bp_section2(C.byref(sysd), C.byref(integ), C.byref(sec), 8, None, None, 1, None, 0, None, None, None,
                                 None, None, None, 0, None, None, None, 0) < 0    # m < 2, no workspace
    assert lib.hb_cr3bp_section3(C.byref(sysd), C.byref(integ), C.byref(sec), 8, None, None, 1, None, 0, None, None, None,
                                 None, None, None, 0, None, None, None) < 0
    # section3 scratch: 4.2 KB of lists per trajectory + 512 B per pooled record, monotone in both arguments
    a3, b3 = lib.hb_section3_scratch_bytes(1000, 4), lib.hb_section3_scratch_bytes(1000, 8)
    assert b3 - a3 == 1000 * 4 * 512 and lib.hb_section3_scratch_bytes(2000, 4) > a3
    assert lib.hb_section3_scratch_bytes(-1, 4) < 0 and lib.hb_section3_scratch_bytes(10, 0) < 0
    assert lib.hb_read_record_overflow(None, None, None) < 0
    # manifold-tube entry points (8f#3)
    assert lib.hb_manifold_ics(None, None, 2000, 2.75, None, 1, None, 0, None, 5, None, None, None) == 0   # empty tube
    assert lib.hb_manifold_ics(None, None, 2000, 2.75, None, 1, None, 4, None, 5, None, None, None) < 0    # null arrays
    assert lib.hb_manifold_ics(None, None, 2000, 2.75, None, 0, None, 0, None, 0, None, None, None) < 0    # direction 0
    fo = _lib.HbTubeFilterOpts(0.0121, 0.0, 0.0, 1e-6)
    assert lib.hb_tube_filter(C.byref(fo), 0, None, 10, None, None, None) == 0
    assert lib.hb_tube_filter(C.byref(fo), 3, None, 10, None, None, None) < 0
    assert lib.hb_tube_filter(C.byref(fo), 3, None, 0, None, None, None) < 0
    # batched corrector (8f#4)
    from hiten_b200 import corrector
    assert lib.hb_correct_scratch_bytes(1000) > 1000 * 8 * 100 and lib.hb_correct_scratch_bytes(-1) < 0
    co = corrector.make_opts("halo")
    s6 = i64(5)
    assert lib.hb_correct_orbits(C.byref(sysd), C.byref(integ), C.byref(co), 0, None, None, None, None, None, None,
                                 C.byref(s6), None, None, 0, (C.c_char * 256)(), None) == 0 and s6.value == 0
    assert lib.hb_correct_orbits(C.byref(sysd), C.byref(integ), C.byref(co), 4, None, None, None, None, None, None,
                                 None, None, None, 0, (C.c_char * 256)(), None) < 0            # null arrays
    back = _lib.HbCr3bp(0.0121, -1, -1, -1, 0)
    assert lib.hb_correct_orbits(C.byref(back), C.byref(integ), C.byref(co), 0, None, None, None, None, None, None,
                                 None, None, None, 0, (C.c_char * 256)(), None) == -2          # backward: unsupported
    bad = corrector.make_opts("halo", control_indices=(0, 0))
    assert lib.hb_correct_orbits(C.byref(sysd), C.byref(integ), C.byref(bad), 0, None, None, None, None, None, None,
                                 None, None, None, 0, (C.c_char * 256)(), None) == -1
    with pytest.raises(ValueError):
        corrector.make_opts(control_indices=(0, 4, 5), residual_indices=(3, 5), event_idx=1)
    # peer put (8e): world size, alignment and parity of the slot are checked before any launch
    slots = (C.c_void_p * 2)(0x1000, 0x2000)
    assert lib.hb_peer_put(None, 2, 0, None, 0, None, 0, 0x1000, None) < 0                       # no slot table
    assert lib.hb_peer_put(slots, 17, 0, None, 0, None, 0, 0x1000, None) < 0                     # more than 16 ranks
    assert lib.hb_peer_put(slots, 2, 2, None, 0, None, 0, 0x1000, None) < 0                      # destination outside the group
    assert lib.hb_peer_put(slots, 2, 0, 0x3000, 3, None, 0, 0x1000, None) < 0                    # odd number of record slots
    assert lib.hb_peer_put((C.c_void_p * 2)(0x1008, 0x2000), 2, 0, None, 0, None, 0, 0x1000, None) < 0   # slot not 16-byte aligned
    assert lib.hb_peer_put(slots, 2, 0, None, 0, None, 0, None, None) < 0                        # no workspace
    assert hiten_b200.make_integ(max_ctas=74).max_ctas == 74 and hiten_b200.make_integ().max_ctas == 0
    # hb_integ.order (scheduling hint): null by default; with_order copies the settings and checks the tensor
    base = hiten_b200.make_integ(rtol=1e-10, max_ctas=3)
    assert base.order is None and C.sizeof(base) == 64
    same = hiten_b200.with_order(base, None)
    assert same is not base and same.order is None and same.rtol == 1e-10 and same.max_ctas == 3
    with pytest.raises(ValueError):
        hiten_b200.with_order(base, [2, 0, 1])                                                  # not a CUDA int32 tensor
    # the launch order is checked on the host side of the boundary (the kernels trust it): a permutation, nothing else
    import torch
    from hiten_b200 import propagate as P_
    cost = torch.tensor([5, 9, 5, 1, 9], dtype=torch.int32)
    order = hiten_b200.cost_order(cost)
    assert order.dtype == torch.int32 and order.tolist() == [1, 4, 0, 2, 3]                    # most expensive first, stable
    P_.check_order(order, 5)
    for bad_order, n_ in ((torch.tensor([0, 1, 1], dtype=torch.int32), 3), (torch.tensor([0, 1, 3], dtype=torch.int32), 3),
                          (torch.tensor([0, -1, 2], dtype=torch.int32), 3), (order, 4)):
        with pytest.raises(ValueError):
            P_.check_order(bad_order, n_)


def test_tao_grid_table_is_host_only_and_matches_the_reference_formulas():
    """hb_tao_grid_prepare (no GPU): per-interval omega = (c*dt)^-order (symplectic.py:38-60), the triple-jump sub-steps of
    _recursive_update_poly (:543-560) and cos / sin(2*omega*ts), for ascending and descending (fwd = -1) grids."""
    import math
    from hiten_b200 import symplectic as S
    t = np.linspace(0.5, 2.5, 41)
    for order, n_expect in ((2, 1), (4, 3), (6, 9), (8, 27)):
        for sign in (1.0, -1.0):
            n_sub, tab = S.tao_grid_table(t * sign, order, 20.0)
            assert n_sub == n_expect and tab.shape == (40, 3, n_expect)
            dts = np.diff(t * sign)

            def sched(ts, k):
                if k == 2:
                    return [ts]
                gam = 1.0 / (2.0 - 2.0 ** (1.0 / (float(k) + 1.0)))
                return sched(gam * ts, k - 2) + sched((1.0 - 2.0 * gam) * ts, k - 2) + sched(gam * ts, k - 2)
            for i in (0, 17, 39):
                ts = sched(float(dts[i]), order)
                omega = (20.0 * float(dts[i])) ** (-float(order))
                assert tab[i, 0].tolist() == ts
                assert tab[i, 1].tolist() == [math.cos(2 * omega * x) for x in ts]
                assert tab[i, 2].tolist() == [math.sin(2 * omega * x) for x in ts]
    with pytest.raises(ValueError):
        S.tao_grid_table([0.0], 4)
    with pytest.raises(ValueError):
        S.tao_grid_table([0.0, 1.0], 3)
    lib, C = _lib.load(), ctypes
    ham = _lib.HbPolyHam(3, 6, (C.c_int64 * 7)(0, 0, 0, 0, 0, 0, 0), None)
    o = _lib.HbSympOpts(4, 0, 10, 0)                                     # n_sub = 0: prepare was not called
    assert lib.hb_ham_symplectic_dense(C.byref(ham), C.byref(o), 4, None, None, None, None, None) < 0
    o.n_sub = 3
    assert lib.hb_ham_symplectic_dense(C.byref(ham), C.byref(o), 4, None, None, None, None, None) < 0   # no workspace
    ev = _lib.HbEvent(9, 0, 0.0, 1e-12, 1e-12)                           # component 9 does not exist
    assert lib.hb_ham_symplectic_event(C.byref(ham), C.byref(o), C.byref(ev), 0, None, None, None, None, None, None, None,
                                       None, None, None) < 0


def test_specialised_grid_kernel_compiles_offline():
    """hb_symp_jit_compile_host: the run-time specialised Tao grid kernel (generated gradient + grid / event loop) compiles
    for sm_100a with NVRTC without a GPU, in both arithmetic variants, for the CM table and for the pendulum."""
    from hiten_b200 import symplectic as S
    from hiten_b200.centermanifold import PolyTable
    for f in ("cm_map.npz", "pendulum.npz"):
        g = np.load(os.path.join(REPO, "tests", "golden", f))
        tab = PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
        for arith in ("parity", "fast"):
            assert S.jit_compile_host(tab, arith) > 4096


def test_tube_section_chunk_plan_is_host_arithmetic():
    """synodic._auto_plan (tube_section's "auto" form): whole batch when its step scratch fits half of the free memory,
    equal 256-aligned chunks that do when it does not, the fused kernel for tiny batches or no room at all."""
    from hiten_b200 import synodic
    per = _lib.load().hb_section2_scratch_bytes(1024, 128) / 1024.0 + 400.0
    assert synodic._auto_plan(100, "cuda:0", "near", free_bytes=1e12) == (100, 0)                  # tiny: fused kernel
    assert synodic._auto_plan(1_000_000, "cuda:0", "near", free_bytes=180e9) == (1_000_000, 128)   # fits a B200
    assert synodic._auto_plan(1_000_000, "cuda:0", "all", free_bytes=180e9)[1] == 160
    chunk, cap = synodic._auto_plan(10_000_000, "cuda:0", "near", free_bytes=170e9)
    assert cap == 128 and chunk % 256 == 0 and chunk * per <= 0.5 * 170e9 + 256 * per
    n_chunks = -(-10_000_000 // chunk)
    assert 8 <= n_chunks <= 10 and (n_chunks - 1) * chunk < 10_000_000
    assert synodic._auto_plan(50_000, "cuda:0", "near", free_bytes=10e6) == (50_000, 0)            # no room: fused kernel
