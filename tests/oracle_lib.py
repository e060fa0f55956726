"""ctypes binding of the CPU parity oracle (oracle/libhiten_oracle.so).

Test infrastructure: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
ORACLE_DIR = os.path.join(REPO, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libhiten_oracle.so")

SYS_CR3BP6, SYS_VAR42, SYS_POLYHAM = 0, 1, 2
RK4, RK6, RK8, RK45, DOP853 = 4, 6, 8, 45, 853


class HoSystem(C.Structure):
    _fields_ = [("kind", C.c_int), ("dim", C.c_int), ("mu", C.c_double), ("fwd", C.c_int),
                ("flip_lo", C.c_int), ("flip_hi", C.c_int), ("ham", C.c_void_p)]


class HoEvent(C.Structure):
    _fields_ = [("idx", C.c_int), ("offset", C.c_double), ("direction", C.c_int),
                ("xtol", C.c_double), ("gtol", C.c_double)]


class HoTol(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("max_step", C.c_double), ("min_step", C.c_double)]


_lib = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int64)


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", ORACLE_DIR, "CC=gcc"], check=True, capture_output=True, env=env)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.ho_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(dp)


def default_tol(rtol=1e-12, atol=1e-12, max_step=1e4, min_step=None):
    if min_step is None:
        min_step = 10.0 * np.finfo(float).eps       # rk.py:833-834
    return HoTol(rtol, atol, max_step, min_step)


def system(kind=SYS_CR3BP6, mu=0.0, fwd=1, flip=None, ham=None):
    dim = {SYS_CR3BP6: 6, SYS_VAR42: 42}.get(kind)
    if kind == SYS_POLYHAM:
        dim = ham.dim
    lo, hi = (-1, -1) if flip is None else flip
    return HoSystem(kind, dim, float(mu), int(fwd), lo, hi, None if ham is None else ham.handle)


def crtbp_accel(state, mu):
    out = np.empty(6)
    s = np.ascontiguousarray(state, dtype=np.float64)
    lib().ho_crtbp_accel(_p(s), C.c_double(mu), _p(out))
    return out


def var_equations(phi, mu):
    out = np.empty(42)
    s = np.ascontiguousarray(phi, dtype=np.float64)
    lib().ho_var_equations(_p(s), C.c_double(mu), _p(out))
    return out


def adaptive_final(sys_, method, tol, y0, t0, tf):
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    yf = np.empty(sys_.dim)
    counts = np.zeros(2, dtype=np.int64)
    lib().ho_adaptive_final(C.byref(sys_), method, C.byref(tol), _p(y0), C.c_double(t0), C.c_double(tf), _p(yf),
                            counts.ctypes.data_as(ip))
    return yf, counts


def adaptive_dense(sys_, method, tol, y0, t_eval):
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    out = np.empty((t_eval.size, sys_.dim))
    counts = np.zeros(2, dtype=np.int64)
    lib().ho_adaptive_dense(C.byref(sys_), method, C.byref(tol), _p(y0), _p(t_eval), t_eval.size, _p(out),
                            counts.ctypes.data_as(ip))
    return out, counts


def adaptive_event(sys_, method, tol, ev, y0, t0, tmax):
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    th = C.c_double(0.0)
    yh = np.empty(sys_.dim)
    yl = np.empty(sys_.dim)
    counts = np.zeros(2, dtype=np.int64)
    hit = lib().ho_adaptive_event(C.byref(sys_), method, C.byref(tol), C.byref(ev), _p(y0), C.c_double(t0),
                                  C.c_double(tmax), C.byref(th), _p(yh), _p(yl), counts.ctypes.data_as(ip))
    return bool(hit), th.value, yh, yl, counts


def fixed_dense(sys_, method, y0, t_vals):
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    t_vals = np.ascontiguousarray(t_vals, dtype=np.float64)
    out = np.empty((t_vals.size, sys_.dim))
    lib().ho_fixed_dense(C.byref(sys_), method, _p(y0), _p(t_vals), t_vals.size, _p(out))
    return out


def fixed_event(sys_, method, ev, y0, t_vals):
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    t_vals = np.ascontiguousarray(t_vals, dtype=np.float64)
    th = C.c_double(0.0)
    yh = np.empty(sys_.dim)
    hit = lib().ho_fixed_event(C.byref(sys_), method, C.byref(ev), _p(y0), _p(t_vals), t_vals.size, C.byref(th), _p(yh))
    return bool(hit), th.value, yh


def batch_final(sys_, method, tol, y0, t0, tf, n_threads=1):
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    n = y0.shape[0]
    yf = np.empty_like(y0)
    counts = np.zeros((n, 2), dtype=np.int64)
    lib().ho_batch_final(C.byref(sys_), method, C.byref(tol), _p(y0), C.c_int64(n), C.c_double(t0), C.c_double(tf),
                         _p(yf), counts.ctypes.data_as(ip), n_threads)
    return yf, counts


def batch_dense(sys_, method, tol, y0, t_eval, n_threads=1):
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    t_eval = np.ascontiguousarray(t_eval, dtype=np.float64)
    n = y0.shape[0]
    out = np.empty((n, t_eval.size, sys_.dim))
    counts = np.zeros((n, 2), dtype=np.int64)
    lib().ho_batch_dense(C.byref(sys_), method, C.byref(tol), _p(y0), C.c_int64(n), _p(t_eval), t_eval.size,
                         _p(out), counts.ctypes.data_as(ip), n_threads)
    return out, counts


def synodic_detect(times, states, idx, offset=0.0, direction=0, proj=(0, 2), segment_refine=50,
                   tol_on_surface=1e-6, dedup_time_tol=1e-9, dedup_point_tol=1e-6, max_hits=0, cap=64):
    """Hits (times[K], states[K, dim]) of one sampled trajectory, reference detector semantics."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    states = np.ascontiguousarray(states, dtype=np.float64)
    m, dim = states.shape
    ht = np.empty(cap)
    hs = np.empty((cap, dim))
    lib().ho_synodic_detect.restype = C.c_int
    k = lib().ho_synodic_detect(_p(times), _p(states), m, dim, int(idx), C.c_double(offset), int(direction),
                                int(proj[0]), int(proj[1]), int(segment_refine), C.c_double(tol_on_surface),
                                C.c_double(dedup_time_tol), C.c_double(dedup_point_tol), int(max_hits), _p(ht),
                                _p(hs), cap)
    return ht[:k].copy(), hs[:k].copy()


def synodic_detect_cubic(times, states, idx, offset=0.0, direction=0, proj=(0, 2), segment_refine=50,
                         tol_on_surface=1e-6, dedup_time_tol=1e-9, dedup_point_tol=1e-6, max_hits=0, newton_max_iter=4,
                         cap=64):
    """The same with interp_kind="cubic" (backend.py:274-379, 541-645)."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    states = np.ascontiguousarray(states, dtype=np.float64)
    m, dim = states.shape
    ht = np.empty(cap)
    hs = np.empty((cap, dim))
    lib().ho_synodic_detect_cubic.restype = C.c_int
    k = lib().ho_synodic_detect_cubic(_p(times), _p(states), m, dim, int(idx), C.c_double(offset), int(direction),
                                      int(proj[0]), int(proj[1]), int(segment_refine), C.c_double(tol_on_surface),
                                      C.c_double(dedup_time_tol), C.c_double(dedup_point_tol), int(max_hits),
                                      int(newton_max_iter), _p(ht), _p(hs), cap)
    return ht[:k].copy(), hs[:k].copy()


class HoPolyHam(C.Structure):
    _fields_ = [("n_dof", C.c_int), ("max_deg", C.c_int), ("ptr", C.c_int64 * 7), ("deg", C.c_void_p),
                ("coef", C.c_void_p), ("exp", C.c_void_p)]


class PolyHam:
    """Oracle-side polynomial Hamiltonian table (keeps the numpy arrays alive)."""

    def __init__(self, ptr, deg, coef, exp):
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        self.deg = np.ascontiguousarray(deg, dtype=np.int32)
        self.coef = np.ascontiguousarray(coef, dtype=np.float64)
        self.exp = np.ascontiguousarray(exp, dtype=np.int32).reshape(-1, 6)
        self.dim = 6
        self.struct = HoPolyHam(3, int(self.exp.max()) if self.exp.size else 0, (C.c_int64 * 7)(*self.ptr.tolist()),
                                self.deg.ctypes.data, self.coef.ctypes.data, self.exp.ctypes.data)
        self.handle = C.addressof(self.struct)


SECTION = {"q2": 0, "p2": 1, "q3": 2, "p3": 3}


def polyham_rhs(ham, y):
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty(6)
    lib().ho_polyham_rhs(C.byref(ham.struct), _p(y), _p(out))
    return out


def cm_poincare_map(ham, seeds, dt, order, max_steps, use_symplectic, section, c_omega=20.0, n_threads=1):
    seeds = np.ascontiguousarray(seeds, dtype=np.float64)
    n = seeds.shape[0]
    flags = np.zeros(n, dtype=np.int64)
    out = np.zeros((n, 4))
    tt = np.zeros(n)
    rc = lib().ho_cm_poincare_map(C.byref(ham.struct), _p(seeds), C.c_int64(n), C.c_double(dt), int(order),
                                  int(max_steps), int(bool(use_symplectic)), SECTION[section], C.c_double(c_omega),
                                  flags.ctypes.data_as(ip), _p(out), _p(tt), int(n_threads))
    assert rc == 0
    return flags, out, tt


def symplectic_dense(ham, y0, t_vals_signed, order, c_omega=20.0):
    """_integrate_symplectic on the signed grid (t_vals * fwd): traj[m][6]."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    t = np.ascontiguousarray(t_vals_signed, dtype=np.float64)
    out = np.empty((t.size, 6))
    rc = lib().ho_symplectic_dense(C.byref(ham.struct), _p(y0), _p(t), int(t.size), int(order), C.c_double(c_omega),
                                   _p(out))
    assert rc == 0
    return out


def symplectic_event(ham, ev, y0, t_vals_signed, order, c_omega=20.0):
    """_integrate_symplectic_until_event: (hit, t_hit (signed-grid time), y_hit[6], traj[n_rows][6])."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    t = np.ascontiguousarray(t_vals_signed, dtype=np.float64)
    traj = np.zeros((t.size, 6))
    th = C.c_double(0.0)
    yh = np.empty(6)
    nr = C.c_int(0)
    hit = lib().ho_symplectic_event(C.byref(ham.struct), C.byref(ev), _p(y0), _p(t), int(t.size), int(order),
                                    C.c_double(c_omega), C.byref(th), _p(yh), _p(traj), C.byref(nr))
    assert hit in (0, 1)
    return bool(hit), th.value, yh, traj[: nr.value]


def batch_synodic_count(times, dense, idx, offset, direction, proj, segment_refine, tol_on_surface, dedup_time_tol,
                        dedup_point_tol, n_threads=1):
    """Total hit count over a uniformly sampled batch dense[N, m, dim] (bench.py CPU legs)."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    dense = np.ascontiguousarray(dense, dtype=np.float64)
    n, m, dim = dense.shape
    f = lib().ho_batch_synodic_count
    f.restype = C.c_int64
    return int(f(_p(times), _p(dense), C.c_int64(n), m, dim, int(idx), C.c_double(offset), int(direction),
                 int(proj[0]), int(proj[1]), int(segment_refine), C.c_double(tol_on_surface),
                 C.c_double(dedup_time_tol), C.c_double(dedup_point_tol), int(n_threads)))


def single_poly(deg, coef, exp):
    """Oracle table holding ONE polynomial (e.g. the Hamiltonian itself) as polynomial 0."""
    t = len(deg)
    return PolyHam([0, t, t, t, t, t, t], deg, coef, exp)


def cm_lift(ham_H, section, pts, h0, initial_guess=1e-3, expand_factor=2.0, max_expand=40, symmetric=False, xtol=1e-12):
    """lift_plane_point over a batch: returns ok[n] (int64), states[n, 4] = (q2, p2, q3, p3)."""
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n = pts.shape[0]
    ok = np.zeros(n, dtype=np.int64)
    out = np.zeros((n, 4))
    rc = lib().ho_cm_lift(C.byref(ham_H.struct), SECTION[section], _p(pts), C.c_int64(n), C.c_double(h0),
                          C.c_double(initial_guess), C.c_double(expand_factor), int(max_expand), int(bool(symmetric)),
                          C.c_double(xtol), ok.ctypes.data_as(ip), _p(out))
    assert rc == 0
    return ok, out


def connections(pu, ps, Xu, Xs, eps, dv_tol, bal_tol):
    """_ConnectionsBackend.run restated: dict of result arrays sorted by delta_v, plus pairs_considered."""
    pu, ps = np.ascontiguousarray(pu, dtype=np.float64), np.ascontiguousarray(ps, dtype=np.float64)
    Xu, Xs = np.ascontiguousarray(Xu, dtype=np.float64), np.ascontiguousarray(Xs, dtype=np.float64)
    n, m = len(pu), len(ps)
    cap = max(min(n, m), 1)
    kind, iu, is_ = (np.zeros(cap, dtype=np.int64) for _ in range(3))
    dv, pt, su, ss = np.zeros(cap), np.zeros((cap, 2)), np.zeros((cap, 6)), np.zeros((cap, 6))
    considered = C.c_int64(0)
    f = lib().ho_connections
    f.restype = C.c_int64
    k = f(_p(pu), C.c_int64(n), _p(ps), C.c_int64(m), _p(Xu), _p(Xs), C.c_double(eps), C.c_double(dv_tol),
          C.c_double(bal_tol), C.c_int64(cap), kind.ctypes.data_as(ip), _p(dv), _p(pt), _p(su), _p(ss),
          iu.ctypes.data_as(ip), is_.ctypes.data_as(ip), C.byref(considered))
    assert 0 <= k <= cap
    return dict(kind=kind[:k], dv=dv[:k], pt=pt[:k], su=su[:k], ss=ss[:k], iu=iu[:k], is_=is_[:k],
                pairs_considered=int(considered.value))


def manifold_ics(phi_dense, tt, period, eigvec_re, direction, fractions, displacements):
    """ho_manifold_ics: x0W[D*K, 6] (displacement-major) and the STM sample index of every fraction."""
    phi = np.ascontiguousarray(phi_dense, dtype=np.float64)
    tt = np.ascontiguousarray(tt, dtype=np.float64)
    ev = np.ascontiguousarray(eigvec_re, dtype=np.float64)
    fr = np.ascontiguousarray(np.atleast_1d(fractions), dtype=np.float64)
    dd = np.ascontiguousarray(np.atleast_1d(displacements), dtype=np.float64)
    out = np.empty((dd.size * fr.size, 6))
    idx = np.empty(fr.size, dtype=np.int64)
    lib().ho_manifold_ics(_p(phi), _p(tt), C.c_int64(tt.size), C.c_double(period), _p(ev), C.c_int(int(direction)),
                          _p(fr), C.c_int64(fr.size), _p(dd), C.c_int64(dd.size), _p(out),
                          idx.ctypes.data_as(ip))
    return out, idx


def tube_filter(states, mu):
    """ho_batch_tube_filter on states[n][m][6]: (min r1, min r2, max relative Jacobi error) per trajectory."""
    s = np.ascontiguousarray(states, dtype=np.float64)
    n, m = s.shape[0], s.shape[1]
    out = np.empty((n, 3))
    lib().ho_batch_tube_filter(_p(s), C.c_int64(n), C.c_int(m), C.c_double(mu), _p(out))
    return out


class HoCorrectOpts(C.Structure):
    _fields_ = [("ctrl", C.c_int * 2), ("res", C.c_int * 2), ("target", C.c_double * 2), ("event_idx", C.c_int),
                ("event_offset", C.c_double), ("halo_quadratic", C.c_int), ("finite_difference", C.c_int),
                ("tol", C.c_double), ("max_attempts", C.c_int), ("max_delta", C.c_double), ("fd_step", C.c_double),
                ("line_search", C.c_int), ("alpha_reduction", C.c_double), ("min_alpha", C.c_double),
                ("armijo_c", C.c_double)]


# family -> (control indices, residual indices, event index, halo quadratic term, finite differences)
# (algorithms/types/services/orbits.py:862-880, 1249-1262, 1440-1453)
CORRECTION_FAMILIES = {"halo": ((0, 4), (3, 5), 1, 1, 0), "lyapunov": ((4, 5), (3, 2), 1, 0, 0),
                       "vertical": ((5, 4), (3, 1), 2, 0, 1)}


def correct_opts(family, tol=1e-12, max_attempts=50, max_delta=1e-2, fd_step=1e-8, line_search=True,
                 alpha_reduction=0.5, min_alpha=1e-4, armijo_c=0.1):
    ctrl, res, ev, quad, fd = CORRECTION_FAMILIES[family]
    return HoCorrectOpts((C.c_int * 2)(*ctrl), (C.c_int * 2)(*res), (C.c_double * 2)(0.0, 0.0), ev, 0.0, quad, fd,
                         tol, max_attempts, max_delta, fd_step, int(line_search), alpha_reduction, min_alpha,
                         armijo_c)


def correct_orbits(x0, mu, opts):
    """ho_correct_orbit over rows of x0[N,6]: (x_corrected, half_period, iterations, residual_norm, status)."""
    x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
    n = len(x0)
    xc, hp, it, rn, st = np.empty((n, 6)), np.empty(n), np.zeros(n, np.int32), np.empty(n), np.zeros(n, np.int32)
    f = lib().ho_correct_orbit
    f.restype = C.c_int
    for i in range(n):
        xo, h, k, r = np.empty(6), C.c_double(), C.c_int(), C.c_double()
        st[i] = f(C.c_double(mu), C.byref(opts), _p(x0[i]), _p(xo), C.byref(h), C.byref(k), C.byref(r))
        xc[i], hp[i], it[i], rn[i] = xo, h.value, k.value, r.value
    return xc, hp, it, rn, st
