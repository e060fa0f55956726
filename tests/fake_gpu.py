"""Oracle-backed stand-ins for the GPU entry points, for CPU tests of the HOST plumbing only.

tests/test_dropin_reference.py runs the real reference (when /root/reference is present) with
hiten_b200.install() active and these fakes patched over hiten_b200.propagate / synodic / centermanifold, so
that argument translation, result types and filters of the drop-in are checked against the reference's own
objects without a GPU.  Product code never imports this module.
"""
import numpy as np

import oracle_lib as O
from hiten_b200.propagate import BatchResult
from hiten_b200.synodic import SectionHits


def _tol(integ):
    return O.HoTol(integ.rtol, integ.atol, integ.max_step, integ.min_step)


def cr3bp_dense(y0, mu, t_eval, *, forward=1, flip=None, integ=None, **kw):
    s = O.system(O.SYS_CR3BP6, mu, fwd=forward, flip=flip)
    if integ.method in (4, 6, 8):                      # fixed-step: one step per grid interval
        y0 = np.asarray(y0)
        dense = np.stack([O.fixed_dense(s, integ.method, y, np.asarray(t_eval)) for y in y0])
        z = np.zeros(len(y0), np.int32)
        return BatchResult(None, z, z, z, states=dense)
    dense, counts = O.batch_dense(s, integ.method, _tol(integ), np.asarray(y0), np.asarray(t_eval), 4)
    n = len(dense)
    return BatchResult(None, counts[:, 0].astype(np.int32), counts[:, 1].astype(np.int32), np.zeros(n, np.int32),
                       states=dense)


def cr3bp_stm_dense(x0, mu, t_eval, *, forward=1, flip=(36, 42), integ=None, **kw):
    s = O.system(O.SYS_VAR42, mu, fwd=forward, flip=flip)
    x0 = np.asarray(x0)
    y0 = np.concatenate([np.tile(np.eye(6).ravel(), (len(x0), 1)), x0], axis=1)
    method = O.DOP853 if integ is None else integ.method
    if method in (4, 6, 8):                             # fixed-step: one step per grid interval
        dense = np.stack([O.fixed_dense(s, method, y, np.asarray(t_eval)) for y in y0])
        counts = np.zeros((len(x0), 2), np.int64)
    else:
        dense, counts = O.batch_dense(s, method, _tol(integ), y0, np.asarray(t_eval), 4)
    return BatchResult(None, counts[:, 0].astype(np.int32), counts[:, 1].astype(np.int32),
                       np.zeros(len(x0), np.int32), states=dense)


def cr3bp_event(y0, mu, tmax, event_idx, *, event_offset=0.0, direction=0, xtol=1e-12, gtol=1e-12, t0=0.0,
                forward=1, flip=None, integ=None, **kw):
    s = O.system(O.SYS_CR3BP6, mu, fwd=forward, flip=flip)
    ev = O.HoEvent(int(event_idx), float(event_offset), int(direction), xtol, gtol)
    y0 = np.asarray(y0)
    yf, th, st = np.empty_like(y0), np.empty(len(y0)), np.zeros(len(y0), np.int32)
    for i in range(len(y0)):
        if integ.method in (4, 6, 8):
            hit, t, yh = O.fixed_event(s, integ.method, ev, y0[i], np.linspace(t0, tmax, integ.n_fixed_steps + 1))
        else:
            hit, t, yh, yl, _ = O.adaptive_event(s, integ.method, _tol(integ), ev, y0[i], t0, tmax)
        yf[i], th[i], st[i] = yh, t, 1 if hit else 0
    return BatchResult(yf, np.zeros(len(y0), np.int32), np.zeros(len(y0), np.int32), st, t_hit=th)


def detect(states, times, section, *, offsets=None, interp_kind="linear", newton_max_iter=4, **kw):
    states, times = np.asarray(states), np.asarray(times)
    cubic = interp_kind == "cubic"
    if offsets is None:
        n, m = states.shape[:2]
        offsets = np.arange(n + 1) * m
        states = states.reshape(-1, 6)
        if times.size == m:
            times = np.tile(times, n)
    ti, tt, ss = [], [], []
    per = np.zeros(len(offsets) - 1, np.int32)
    for k in range(len(offsets) - 1):
        a, b = offsets[k], offsets[k + 1]
        args = (times[a:b], states[a:b], section.idx, section.offset, section.direction,
                (section.proj_i, section.proj_j), section.segment_refine, section.tol_on_surface,
                section.dedup_time_tol, section.dedup_point_tol, section.max_hits_per_traj)
        t, x = (O.synodic_detect_cubic(*args, newton_max_iter=newton_max_iter, cap=256) if cubic
                else O.synodic_detect(*args, cap=256))
        per[k] = len(t)
        ti += [k] * len(t); tt += list(t); ss += list(x)
    ss = np.array(ss).reshape(-1, 6)
    return SectionHits(np.array(ti, dtype=np.int64), np.array(tt), ss, ss[:, [section.proj_i, section.proj_j]], per)


def poincare_map(table, seeds, opts, **kw):
    ham = O.PolyHam(table.ptr, table.deg, table.coef, table.exp)
    sec = {0: "q2", 1: "p2", 2: "q3", 3: "p3"}[opts.section]
    symp = opts.method == 2
    order = opts.order if symp else {4: 4, 6: 6, 8: 8}[opts.method]
    c_omega = 20.0
    if symp:                                   # recover c_omega from omega = (c*dt)^-order is not needed: sub_* carry it
        c_omega = getattr(opts, "_c_omega", 20.0)
    return O.cm_poincare_map(ham, seeds, opts.dt, order, opts.max_steps, symp, sec, c_omega, 4)


def lift_plane_points(H_table, section_coord, plane_points, h0, *, initial_guess=1e-3, expand_factor=2.0, max_expand=40,
                      symmetric=False, xtol=1e-12, **kw):
    ham = O.PolyHam(H_table.ptr, H_table.deg, H_table.coef, H_table.exp)
    ok, out = O.cm_lift(ham, section_coord, np.asarray(plane_points, dtype=np.float64), h0, initial_guess, expand_factor,
                        max_expand, symmetric, xtol)
    return ok.astype(bool), out


def find_connections(points_u, points_s, states_u, states_s, eps, dv_tol, bal_tol, *, traj_indices_u=None,
                     traj_indices_s=None, **kw):
    from hiten_b200.connections import Connections
    r = O.connections(points_u, points_s, states_u, states_s, eps, dv_tol, bal_tol)
    tu = np.asarray(traj_indices_u)[r["iu"]] if traj_indices_u is not None else np.zeros(len(r["iu"]), np.int64)
    ts = np.asarray(traj_indices_s)[r["is_"]] if traj_indices_s is not None else np.zeros(len(r["iu"]), np.int64)
    return Connections(r["kind"], r["dv"], r["pt"], r["su"], r["ss"], r["iu"], r["is_"], tu, ts, r["pairs_considered"])


def tube_initial_conditions(phi_dense, tt, period, eigvec, direction, fractions, displacements, **kw):
    x0, idx = O.manifold_ics(phi_dense, tt, period, np.asarray(eigvec).real, direction, fractions, displacements)
    return x0, idx.astype(np.int32)       # host [N, 6]: what the fake cr3bp_dense takes


def tube_filter(states, mu, *, safe_r1=0.0, safe_r2=0.0, energy_tol=np.inf, **kw):
    out = O.tube_filter(states, mu)
    keep = ~((out[:, 0] < safe_r1) | (out[:, 1] < safe_r2)) & ~(out[:, 2] > energy_tol)
    return out, keep.astype(np.int32)


def correct_orbits(x0, mu, opts, **kw):
    from hiten_b200.corrector import CorrectionBatch
    o = O.HoCorrectOpts((O.C.c_int * 2)(*opts.ctrl), (O.C.c_int * 2)(*opts.res), (O.C.c_double * 2)(*opts.target),
                        opts.event_idx, opts.event_offset, opts.halo_quadratic, opts.finite_difference, opts.tol,
                        opts.max_attempts, opts.max_delta, opts.fd_step, opts.line_search, opts.alpha_reduction,
                        opts.min_alpha, opts.armijo_c)
    xc, half, it, rn, st = O.correct_orbits(x0, mu, o)
    return CorrectionBatch(xc, half, it, rn, st, 0, 0)


def integrate_symplectic(table, y0, t_vals_signed, order, *, c_omega_heuristic=20.0, **kw):
    ham = O.PolyHam(table.ptr, table.deg, table.coef, table.exp)
    return np.stack([O.symplectic_dense(ham, y, t_vals_signed, order, c_omega_heuristic) for y in np.asarray(y0)])


def integrate_symplectic_until_event(table, y0, t_vals_signed, order, event, *, c_omega_heuristic=20.0,
                                     want_trajectory=False, **kw):
    from hiten_b200.symplectic import SymplecticEventResult
    ham = O.PolyHam(table.ptr, table.deg, table.coef, table.exp)
    idx, offset, direction, xtol, gtol = event
    ev = O.HoEvent(int(idx), float(offset), int(direction), xtol, gtol)
    y0 = np.asarray(y0)
    m = len(t_vals_signed)
    hit, th, yh, nr, traj = np.zeros(len(y0), bool), np.zeros(len(y0)), np.zeros((len(y0), 6)), \
        np.zeros(len(y0), np.int64), np.zeros((len(y0), m, 6))
    for i in range(len(y0)):
        hit[i], th[i], yh[i], rows = O.symplectic_event(ham, ev, y0[i], t_vals_signed, order, c_omega_heuristic)
        nr[i] = len(rows)
        traj[i, : len(rows)] = rows
    return SymplecticEventResult(hit, th, yh, nr, traj if want_trajectory else None)


def integrate_rk_ham(table, y0, t_vals, order, *, want_derivatives=True, **kw):
    ham = O.PolyHam(table.ptr, table.deg, table.coef, table.exp)
    s = O.system(O.SYS_POLYHAM, ham=ham)
    states = np.stack([O.fixed_dense(s, int(order), y, np.asarray(t_vals)) for y in np.asarray(y0)])
    derivs = np.stack([[O.polyham_rhs(ham, row) for row in tr] for tr in states]) if want_derivatives else None
    return states, derivs


def integrate_rk_ham_until_event(table, y0, t_vals, order, event, *, want_trajectory=False, **kw):
    from hiten_b200.symplectic import SymplecticEventResult
    ham = O.PolyHam(table.ptr, table.deg, table.coef, table.exp)
    s = O.system(O.SYS_POLYHAM, ham=ham)
    idx, offset, direction, xtol, gtol = event
    ev = O.HoEvent(int(idx), float(offset), int(direction), xtol, gtol)
    y0 = np.asarray(y0)
    hit, th, yh = np.zeros(len(y0), bool), np.zeros(len(y0)), np.zeros((len(y0), 6))
    for i in range(len(y0)):
        hit[i], th[i], yh[i] = O.fixed_event(s, int(order), ev, y0[i], np.asarray(t_vals))
    return SymplecticEventResult(hit, th, yh, np.zeros(len(y0), np.int64), None)


def integrate_adaptive_ham(table, y0, t_eval, *, integ=None, want_derivatives=True, **kw):
    from hiten_b200.symplectic import HamAdaptiveResult
    ham = O.PolyHam(table.ptr, table.deg, table.coef, table.exp)
    s = O.system(O.SYS_POLYHAM, ham=ham)
    y0 = np.asarray(y0)
    states = np.stack([O.adaptive_dense(s, integ.method, _tol(integ), y, np.asarray(t_eval))[0] for y in y0])
    derivs = np.stack([[O.polyham_rhs(ham, row) for row in tr] for tr in states]) if want_derivatives else None
    z = np.zeros(len(y0), np.int32)
    return HamAdaptiveResult(states, derivs, None, None, z, z, z)


def integrate_adaptive_ham_until_event(table, y0, t0, tmax, event, *, integ=None, **kw):
    from hiten_b200.symplectic import HamAdaptiveResult
    ham = O.PolyHam(table.ptr, table.deg, table.coef, table.exp)
    s = O.system(O.SYS_POLYHAM, ham=ham)
    idx, offset, direction, xtol, gtol = event
    ev = O.HoEvent(int(idx), float(offset), int(direction), xtol, gtol)
    y0 = np.asarray(y0)
    th, yh, st = np.zeros(len(y0)), np.zeros((len(y0), 6)), np.zeros(len(y0), np.int32)
    for i in range(len(y0)):
        hit, t, y, yl, _ = O.adaptive_event(s, integ.method, _tol(integ), ev, y0[i], t0, tmax)
        th[i], yh[i], st[i] = t, y, 1 if hit else 0
    z = np.zeros(len(y0), np.int32)
    return HamAdaptiveResult(None, None, th, yh, z, z, st)


def patch(monkeypatch):
    import hiten_b200.corrector as corr
    monkeypatch.setattr(corr, "correct_orbits", correct_orbits)
    import hiten_b200.manifold as man
    monkeypatch.setattr(man, "tube_initial_conditions", tube_initial_conditions)
    monkeypatch.setattr(man, "tube_filter", tube_filter)
    import hiten_b200.centermanifold as cm
    import hiten_b200.propagate as prop
    import hiten_b200.synodic as syn
    monkeypatch.setattr(prop, "cr3bp_dense", cr3bp_dense)
    monkeypatch.setattr(prop, "cr3bp_stm_dense", cr3bp_stm_dense)
    monkeypatch.setattr(prop, "cr3bp_event", cr3bp_event)
    monkeypatch.setattr(syn, "detect", detect)
    monkeypatch.setattr(cm, "poincare_map", poincare_map)
    monkeypatch.setattr(cm, "lift_plane_points", lift_plane_points)
    import hiten_b200.symplectic as symp
    monkeypatch.setattr(symp, "integrate_symplectic", integrate_symplectic)
    monkeypatch.setattr(symp, "integrate_symplectic_until_event", integrate_symplectic_until_event)
    monkeypatch.setattr(symp, "integrate_rk_ham", integrate_rk_ham)
    monkeypatch.setattr(symp, "integrate_adaptive_ham", integrate_adaptive_ham)
    monkeypatch.setattr(symp, "integrate_adaptive_ham_until_event", integrate_adaptive_ham_until_event)
    monkeypatch.setattr(symp, "integrate_rk_ham_until_event", integrate_rk_ham_until_event)
    import hiten_b200.connections as conn
    monkeypatch.setattr(conn, "find_connections", find_connections)
