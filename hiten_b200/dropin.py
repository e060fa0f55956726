"""Drop-in installer: rebinds HITEN's propagation funnels to the GPU path.

    import hiten, hiten_b200
    hiten_b200.install()          # orbit.manifold().compute(), SynodicMap.compute(), cm.poincare_map().compute(),
                                  # _propagate_dynsys / _compute_stm / the DOP853 integrator class now run on the GPU
    hiten_b200.uninstall()

No reference file is edited.  What is rebound (paths relative to hiten/):
  * `_propagate_dynsys`  (algorithms/dynamics/base.py:346) in every module that imported it by value
    -> CR3BP 6-state and 42-state systems with method="adaptive", order=8 go to the DOP853 kernels
       (dense grid, or a recognised plane event); `_compute_stm` (algorithms/dynamics/rtbp.py:258) follows
       because it calls the rebound name;
  * `_ManifoldDynamicsService._run_compute` (algorithms/types/services/manifold.py:293): the serial fraction loop
    becomes ONE batched dense propagation;
  * `_SynodicDetectionBackend.run` (algorithms/poincare/synodic/backend.py:823);
  * `_CenterManifoldBackend.run` (algorithms/poincare/centermanifold/backend.py:404);
  * `_DOP853.integrate` (algorithms/integrators/rk.py:2221), `_RK45.integrate` (:1138) and
    `_FixedStepRK.integrate` (:422; _RK4 / _RK6 / _RK8) for the 6-state CR3BP system;
  * `_ExtendedSymplectic.integrate` (algorithms/integrators/symplectic.py:877) for the polynomial Hamiltonian systems
    (grid and plane-event forms; `_propagate_dynsys(method="symplectic")` builds this class);
  * the `_ham` branches of the three RK classes above for the reference's bare `_HamiltonianSystem` (grid with
    derivatives, plane events).
Anything the GPU path cannot express (user-defined RHS or event callables, a `_DirectedSystem` around a Hamiltonian
system in the RK classes -- which raises inside the reference --, the 42-state system handed directly to the RK45 /
fixed-step CLASSES instead of through _propagate_dynsys / _compute_stm) is handed to the reference's ORIGINAL function -- that is the reference's
own code for inputs outside this path, not a fallback of the kernels: for recognised inputs a missing library
or GPU raises.
"""
import sys

import numpy as np

from . import _lib as _L
from . import centermanifold as _cm
from . import connections as _conn
from . import manifold as _man
from . import propagate as _prop
from . import symplectic as _symp
from . import synodic as _syn

_STATE = {"installed": False, "orig": {}, "patched_modules": [], "arith": "parity"}


# ------------------------------------------------------------------------------------------------
# recognition helpers
# ------------------------------------------------------------------------------------------------
def _norm_flip(flip, dim):
    """flip_indices of _DirectedSystem -> (lo, hi) or None (= all); raises LookupError if not a contiguous range."""
    if flip is None:
        return None
    if isinstance(flip, slice):
        lo, hi, st = flip.indices(dim)
        if st != 1:
            raise LookupError("strided flip")
        return (lo, hi)
    idx = np.asarray(flip, dtype=np.int64).ravel()
    if idx.size == 0:
        raise LookupError("empty flip")
    if np.any(idx < -dim) or np.any(idx >= dim):
        raise LookupError("flip index out of range")           # the reference's own IndexError
    idx = idx % dim                                             # numpy's negative indexing: -1 is the last component
    srt = np.sort(idx)
    if not np.array_equal(srt, np.arange(srt[0], srt[0] + idx.size)):
        raise LookupError("non-contiguous flip")
    return (int(srt[0]), int(srt[0]) + idx.size)


def recognise_system(dynsys):
    """-> (dim, mu, fwd, flip) for the CR3BP 6- or 42-state systems (possibly wrapped by _DirectedSystem), else None."""
    from hiten.algorithms.dynamics.base import _DirectedSystem
    from hiten.algorithms.dynamics.rtbp import _RTBPRHS, _VarEqRHS
    fwd, flip, base = 1, None, dynsys
    if isinstance(dynsys, _DirectedSystem):
        base = dynsys._base
        fwd = int(dynsys._fwd)
        try:
            flip = _norm_flip(dynsys._flip_idx, dynsys.dim)
        except LookupError:
            return None
        if isinstance(base, _DirectedSystem):
            return None
    if type(base) is _RTBPRHS:
        return 6, float(base.mu), fwd, flip
    if type(base) is _VarEqRHS:
        return 42, float(base.mu), fwd, flip
    return None


def recognise_event(event_fn):
    """-> (idx, offset) for g(t,y) = y[idx] - offset plane events (singlehit/backend.py:30-67), else None."""
    from hiten.algorithms.poincare.singlehit import backend as sh
    for i, fn in enumerate((sh._g_x0, sh._g_y0, sh._g_z0)):
        if event_fn is fn:
            return i, 0.0
    for (idx, off), fn in list(sh._PLANE_EVENT_FN_CACHE.items()):
        if event_fn is fn:
            return int(idx), float(off)
    return None


def _integ(kwargs=None, rtol=None, atol=None, max_step=None):
    kwargs = kwargs or {}
    return _prop.make_integ(arith=_STATE["arith"],
                            rtol=kwargs.get("rtol", 1e-12) if rtol is None else rtol,
                            atol=kwargs.get("atol", 1e-12) if atol is None else atol,
                            max_step=kwargs.get("max_step", 1e4) if max_step is None else max_step)


def _hb_method(method, order):
    """(method, order) of _propagate_dynsys -> HB_* integrator id, or None (base.py:430-453)."""
    if method == "adaptive":
        return {8: _L.HB_DOP853, 5: _L.HB_RK45}.get(order)
    if method == "fixed":
        return {4: _L.HB_RK4, 6: _L.HB_RK6, 8: _L.HB_RK8}.get(order)
    return None


def _rhs6_numpy(states, mu, fwd, flip):
    """Vectorised _crtbp_accel (+ direction wrapper) for the `derivatives` field of _Solution."""
    x, y, z, vx, vy, vz = states.T
    r1 = np.sqrt((x + mu) ** 2 + y ** 2 + z ** 2)
    r2 = np.sqrt((x - (1 - mu)) ** 2 + y ** 2 + z ** 2)
    ax = 2 * vy + x - (1 - mu) * (x + mu) / r1 ** 3 - mu * (x - 1 + mu) / r2 ** 3
    ay = -2 * vx + y - (1 - mu) * y / r1 ** 3 - mu * y / r2 ** 3
    az = -(1 - mu) * z / r1 ** 3 - mu * z / r2 ** 3
    out = np.column_stack([vx, vy, vz, ax, ay, az])
    if fwd == -1:
        lo, hi = (0, 6) if flip is None else flip
        out[:, lo:hi] *= -1
    return out


# ------------------------------------------------------------------------------------------------
# shared GPU integration of ONE trajectory on a grid / to an event (used by the two integrate funnels)
# ------------------------------------------------------------------------------------------------
def _gpu_integrate(dim, mu, fwd, flip, y0, t_vals, integ, event=None):
    """Returns (times, states) like _DOP853.integrate: the dense grid, or the 2-row event solution."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    t_vals = np.ascontiguousarray(t_vals, dtype=np.float64)
    if event is not None:
        idx, off, direction, xtol, gtol = event
        if dim != 6:
            raise LookupError("event propagation is a 6-state path")
        res = _prop.cr3bp_event(y0[None, :], mu, float(t_vals[-1]), idx, event_offset=off, direction=direction,
                                xtol=xtol, gtol=gtol, t0=float(t_vals[0]), forward=fwd, flip=flip, integ=integ)
        _raise_on_status(res.status)
        if int(res.status[0]) == 1:
            return np.array([t_vals[0], res.t_hit[0]]), np.vstack([y0, res.yf[0]])
        return np.array([t_vals[0], t_vals[-1]]), np.vstack([y0, res.yf[0]])
    if dim == 6:
        res = _prop.cr3bp_dense(y0[None, :], mu, t_vals, forward=fwd, flip=flip, integ=integ)
        _raise_on_status(res.status)
        return t_vals.copy(), res.states[0]
    if y0.shape != (42,) or not np.array_equal(y0[:36], np.eye(6).ravel()):
        raise LookupError("42-state kernel starts from PHI0 = [I, x0]")
    res = _prop.cr3bp_stm_dense(y0[None, 36:], mu, t_vals, forward=fwd, flip=(0, 42) if flip is None else flip,
                                integ=integ)
    _raise_on_status(res.status)
    return t_vals.copy(), res.states[0]


def _host(a):
    """Device tensor -> numpy (one D2H copy); numpy stays numpy."""
    return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)


def _raise_on_status(status):
    bad = np.nonzero(np.asarray(status) > 1)[0]
    if bad.size:
        from hiten.algorithms.types.exceptions import ConvergenceError
        raise ConvergenceError(f"GPU propagation failed for trajectory {int(bad[0])} (status {int(status[bad[0]])})")


# ------------------------------------------------------------------------------------------------
# replacements
# ------------------------------------------------------------------------------------------------
def _make_propagate_dynsys(orig):
    def _propagate_dynsys(dynsys, state0, t0, tf, forward=1, steps=1000, method="adaptive", order=8,
                          flip_indices=None, **kwargs):
        from hiten.algorithms.dynamics.base import _validate_initial_state
        from hiten.algorithms.integrators.types import _Solution
        rec = recognise_system(dynsys)
        known = {"rtol", "atol", "max_step", "event_fn", "event_cfg", "event_options"}
        hb_method = _hb_method(method, order)
        if rec is None or rec[2] != 1 or hb_method is None or (set(kwargs) - known):
            return orig(dynsys, state0, t0, tf, forward=forward, steps=steps, method=method, order=order,
                        flip_indices=flip_indices, **kwargs)
        dim, mu, _, _ = rec
        try:
            flip = _norm_flip(flip_indices, dim)
        except LookupError:
            return orig(dynsys, state0, t0, tf, forward=forward, steps=steps, method=method, order=order,
                        flip_indices=flip_indices, **kwargs)
        fwd = 1 if forward >= 0 else -1
        event = None
        event_fn = kwargs.get("event_fn")
        if event_fn is not None:
            ev = recognise_event(event_fn)
            if ev is None or dim != 6:
                return orig(dynsys, state0, t0, tf, forward=forward, steps=steps, method=method, order=order,
                            flip_indices=flip_indices, **kwargs)
            cfg, opt = kwargs.get("event_cfg"), kwargs.get("event_options")
            event = (ev[0], ev[1], 0 if cfg is None else int(cfg.direction),
                     1e-12 if opt is None else float(opt.xtol), 1e-12 if opt is None else float(opt.gtol))
        state0_np = _validate_initial_state(state0, dynsys.dim)
        t_eval = np.linspace(t0, tf, steps)
        if steps >= 2 and np.isclose(t_eval[0], t_eval[-1]):          # base.py:421-424
            return _Solution(forward * t_eval, np.repeat(state0_np[None, :], repeats=len(t_eval), axis=0))
        if hb_method == _L.HB_DOP853:
            integ = _integ(kwargs)
        elif hb_method == _L.HB_RK45:                     # AdaptiveRK(order=5, max_step, rtol, atol), base.py:446-450
            integ = _prop.make_integ(method=_L.HB_RK45, arith=_STATE["arith"], rtol=kwargs.get("rtol", 1e-12),
                                     atol=kwargs.get("atol", 1e-12), max_step=kwargs.get("max_step", 1e4))
        else:                                             # RungeKutta(order=4|6|8): one step per grid interval
            integ = _prop.make_integ(method=hb_method, arith=_STATE["arith"], n_fixed_steps=steps - 1)
        try:
            times, states = _gpu_integrate(dim, mu, fwd, flip, state0_np, t_eval, integ, event)
        except LookupError:
            return orig(dynsys, state0, t0, tf, forward=forward, steps=steps, method=method, order=order,
                        flip_indices=flip_indices, **kwargs)
        return _Solution(forward * times, states)

    _propagate_dynsys.__wrapped__ = orig
    _propagate_dynsys.__doc__ = orig.__doc__
    return _propagate_dynsys


def _make_dop853_integrate(orig):
    def integrate(self, system, y0, t_vals, *, event_fn=None, event_cfg=None, event_options=None, **kwargs):
        from hiten.algorithms.integrators.types import _Solution
        rec = recognise_system(system)
        ev = None
        if event_fn is not None:
            ev = recognise_event(event_fn)
        if rec is None and not kwargs and (event_fn is None or ev is not None):
            sol = _adaptive_ham(self, system, y0, t_vals, _L.HB_DOP853, ev, event_cfg, event_options)
            if sol is not None:
                return sol
        if rec is None or (event_fn is not None and (ev is None or rec[0] != 6)):
            return orig(self, system, y0, t_vals, event_fn=event_fn, event_cfg=event_cfg,
                        event_options=event_options, **kwargs)
        self.validate_inputs(system, y0, t_vals)
        const = self._maybe_constant_solution(system, y0, t_vals)
        if const is not None:
            return const
        t_vals = np.asarray(t_vals, dtype=np.float64)
        if not np.all(np.diff(t_vals) > 0):
            return orig(self, system, y0, t_vals, event_fn=event_fn, event_cfg=event_cfg,
                        event_options=event_options, **kwargs)
        dim, mu, fwd, flip = rec
        integ = _prop.make_integ(arith=_STATE["arith"], rtol=self._rtol, atol=self._atol,
                                 max_step=min(float(self._max_step), 1e300), min_step=self._min_step)
        event = None
        if ev is not None:
            event = (ev[0], ev[1], 0 if event_cfg is None else int(event_cfg.direction),
                     float(event_options.xtol if event_options is not None else 1.0e-12),
                     float(event_options.gtol if event_options is not None else 1.0e-12))
        try:
            times, states = _gpu_integrate(dim, mu, fwd, flip, np.asarray(y0, dtype=np.float64), t_vals, integ, event)
        except LookupError:
            return orig(self, system, y0, t_vals, event_fn=event_fn, event_cfg=event_cfg,
                        event_options=event_options, **kwargs)
        if event is not None:
            return _Solution(times=times, states=states)
        derivs = _rhs6_numpy(states, mu, fwd, flip) if dim == 6 else None
        return _Solution(times=times, states=states, derivatives=derivs)

    integrate.__wrapped__ = orig
    integrate.__doc__ = orig.__doc__
    return integrate


def _adaptive_ham(self, system, y0, t_vals, method, ev, event_cfg, event_options):
    """The `_ham` branch of _DOP853.integrate (rk.py:2256-2360) / _RK45.integrate (:1170-1262) for the reference's bare
    polynomial `_HamiltonianSystem`.  Returns None when the system is not that (or the grid is not ascending)."""
    from hiten.algorithms.dynamics.hamiltonian import _HamiltonianSystem
    from hiten.algorithms.integrators.types import _Solution
    if type(system) is not _HamiltonianSystem or system.n_dof != 3:
        return None
    self.validate_inputs(system, y0, t_vals)
    const = self._maybe_constant_solution(system, y0, t_vals)
    if const is not None:
        return const
    t_vals = np.asarray(t_vals, dtype=np.float64)
    if not np.all(np.diff(t_vals) > 0):
        return None
    y0 = np.asarray(y0, dtype=np.float64)
    table = _poly_table(system.jac_H, system.clmo_H)
    integ = _prop.make_integ(method=method, arith=_STATE["arith"], rtol=self._rtol, atol=self._atol,
                             max_step=min(float(self._max_step), 1e300), min_step=self._min_step)
    if ev is None:
        r = _symp.integrate_adaptive_ham(table, y0[None, :], t_vals, integ=integ)
        _raise_on_status(r.status)
        return _Solution(times=t_vals.copy(), states=r.states[0], derivatives=r.derivatives[0])
    event = (ev[0], ev[1], 0 if event_cfg is None else int(event_cfg.direction),
             float(event_options.xtol if event_options is not None else 1.0e-12),
             float(event_options.gtol if event_options is not None else 1.0e-12))
    r = _symp.integrate_adaptive_ham_until_event(table, y0[None, :], float(t_vals[0]), float(t_vals[-1]), event,
                                                 integ=integ)
    _raise_on_status(r.status)
    t_end = float(r.t_hit[0]) if int(r.status[0]) == 1 else t_vals[-1]
    return _Solution(times=np.array([t_vals[0], t_end], dtype=np.float64), states=np.vstack([y0, r.y_hit[0]]))


def _fixed_rk_ham(self, system, y0, t_vals, method, ev, event_cfg, event_options):
    """The `_ham` branch of _FixedStepRK.integrate (rk.py:468-530) for the reference's bare polynomial
    `_HamiltonianSystem` (a _DirectedSystem around it never reaches the `_ham` kernels in the reference -- it raises
    there -- and is left to the reference's own method).  Returns None when the system is not that."""
    from hiten.algorithms.dynamics.hamiltonian import _HamiltonianSystem
    from hiten.algorithms.integrators.types import _Solution
    if type(system) is not _HamiltonianSystem or system.n_dof != 3:
        return None
    self.validate_inputs(system, y0, t_vals)
    const = self._maybe_constant_solution(system, y0, t_vals)
    if const is not None:
        return const
    t_vals = np.asarray(t_vals, dtype=np.float64)
    y0 = np.asarray(y0, dtype=np.float64)
    table = _poly_table(system.jac_H, system.clmo_H)
    order = {_L.HB_RK4: 4, _L.HB_RK6: 6, _L.HB_RK8: 8}[method]
    if ev is None:
        states, derivs = _symp.integrate_rk_ham(table, y0[None, :], t_vals, order, arith=_STATE["arith"])
        return _Solution(times=t_vals[: states.shape[1]], states=states[0], derivatives=derivs[0])
    event = (ev[0], ev[1], 0 if event_cfg is None else int(event_cfg.direction),
             float(event_options.xtol if event_options is not None else 1.0e-12),
             float(event_options.gtol if event_options is not None else 1.0e-12))
    r = _symp.integrate_rk_ham_until_event(table, y0[None, :], t_vals, order, event, arith=_STATE["arith"])
    t_end = float(r.t_hit[0]) if bool(r.hit[0]) else t_vals[-1]                    # rk.py:512-515
    return _Solution(times=np.array([t_vals[0], t_end], dtype=np.float64), states=np.vstack([y0, r.y_hit[0]]))


def _make_rk_integrate(orig, kind):
    """_RK45.integrate (rk.py:1138) / _FixedStepRK.integrate (rk.py:422) for the 6-state CR3BP system: grid
    integration (hb_cr3bp_dense with the class's method) and recognised plane events (hb_cr3bp_event).  The fixed-step
    event scan runs on the GPU only for a uniform grid (the kernel rebuilds linspace(t0, tf, n + 1)); anything else
    goes to the reference's own method."""
    def integrate(self, system, y0, t_vals, *, event_fn=None, event_cfg=None, event_options=None, **kwargs):
        from hiten.algorithms.integrators.types import _Solution
        rec = recognise_system(system)
        ev = recognise_event(event_fn) if event_fn is not None else None
        method = _L.HB_RK45 if kind == "rk45" else {"_RK4": _L.HB_RK4, "_RK6": _L.HB_RK6,
                                                    "_RK8": _L.HB_RK8}.get(type(self).__name__)
        if rec is None and method is not None and not kwargs and (event_fn is None or ev is not None):
            sol = (_fixed_rk_ham if kind == "fixed" else _adaptive_ham)(self, system, y0, t_vals, method, ev, event_cfg,
                                                                        event_options)
            if sol is not None:
                return sol
        if rec is None or rec[0] != 6 or method is None or (event_fn is not None and ev is None) or kwargs:
            return orig(self, system, y0, t_vals, event_fn=event_fn, event_cfg=event_cfg,
                        event_options=event_options, **kwargs)
        self.validate_inputs(system, y0, t_vals)
        if kind == "fixed":
            const = self._maybe_constant_solution(system, y0, t_vals)
            if const is not None:
                return const
        t_vals = np.asarray(t_vals, dtype=np.float64)
        uniform = t_vals.size >= 2 and np.array_equal(t_vals, np.linspace(t_vals[0], t_vals[-1], t_vals.size))
        if not np.all(np.diff(t_vals) > 0) or (kind == "fixed" and event_fn is not None and not uniform):
            return orig(self, system, y0, t_vals, event_fn=event_fn, event_cfg=event_cfg,
                        event_options=event_options, **kwargs)
        dim, mu, fwd, flip = rec
        if kind == "rk45":
            integ = _prop.make_integ(method=method, arith=_STATE["arith"], rtol=self._rtol, atol=self._atol,
                                     max_step=min(float(self._max_step), 1e300), min_step=self._min_step)
        else:
            integ = _prop.make_integ(method=method, arith=_STATE["arith"], n_fixed_steps=t_vals.size - 1)
        event = None
        if ev is not None:
            event = (ev[0], ev[1], 0 if event_cfg is None else int(event_cfg.direction),
                     float(event_options.xtol if event_options is not None else 1.0e-12),
                     float(event_options.gtol if event_options is not None else 1.0e-12))
        times, states = _gpu_integrate(dim, mu, fwd, flip, np.asarray(y0, dtype=np.float64), t_vals, integ, event)
        if event is not None:
            return _Solution(times=times, states=states)
        return _Solution(times=times, states=states, derivatives=_rhs6_numpy(states, mu, fwd, flip))

    integrate.__wrapped__ = orig
    integrate.__doc__ = orig.__doc__
    return integrate


def _make_run_compute(orig):
    def _run_compute(self, *, step, integration_fraction, NN, displacement, method, order, dt, energy_tol,
                     safe_distance, show_progress):
        if method != "adaptive" or order != 8:
            return orig(self, step=step, integration_fraction=integration_fraction, NN=NN,
                        displacement=displacement, method=method, order=order, dt=dt, energy_tol=energy_tol,
                        safe_distance=safe_distance, show_progress=show_progress)
        orbit = self.orbit
        mu, forward = self.mu, self.forward
        dist_m = self.system.distance * 1e3                                 # manifold.py:341-345 (kept as is)
        safe_r1 = safe_distance * (self.system.primary.radius / dist_m)
        safe_r2 = safe_distance * (self.system.secondary.radius / dist_m)
        sn, un, _ = self.eigenvalues
        Ws, Wu, _ = self.eigenvectors
        _, snreal_vecs = self.stability.get_real_eigenvectors(Ws, sn)
        _, unreal_vecs = self.stability.get_real_eigenvectors(Wu, un)
        col_idx = NN - 1
        vecs, label = (snreal_vecs, "stable") if self.stable == 1 else (unreal_vecs, "unstable")
        if vecs.shape[1] <= col_idx or col_idx < 0:
            raise ValueError(f"Requested {label} eigenvector {NN} not available. "
                             f"Only {vecs.shape[1]} real {label} eigenvectors found.")
        eigvec = vecs[:, col_idx]
        fractions = np.arange(0.0, 1.0, step)
        ysos, dysos, states_list, times_list = [], [], [], []
        attempts = len(fractions)
        if attempts == 0:
            return (ysos, dysos, states_list, times_list, 0, 0)
        if np.any(np.asarray(eigvec).imag != 0.0):
            # MAN.real of a genuinely complex eigenvector: not the device path's contract -> the reference's loop
            return orig(self, step=step, integration_fraction=integration_fraction, NN=NN,
                        displacement=displacement, method=method, order=order, dt=dt, energy_tol=energy_tol,
                        safe_distance=safe_distance, show_progress=show_progress)
        _, tt, _, PHI = self.compute_stm(steps=2000)
        tf = integration_fraction * 2 * np.pi
        steps = max(int(abs(tf) / dt) + 1, 100)
        t_eval = np.linspace(0.0, tf, steps)
        # the batch axis: initial conditions of every fraction (manifold.py:470-537), ONE dense propagation
        # (replaces the loop at manifold.py:381-440) and the two filters (manifold.py:412-432), all in HBM;
        # only the tubes that pass come back to the host
        x0W, _ = _man.tube_initial_conditions(PHI, tt, orbit.period, eigvec, self.direction, fractions,
                                              [displacement])
        res = _prop.cr3bp_dense(x0W, mu, t_eval, forward=forward, flip=(0, 6), integ=_integ(), keep_on_device=True)
        _, keep = _man.tube_filter(res.states, mu, safe_r1=safe_r1, safe_r2=safe_r2, energy_tol=energy_tol)
        keep = _host(keep).astype(bool) & (_host(res.status) == 0)         # "discard and continue", manifold.py:438-440
        sel = np.nonzero(keep)[0]
        times = forward * t_eval
        if sel.size:
            kept = _host(res.states[sel] if sel.size < attempts else res.states)
            states_list = [kept[i] for i in range(sel.size)]
            times_list = [times.copy() for _ in range(sel.size)]
        successes = int(sel.size)
        return (ysos, dysos, states_list, times_list, successes, attempts)

    _run_compute.__wrapped__ = orig
    return _run_compute


def _make_synodic_run(orig):
    def run(self, request):
        from hiten.algorithms.poincare.core.types import _SectionHit
        from hiten.algorithms.poincare.synodic.types import SynodicBackendResponse
        normal = np.asarray(request.normal, dtype=float).ravel()
        nz = np.nonzero(normal)[0]
        trajs = list(request.trajectories)
        one_hot = normal.size == 6 and nz.size == 1 and normal[nz[0]] == 1.0
        cubic = request.interp_kind == "cubic"         # the reference's own test (backend.py:762)
        names_ok = all(isinstance(c, str) and c.lower() in _syn.IDX for c in request.plane_coords)
        if not (one_hot and names_ok) or any(np.asarray(s).shape[1] != 6 for _, s in trajs if len(s)):
            return orig(self, request)
        sec = _syn.make_section(int(nz[0]), float(request.offset), request.plane_coords, request.direction,
                                int(request.segment_refine), request.tol_on_surface, request.dedup_time_tol,
                                request.dedup_point_tol, request.max_hits_per_traj)
        idxs = [int(i) for i in request.trajectory_indices]
        n = min(len(idxs), len(trajs))
        lens = [len(trajs[k][0]) for k in range(n)]
        hits = [[] for _ in range(n)]
        if n and sum(lens):
            states = np.concatenate([np.asarray(trajs[k][1], dtype=np.float64).reshape(-1, 6) for k in range(n)])
            times = np.concatenate([np.asarray(trajs[k][0], dtype=np.float64).ravel() for k in range(n)])
            off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            got = _syn.detect(states, times, sec, offsets=off, interp_kind="cubic" if cubic else "linear",
                              newton_max_iter=int(request.newton_max_iter))
            for k, t, s, p in zip(got.trajectory_indices, got.times, got.states, got.points):
                hits[int(k)].append(_SectionHit(time=float(t), state=s.copy(), point2d=p.copy(),
                                                trajectory_index=idxs[int(k)]))
        pts, sts, ts, tis = [], [], [], []
        for th in hits:
            for h in th:
                pts.append(h.point2d); sts.append(h.state); ts.append(h.time); tis.append(h.trajectory_index)
        return SynodicBackendResponse(
            hits=hits,
            points=np.asarray(pts, dtype=float) if pts else np.empty((0, 2)),
            states=np.asarray(sts, dtype=float) if sts else np.empty((0, 6)),
            times=np.asarray(ts, dtype=float) if ts else None,
            trajectory_indices=np.asarray(tis, dtype=int) if tis else np.empty((0,), dtype=int),
            metadata={},
        )

    run.__wrapped__ = orig
    return run


_TABLES = {}


def _poly_table(jac_H, clmo):
    """Sparse term table of a reference Hamiltonian (cached per jac_H / clmo object pair)."""
    key = (id(jac_H), id(clmo))
    if key not in _TABLES:
        if len(_TABLES) > 16:
            _TABLES.clear()
        _TABLES[key] = (_cm.PolyTable.from_reference(jac_H, clmo), jac_H)
    return _TABLES[key][0]


def recognise_hamiltonian(system):
    """-> (base _HamiltonianSystem, fwd) for a 3-dof polynomial Hamiltonian system (possibly wrapped by
    _DirectedSystem), else None."""
    from hiten.algorithms.dynamics.base import _DirectedSystem
    from hiten.algorithms.dynamics.hamiltonian import _HamiltonianSystem
    base, fwd = system, 1
    if isinstance(system, _DirectedSystem):
        base, fwd = system._base, int(system._fwd)
    if type(base) is _HamiltonianSystem and base.n_dof == 3:
        return base, fwd
    return None


def _make_symplectic_integrate(orig):
    """_ExtendedSymplectic.integrate (algorithms/integrators/symplectic.py:877-1004) for the reference's polynomial
    Hamiltonian systems: grid integration (hb_ham_symplectic_dense) and recognised plane events
    (hb_ham_symplectic_event).  `_propagate_dynsys(method="symplectic")` follows because it builds this class
    (dynamics/base.py:436-444) -- including its double application of the direction sign to the returned times."""
    def integrate(self, system, y0, t_vals, *, event_fn=None, event_cfg=None, event_options=None, **kwargs):
        from hiten.algorithms.integrators.types import _Solution
        rec = recognise_hamiltonian(system)
        ev = recognise_event(event_fn) if event_fn is not None else None
        if rec is None or (event_fn is not None and ev is None) or not 2 <= int(self._order) <= 8:
            return orig(self, system, y0, t_vals, event_fn=event_fn, event_cfg=event_cfg,
                        event_options=event_options, **kwargs)
        self.validate_inputs(system, y0, t_vals)
        t_vals = np.asarray(t_vals, dtype=np.float64)
        if not np.all(np.diff(t_vals) != 0.0):
            return orig(self, system, y0, t_vals, event_fn=event_fn, event_cfg=event_cfg,
                        event_options=event_options, **kwargs)
        base, fwd = rec
        y0 = np.asarray(y0, dtype=np.float64)
        table = _poly_table(base.jac_H, base.clmo_H)
        t_int = t_vals if fwd == 1 else t_vals * (-1.0)                      # symplectic.py:963
        if ev is None:
            traj = _symp.integrate_symplectic(table, y0[None, :], t_int, self._order,
                                              c_omega_heuristic=self.c_omega_heuristic, arith=_STATE["arith"])[0]
            return _Solution(times=t_vals.copy() * fwd, states=traj)
        event = (ev[0], ev[1], 0 if event_cfg is None else int(event_cfg.direction),
                 float(event_options.xtol if event_options is not None else 1.0e-12),
                 float(event_options.gtol if event_options is not None else 1.0e-12))
        r = _symp.integrate_symplectic_until_event(table, y0[None, :], t_int, self._order, event,
                                                   c_omega_heuristic=self.c_omega_heuristic, arith=_STATE["arith"],
                                                   want_trajectory=True)
        if bool(r.hit[0]):
            return _Solution(times=np.array([t_vals[0], float(r.t_hit[0]) * fwd], dtype=np.float64),
                             states=np.vstack([y0, r.y_hit[0]]))
        return _Solution(times=t_vals.copy() * fwd, states=r.traj[0])

    integrate.__wrapped__ = orig
    integrate.__doc__ = orig.__doc__
    return integrate


def _make_cm_run(orig):
    def run(self, request):
        from hiten.algorithms.poincare.centermanifold.types import CenterManifoldBackendResponse
        seeds = np.asarray(request.seeds)
        if seeds.size == 0:
            return CenterManifoldBackendResponse(states=np.empty((0, 4)), times=np.empty((0,)),
                                                 flags=np.empty((0,), dtype=np.int64), metadata={})
        if request.method == "adaptive":
            raise NotImplementedError("Adaptive integrator is not implemented in CM backend; use 'fixed' (RK) or 'symplectic'.")
        table = _poly_table(request.jac_H, request.clmo_table)
        opts = _cm.make_opts(request.dt, request.max_steps, "symplectic" if request.method == "symplectic" else "fixed",
                             request.order, request.section_coord, request.c_omega_heuristic, _STATE["arith"])
        flags, out, tt = _cm.poincare_map(table, np.ascontiguousarray(seeds, dtype=np.float64), opts)
        ok = flags.astype(bool)                                             # failed seeds are dropped, backend.py:455-459
        return CenterManifoldBackendResponse(states=np.asarray(out[ok], dtype=np.float64).reshape(-1, 4),
                                             times=np.asarray(tt[ok], dtype=np.float64),
                                             flags=np.asarray(flags, dtype=np.int64), metadata={})

    run.__wrapped__ = orig
    return run


# ------------------------------------------------------------------------------------------------
# centre-manifold seeding: the per-candidate Brent solves of the seeding strategies and of the engine's
# lifting loop become ONE hb_cm_lift batch per map computation
# ------------------------------------------------------------------------------------------------
import threading as _threading

_LIFT_CACHE = {}                 # (id(H_blocks), section, h0, p0, p1) -> (q2, p2, q3, p3) or None, filled by a generate()
_TLS = _threading.local()        # n_seeds of the options of the compute() call running on this thread


def _h_table(H_blocks, clmo_table):
    key = ("H", id(H_blocks), id(clmo_table))
    if key not in _TABLES:
        if len(_TABLES) > 16:
            _TABLES.clear()
        _TABLES[key] = (_cm.PolyTable.from_hamiltonian(H_blocks, clmo_table), H_blocks)
    return _TABLES[key][0]


def _make_strategy_generate(orig):
    """`<seeding strategy>.generate` (algorithms/poincare/centermanifold/strategies.py:67-520).  Every strategy walks
    its candidate plane points and asks `_build_seed` (seeding.py:121-178) whether a point can be lifted onto the
    energy surface -- one Python Brent solve each -- and some stop as soon as enough valid seeds are found.  Here the
    strategy's own code runs twice: a PROBE pass in which `_build_seed` only records the candidate (and says "invalid",
    so no early exit hides later candidates), ONE hb_cm_lift batch over all candidates, then a REPLAY pass -- same
    numpy random state -- in which `_build_seed` answers from the batch.  Same seeds, same order, same early exits; the
    lifted states are kept for the engine's lifting loop (`lift_plane_point`)."""
    def generate(self, *, h0, H_blocks, clmo_table, solve_missing_coord_fn, find_turning_fn):
        section = self.config.section_coord
        turning = {}

        def turning_fn(name):                                   # two more Brent solves per pass otherwise
            if name not in turning:
                turning[name] = find_turning_fn(name)
            return turning[name]

        cands = []
        rng_state = np.random.get_state()
        # _RandomSeeding draws from an unseeded np.random.default_rng(): both passes get the same fresh seed
        entropy = np.random.SeedSequence().entropy
        real_default_rng = np.random.default_rng
        np.random.default_rng = lambda *a, **k: real_default_rng(*a, **k) if (a or k) else real_default_rng(entropy)
        self._build_seed = lambda plane_vals, *, solve_missing_coord_fn: (cands.append(tuple(float(v) for v in plane_vals)),
                                                                          None)[1]
        try:
            orig(self, h0=h0, H_blocks=H_blocks, clmo_table=clmo_table, solve_missing_coord_fn=solve_missing_coord_fn,
                 find_turning_fn=turning_fn)
        except BaseException:
            np.random.default_rng = real_default_rng
            raise
        finally:
            del self._build_seed
        np.random.set_state(rng_state)
        if not cands:
            np.random.default_rng = real_default_rng
            return orig(self, h0=h0, H_blocks=H_blocks, clmo_table=clmo_table,
                        solve_missing_coord_fn=solve_missing_coord_fn, find_turning_fn=turning_fn)
        ok, states = _cm.lift_plane_points(_h_table(H_blocks, clmo_table), section, np.asarray(cands, dtype=np.float64),
                                           float(h0))
        if len(_LIFT_CACHE) > 4_000_000:
            _LIFT_CACHE.clear()
        for c, good, st in zip(cands, ok, states):
            _LIFT_CACHE[(id(H_blocks), section, float(h0), c[0], c[1])] = tuple(float(v) for v in st) if good else None
        answers = iter(zip(cands, ok))

        def build(plane_vals, *, solve_missing_coord_fn):
            c, good = next(answers)
            if c != tuple(float(v) for v in plane_vals):          # the strategy did not replay the same candidates
                return type(self)._build_seed(self, plane_vals, solve_missing_coord_fn=solve_missing_coord_fn)
            return plane_vals if good else None

        self._build_seed = build
        try:
            seeds = orig(self, h0=h0, H_blocks=H_blocks, clmo_table=clmo_table,
                         solve_missing_coord_fn=solve_missing_coord_fn, find_turning_fn=turning_fn)
        finally:
            del self._build_seed
            np.random.default_rng = real_default_rng
        # The engine lifts the returned seeds on the section of the compute() call (engine.py:139-149), which is NOT the
        # strategy's own config.section_coord when the caller passes section_coord at run time (the strategy keeps the
        # config it was built with): one more batch on that section, so that the engine's loop is answered from it too.
        sec_run = getattr(_TLS, "section", None)
        if sec_run is not None and sec_run != section and seeds:
            pts = np.asarray([(float(p_[0]), float(p_[1])) for p_ in seeds], dtype=np.float64)
            ok2, st2 = _cm.lift_plane_points(_h_table(H_blocks, clmo_table), sec_run, pts, float(h0))
            for c, good, st in zip(pts.tolist(), ok2, st2):
                _LIFT_CACHE[(id(H_blocks), sec_run, float(h0), c[0], c[1])] = tuple(float(v) for v in st) if good else None
        return seeds

    generate.__wrapped__ = orig
    return generate


def _make_lift_plane_point(orig):
    """_CenterManifoldInterface.lift_plane_point (interfaces.py:297-337): answered from the batch the seeding strategy
    just lifted (bit-identical: hb_cm_lift reproduces the reference's bracket expansion and Brent iteration); any other
    point, or non-default solver settings, goes through the reference's own method."""
    def lift_plane_point(self, plane, *, section_coord, h0, H_blocks, clmo_table, **kw):
        if not kw:
            key = (id(H_blocks), section_coord, float(h0), float(plane[0]), float(plane[1]))
            if key in _LIFT_CACHE:
                return _LIFT_CACHE[key]
        return orig(self, plane, section_coord=section_coord, h0=h0, H_blocks=H_blocks, clmo_table=clmo_table, **kw)

    lift_plane_point.__wrapped__ = orig
    return lift_plane_point


def _make_create_problem(orig):
    """Remember the n_seeds of the options this compute() call was given: the reference's strategies read
    `getattr(map_config, "n_seeds", 20)` (core/strategies.py:102-121) and the config has no such field, so
    SeedingOptions(n_seeds=...) never reaches them.  With install(cm_seeds_from_options=True) it does."""
    def create_problem(self, *, domain_obj, config, options):
        n = None
        try:
            n = int(options.seeding.n_seeds)
        except Exception:
            pass
        _TLS.n_seeds = n
        _TLS.section = getattr(config, "section_coord", None)      # the section the engine will lift the seeds on
        return orig(self, domain_obj=domain_obj, config=config, options=options)

    create_problem.__wrapped__ = orig
    return create_problem


def _make_n_seeds(orig_prop):
    def n_seeds(self):
        n = getattr(_TLS, "n_seeds", None)
        if _STATE.get("cm_seeds_from_options") and n is not None:
            return n
        return orig_prop.fget(self)
    return property(n_seeds)


def _make_connections_run(orig):
    def run(self, request):
        """_ConnectionsBackend.run (algorithms/connections/backends.py:425-540) through hb_connections."""
        from hiten.algorithms.connections.types import ConnectionsBackendResponse, _ConnectionResult
        pu, ps = np.asarray(request.points_u, dtype=np.float64), np.asarray(request.points_s, dtype=np.float64)
        if pu.size == 0 or ps.size == 0:
            return ConnectionsBackendResponse(results=[], metadata={})
        c = _conn.find_connections(pu, ps, request.states_u, request.states_s, float(request.eps), float(request.dv_tol),
                                   float(request.bal_tol), traj_indices_u=request.traj_indices_u,
                                   traj_indices_s=request.traj_indices_s)
        if c.pairs_considered == 0:
            return ConnectionsBackendResponse(results=[], metadata={})
        results = [_ConnectionResult(kind=_conn.KINDS[int(c.kind[k])], delta_v=float(c.delta_v[k]),
                                     point2d=(float(c.point2d[k, 0]), float(c.point2d[k, 1])),
                                     state_u=c.state_u[k].copy(), state_s=c.state_s[k].copy(),
                                     index_u=int(c.index_u[k]), index_s=int(c.index_s[k]),
                                     trajectory_index_u=int(c.trajectory_index_u[k]),
                                     trajectory_index_s=int(c.trajectory_index_s[k])) for k in range(len(c.delta_v))]
        metadata = dict(request.metadata)
        metadata.update({"pairs_considered": c.pairs_considered, "accepted": len(results)})
        return ConnectionsBackendResponse(results=results, metadata=metadata)

    run.__wrapped__ = orig
    return run


def _make_orbit_correct(orig):
    """_OrbitCorrectionService.correct (algorithms/types/services/orbits.py:114-151) as ONE hb_correct_orbits call
    (batch of one): the whole Newton + Armijo loop runs on the GPU.  Same cache key, same `apply_correction`, same
    (state, period, result) return; non-convergence raises ConvergenceError like the reference's backend.  Only with
    install(corrector="batched"): the result agrees with the reference's own Newton loop to <= 1e-10 (iteration count
    within one), not bit for bit."""
    def correct(self, *, options=None):
        from . import corrector as _corr
        if options is None:
            options = self.correction_options
        orbit = self.domain_obj
        try:
            op = _corr.opts_from_reference(orbit, options)
        except Exception:
            op = None
        if op is None:
            return orig(self, options=options)
        from hiten.algorithms.corrector.types import OrbitCorrectionDomainPayload, OrbitCorrectionResult
        from hiten.algorithms.types.exceptions import ConvergenceError
        cache_key = self.make_key("correct", tuple(sorted(options.to_dict().items())))

        def _factory():
            x0 = np.asarray(orbit.initial_state, dtype=np.float64)[None, :]
            res = _corr.correct_orbits(x0, float(orbit.mu), op, integ=_integ())
            if int(res.status[0]) != 0:
                raise ConvergenceError(f"Newton did not converge ({_corr.STATUS[int(res.status[0])]}, "
                                       f"|R|={float(res.residual_norm[0]):.2e}).")
            result = OrbitCorrectionResult(converged=True, x_corrected=np.asarray(res.x_corrected[0], dtype=float),
                                           residual_norm=float(res.residual_norm[0]),
                                           iterations=int(res.iterations[0]), half_period=float(res.half_period[0]))
            payload = OrbitCorrectionDomainPayload._from_mapping({
                "x_full": result.x_corrected, "half_period": result.half_period, "iterations": result.iterations,
                "residual_norm": result.residual_norm})
            self.apply_correction(payload)
            return result.x_corrected, 2 * result.half_period, payload, result

        state, period, payload, result = self.get_or_create(cache_key, _factory)
        return state, period, result

    correct.__wrapped__ = orig
    return correct


# ------------------------------------------------------------------------------------------------
# install / uninstall
# ------------------------------------------------------------------------------------------------
def install(arith="parity", corrector="reference", cm_seeds_from_options=False):
    """Rebind the reference's funnels.  Requires `hiten` to be importable; idempotent.
    corrector="reference" keeps the reference's own Newton loop (its propagations run on the GPU through the rebound
    `_propagate_dynsys` / `_DOP853.integrate`); corrector="batched" additionally rebinds
    `_OrbitCorrectionService.correct` to hb_correct_orbits (one GPU call per correct()).
    cm_seeds_from_options=True lets `SeedingOptions(n_seeds=...)` reach the centre-manifold seeding strategies (the
    reference always seeds 20, core/strategies.py:102-121), which is what makes BASELINE configs[2]'s 1e5 seeds
    reachable from cm.poincare_map().compute(); False keeps the reference's behaviour."""
    if corrector not in ("reference", "batched"):
        raise ValueError("corrector must be 'reference' or 'batched'")
    if _STATE["installed"]:
        _STATE["arith"] = arith
        _STATE["cm_seeds_from_options"] = bool(cm_seeds_from_options)
        want_batched = corrector == "batched"
        if want_batched != ("orbit_correct" in _STATE["orig"]):           # the corrector choice changed: re-bind it
            from hiten.algorithms.types.services.orbits import _OrbitCorrectionService
            if want_batched:
                _STATE["orig"]["orbit_correct"] = _OrbitCorrectionService.correct
                _OrbitCorrectionService.correct = _make_orbit_correct(_STATE["orig"]["orbit_correct"])
            else:
                _OrbitCorrectionService.correct = _STATE["orig"].pop("orbit_correct")
        return
    import hiten  # noqa: F401
    import hiten.algorithms.dynamics.base as dbase
    from hiten.algorithms.connections.backends import _ConnectionsBackend
    from hiten.algorithms.integrators.rk import _DOP853, _RK45, _FixedStepRK
    from hiten.algorithms.integrators.symplectic import _ExtendedSymplectic
    from hiten.algorithms.poincare.centermanifold.backend import _CenterManifoldBackend
    from hiten.algorithms.poincare.synodic.backend import _SynodicDetectionBackend
    from hiten.algorithms.types.services.manifold import _ManifoldDynamicsService

    _STATE["arith"] = arith
    orig_prop = dbase._propagate_dynsys
    new_prop = _make_propagate_dynsys(orig_prop)
    patched = []
    for name, mod in list(sys.modules.items()):
        if name.startswith("hiten") and mod is not None and getattr(mod, "_propagate_dynsys", None) is orig_prop:
            setattr(mod, "_propagate_dynsys", new_prop)
            patched.append(mod)
    _STATE["patched_modules"] = patched
    _STATE["orig"] = {
        "propagate": orig_prop,
        "dop853": _DOP853.integrate,
        "rk45": _RK45.integrate,
        "fixed_rk": _FixedStepRK.integrate,
        "symplectic": _ExtendedSymplectic.integrate,
        "run_compute": _ManifoldDynamicsService._run_compute,
        "synodic": _SynodicDetectionBackend.run,
        "cm": _CenterManifoldBackend.run,
        "connections": _ConnectionsBackend.run,
    }
    _DOP853.integrate = _make_dop853_integrate(_STATE["orig"]["dop853"])
    _RK45.integrate = _make_rk_integrate(_STATE["orig"]["rk45"], "rk45")
    _FixedStepRK.integrate = _make_rk_integrate(_STATE["orig"]["fixed_rk"], "fixed")
    _ExtendedSymplectic.integrate = _make_symplectic_integrate(_STATE["orig"]["symplectic"])
    _ManifoldDynamicsService._run_compute = _make_run_compute(_STATE["orig"]["run_compute"])
    _SynodicDetectionBackend.run = _make_synodic_run(_STATE["orig"]["synodic"])
    _CenterManifoldBackend.run = _make_cm_run(_STATE["orig"]["cm"])
    _ConnectionsBackend.run = _make_connections_run(_STATE["orig"]["connections"])
    # centre-manifold seeding: batched lifting behind the strategies and the engine's lifting loop
    import hiten.algorithms.poincare.centermanifold.strategies as cm_strat
    from hiten.algorithms.poincare.centermanifold.interfaces import _CenterManifoldInterface
    from hiten.algorithms.poincare.core.strategies import _SeedingStrategyBase
    _STATE["cm_seeds_from_options"] = bool(cm_seeds_from_options)
    strat_classes = [c for c in vars(cm_strat).values()
                     if isinstance(c, type) and c.__module__ == cm_strat.__name__ and "generate" in vars(c)]
    _STATE["orig"]["cm_generate"] = {c: c.generate for c in strat_classes}
    for c in strat_classes:
        c.generate = _make_strategy_generate(c.generate)
    _STATE["orig"]["cm_lift"] = _CenterManifoldInterface.lift_plane_point
    _STATE["orig"]["cm_create_problem"] = _CenterManifoldInterface.create_problem
    _STATE["orig"]["cm_n_seeds"] = _SeedingStrategyBase.__dict__["n_seeds"]
    _CenterManifoldInterface.lift_plane_point = _make_lift_plane_point(_STATE["orig"]["cm_lift"])
    _CenterManifoldInterface.create_problem = _make_create_problem(_STATE["orig"]["cm_create_problem"])
    _SeedingStrategyBase.n_seeds = _make_n_seeds(_STATE["orig"]["cm_n_seeds"])
    if corrector == "batched":
        from hiten.algorithms.types.services.orbits import _OrbitCorrectionService
        _STATE["orig"]["orbit_correct"] = _OrbitCorrectionService.correct
        _OrbitCorrectionService.correct = _make_orbit_correct(_STATE["orig"]["orbit_correct"])
    _STATE["installed"] = True


def uninstall():
    if not _STATE["installed"]:
        return
    from hiten.algorithms.integrators.rk import _DOP853
    from hiten.algorithms.poincare.centermanifold.backend import _CenterManifoldBackend
    from hiten.algorithms.poincare.synodic.backend import _SynodicDetectionBackend
    from hiten.algorithms.types.services.manifold import _ManifoldDynamicsService
    o = _STATE["orig"]
    for mod in _STATE["patched_modules"]:
        setattr(mod, "_propagate_dynsys", o["propagate"])
    _DOP853.integrate = o["dop853"]
    from hiten.algorithms.integrators.rk import _RK45, _FixedStepRK
    _RK45.integrate = o["rk45"]
    _FixedStepRK.integrate = o["fixed_rk"]
    from hiten.algorithms.integrators.symplectic import _ExtendedSymplectic
    _ExtendedSymplectic.integrate = o["symplectic"]
    _ManifoldDynamicsService._run_compute = o["run_compute"]
    _SynodicDetectionBackend.run = o["synodic"]
    _CenterManifoldBackend.run = o["cm"]
    from hiten.algorithms.connections.backends import _ConnectionsBackend
    _ConnectionsBackend.run = o["connections"]
    if "orbit_correct" in o:
        from hiten.algorithms.types.services.orbits import _OrbitCorrectionService
        _OrbitCorrectionService.correct = o["orbit_correct"]
    if "cm_generate" in o:
        from hiten.algorithms.poincare.centermanifold.interfaces import _CenterManifoldInterface
        from hiten.algorithms.poincare.core.strategies import _SeedingStrategyBase
        for c, fn in o["cm_generate"].items():
            c.generate = fn
        _CenterManifoldInterface.lift_plane_point = o["cm_lift"]
        _CenterManifoldInterface.create_problem = o["cm_create_problem"]
        _SeedingStrategyBase.n_seeds = o["cm_n_seeds"]
    _STATE.update(installed=False, orig={}, patched_modules=[])
    _TABLES.clear()
    _LIFT_CACHE.clear()
    _TLS.n_seeds = None


def is_installed():
    return _STATE["installed"]
