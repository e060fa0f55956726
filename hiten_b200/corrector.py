"""Batched single-shooting differential correction.

Host mirror of the reference's corrector for periodic orbits -- `PeriodicOrbit.correct()` ->
`_OrbitCorrectionService.correct` (hiten/algorithms/types/services/orbits.py:114-151) -> `_NewtonBackend.run`
(hiten/algorithms/corrector/backends/newton.py) with the Armijo stepper -- for MANY orbits at once: every event
propagation, STM propagation, 2x2 solve and line-search decision of the batch runs on the GPU in lock-step
(`hb_correct_orbits`, csrc/hb_corrector.cu).  No CPU fallback.
"""
from dataclasses import dataclass

import numpy as np

from . import _lib as L

# family -> (control_indices, residual_indices, event coordinate, halo quadratic term, finite differences):
# the shipped OrbitCorrectionConfig of each orbit family (services/orbits.py:862-880, 1249-1262, 1440-1453)
FAMILIES = {
    "halo": ((0, 4), (3, 5), 1, True, False),
    "lyapunov": ((4, 5), (3, 2), 1, False, False),
    "vertical": ((5, 4), (3, 1), 2, False, True),
}

STATUS = {0: "converged", 1: "max_attempts", 2: "step_failed", 3: "no_event", 4: "singular"}


def make_opts(family=None, *, control_indices=None, residual_indices=None, target=(0.0, 0.0), event_idx=None,
              event_offset=0.0, halo_quadratic=None, finite_difference=None, line_search=True, tol=1e-12,
              max_attempts=50, max_delta=1e-2, fd_step=1e-8, alpha_reduction=0.5, min_alpha=1e-4, armijo_c=0.1):
    """hb_correct_opts from a family name and / or explicit OrbitCorrectionConfig / OrbitCorrectionOptions fields
    (defaults = the reference's: ConvergenceOptions(50, 1e-12, 1e-2), NumericalOptions(1e-8, 0.5, 1e-4, 0.1))."""
    if family is not None:
        ctrl, res, ev, quad, fd = FAMILIES[family]
        control_indices = ctrl if control_indices is None else control_indices
        residual_indices = res if residual_indices is None else residual_indices
        event_idx = ev if event_idx is None else event_idx
        halo_quadratic = quad if halo_quadratic is None else halo_quadratic
        finite_difference = fd if finite_difference is None else finite_difference
    if control_indices is None or residual_indices is None or event_idx is None:
        raise ValueError("give a family or control_indices, residual_indices and event_idx")
    ctrl, res = tuple(int(i) for i in control_indices), tuple(int(i) for i in residual_indices)
    if len(ctrl) != 2 or len(res) != 2 or len(tuple(target)) != 2:
        raise ValueError("the batched corrector handles two controls and two residuals (every shipped family)")
    md = float("inf") if max_delta is None else float(max_delta)
    return L.HbCorrectOpts((L.C.c_int32 * 2)(*ctrl), (L.C.c_int32 * 2)(*res), (L.C.c_double * 2)(*map(float, target)),
                           int(event_idx), int(bool(halo_quadratic)), float(event_offset), int(bool(finite_difference)),
                           int(bool(line_search)), float(tol), md, float(fd_step), float(alpha_reduction),
                           float(min_alpha), float(armijo_c), int(max_attempts), 0)


@dataclass
class CorrectionBatch:
    x_corrected: object        # [N, 6] numpy (or [6, N] device tensor with keep_on_device)
    half_period: object        # [N]; NaN unless converged
    iterations: object         # [N] int32
    residual_norm: object      # [N] inf-norm of the last residual
    status: object             # [N] int32, see STATUS
    rk_steps6: int             # attempted 6-state DOP853 steps of the whole call
    rk_steps42: int            # attempted 42-state DOP853 steps of the whole call

    @property
    def converged(self):
        return self.status == 0


def correct_orbits(x0, mu, opts, *, integ=None, device=None, stream=None, keep_on_device=False, scratch=None):
    """Correct N initial guesses x0[N, 6] (or a device SoA tensor [6, N]) in one lock-step batch."""
    import torch
    from .propagate import _require_cuda, _stream_ptr, _to_device_soa, make_integ, make_sys, workspace
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        x0d, host = _to_device_soa(x0, device)
        n = int(x0d.shape[1])
        f64, i32 = dict(dtype=torch.float64, device=device), dict(dtype=torch.int32, device=device)
        xc, half, rn = torch.empty((6, n), **f64), torch.empty(n, **f64), torch.empty(n, **f64)
        it, st = torch.empty(n, **i32), torch.empty(n, **i32)
        need = int(lib.hb_correct_scratch_bytes(n))
        if scratch is None or scratch.numel() < need:
            scratch = torch.empty(need, dtype=torch.uint8, device=device)
        ws = workspace(device)
        s6, s42 = L.C.c_int64(0), L.C.c_int64(0)
        integ = make_integ() if integ is None else integ
        rc = lib.hb_correct_orbits(make_sys(mu, 1, None), integ, opts, n, x0d.data_ptr(), xc.data_ptr(),
                                   half.data_ptr(), it.data_ptr(), rn.data_ptr(), st.data_ptr(), L.C.byref(s6),
                                   L.C.byref(s42), scratch.data_ptr(), need, ws.data_ptr(), _stream_ptr(stream))
        L.check(rc, "hb_correct_orbits")
        if host and not keep_on_device:
            return CorrectionBatch(xc.t().contiguous().cpu().numpy(), half.cpu().numpy(), it.cpu().numpy(),
                                   rn.cpu().numpy(), st.cpu().numpy(), s6.value, s42.value)
        return CorrectionBatch(xc, half, it, rn, st, s6.value, s42.value)


def opts_from_reference(orbit, options=None):
    """hb_correct_opts from a reference PeriodicOrbit's own correction_config / correction_options (or the given
    OrbitCorrectionOptions), or None when the configuration is outside this path (multiple shooting, != 2 controls,
    a non-plane event, another integration method)."""
    from hiten.algorithms.poincare.singlehit import backend as sh
    cfg, opt = orbit.correction_config, (orbit.correction_options if options is None else options)
    if not hasattr(opt, "base") or not hasattr(opt, "forward"):
        return None
    ev = {sh._x_plane_crossing: 0, sh._y_plane_crossing: 1, sh._z_plane_crossing: 2}.get(getattr(cfg, "event_func", None))
    if ev is None or type(cfg).__name__ != "OrbitCorrectionConfig":      # not the multiple-shooting subclass
        return None
    if len(cfg.control_indices) != 2 or len(cfg.residual_indices) != 2:
        return None
    if cfg.integration.method != "adaptive" or opt.base.integration.order != 8 or int(opt.forward) != 1:
        return None
    quad = cfg.extra_jacobian is not None
    if quad and getattr(cfg.extra_jacobian, "__name__", "") != "_halo_quadratic_term":
        return None
    conv, num = opt.base.convergence, opt.base.numerical
    return make_opts(control_indices=[int(i) for i in cfg.control_indices],
                     residual_indices=[int(i) for i in cfg.residual_indices], target=tuple(cfg.target), event_idx=ev,
                     halo_quadratic=quad, finite_difference=cfg.numerical.finite_difference,
                     line_search=cfg.numerical.line_search_enabled, tol=conv.tol, max_attempts=conv.max_attempts,
                     max_delta=conv.max_delta, fd_step=num.fd_step, alpha_reduction=num.line_search_alpha_reduction,
                     min_alpha=num.line_search_min_alpha, armijo_c=num.line_search_armijo_c)


def correct_many(orbits):
    """Correct a list of reference PeriodicOrbit objects in one GPU batch per configuration and apply the results
    exactly as `_OrbitCorrectionService.correct` does (services/orbits.py:137-164).  Orbits whose configuration is
    outside this path go through their own `correct()`; failures raise the reference's ConvergenceError after the
    successful members have been updated.  Returns the list of OrbitCorrectionResult."""
    from hiten.algorithms.corrector.types import OrbitCorrectionDomainPayload, OrbitCorrectionResult
    from hiten.algorithms.types.exceptions import ConvergenceError
    results = [None] * len(orbits)
    groups = {}
    for k, o in enumerate(orbits):
        op = opts_from_reference(o)
        if op is None:
            results[k] = o.correct()
            continue
        groups.setdefault((float(o.mu), bytes(op)), (op, []))[1].append(k)
    failed = []
    for (mu, _), (op, idx) in groups.items():
        x0 = np.stack([np.asarray(orbits[k].initial_state, dtype=np.float64) for k in idx])
        res = correct_orbits(x0, mu, op)
        for j, k in enumerate(idx):
            if int(res.status[j]) != 0:
                failed.append((k, STATUS[int(res.status[j])], float(res.residual_norm[j])))
                continue
            payload = OrbitCorrectionDomainPayload._from_mapping({
                "x_full": res.x_corrected[j].copy(), "half_period": float(res.half_period[j]),
                "iterations": int(res.iterations[j]), "residual_norm": float(res.residual_norm[j])})
            orbits[k]._correction.apply_correction(payload)
            results[k] = OrbitCorrectionResult(converged=True, x_corrected=payload.x_full,
                                               residual_norm=float(payload.residual_norm),
                                               iterations=int(payload.iterations), half_period=payload.half_period)
    if failed:
        k, why, rn = failed[0]
        raise ConvergenceError(f"{len(failed)} of {len(orbits)} orbits did not converge (first: orbit {k}, {why}, "
                               f"|R|={rn:.2e})")
    return results
