"""The Tao extended-phase-space integrator over a time grid on the GPU (host wrapper over hb_ham_symplectic_dense /
hb_ham_symplectic_event).

Reference: `_ExtendedSymplectic.integrate` (hiten/algorithms/integrators/symplectic.py:877-1004) =
`_integrate_symplectic` (:564-653) and `_integrate_symplectic_until_event` (:657-782); reached from
`_propagate_dynsys(method="symplectic")` (hiten/algorithms/dynamics/base.py:436-444).  A batch of initial states shares
one grid; the per-interval Tao parameters (omega, triple-jump sub-steps, cos / sin) are evaluated once on the host with
libm (`hb_tao_grid_prepare`), exactly as the reference evaluates them per step, and live in HBM as one small table.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from .propagate import _require_cuda, _stream_ptr, workspace


def tao_grid_table(t_vals_signed, order, c_omega_heuristic=20.0):
    """Host-only (no GPU): (n_sub, table[(m-1), 3, n_sub]) = sub-step lengths, cos, sin of every grid interval."""
    t = np.ascontiguousarray(t_vals_signed, dtype=np.float64)
    if t.ndim != 1 or t.size < 2:
        raise ValueError("Must provide at least 2 time points")                 # integrators/base.py:183
    if int(order) <= 0 or int(order) % 2 != 0:
        raise ValueError("Order must be a positive even integer")               # symplectic.py:832-833
    lib = L.load()
    n_sub = L.C.c_int32(0)
    L.check(lib.hb_tao_grid_prepare(t.ctypes.data, int(t.size), int(order), float(c_omega_heuristic),
                                    L.C.byref(n_sub), None, 0), "hb_tao_grid_prepare")
    tab = np.empty((t.size - 1, 3, n_sub.value), dtype=np.float64)
    L.check(lib.hb_tao_grid_prepare(t.ctypes.data, int(t.size), int(order), float(c_omega_heuristic),
                                    L.C.byref(n_sub), tab.ctypes.data, tab.size), "hb_tao_grid_prepare")
    return int(n_sub.value), tab


def _prep(table, y0, t_vals_signed, order, c_omega_heuristic, arith, device):
    host = not (isinstance(y0, torch.Tensor) and y0.is_cuda)
    yd = torch.from_numpy(np.ascontiguousarray(y0, dtype=np.float64)).to(device) if host else y0.contiguous()
    if yd.dim() != 2 or yd.shape[1] != 6:
        raise ValueError("y0 must have shape (N, 6) = [Q, P]")
    t = np.ascontiguousarray(t_vals_signed, dtype=np.float64)
    n_sub, tab = tao_grid_table(t, order, c_omega_heuristic)
    tabd = torch.from_numpy(tab).to(device)
    o = L.HbSympOpts(int(order), {"parity": L.HB_ARITH_PARITY, "fast": L.HB_ARITH_FAST}[arith], int(t.size), n_sub)
    ham, keep = table.device_struct(device)
    return host, yd, t, tabd, o, ham, keep


def jit_compile_host(table, arith="parity"):
    """Generate + compile the specialised grid kernel offline (no GPU): returns the cubin size in bytes."""
    rec = table.packed()
    ptr = (L.C.c_int64 * 7)(*[int(x) for x in table.ptr])
    nbytes = L.C.c_int64(0)
    L.check(L.load().hb_symp_jit_compile_host(rec.ctypes.data if rec.size else None, ptr,
                                              {"parity": L.HB_ARITH_PARITY, "fast": L.HB_ARITH_FAST}[arith],
                                              L.C.byref(nbytes)), "hb_symp_jit_compile_host")
    return int(nbytes.value)


def integrate_symplectic(table, y0, t_vals_signed, order, *, c_omega_heuristic=20.0, arith="parity", device=None,
                         stream=None, jit=True):
    """_integrate_symplectic for a batch: y0 [N, 6] -> traj [N, m, 6] on the signed grid (t_vals * fwd).
    Host ndarray in -> ndarray out; CUDA tensor in -> CUDA tensor out.  Bit-identical to the reference (parity).
    jit=True (default) runs the kernel specialised for this Hamiltonian (hb_ham_symplectic_jit), jit=False the
    table-driven kernel (hb_ham_symplectic_dense); both give the same bits in the parity variant."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        host, yd, t, tabd, o, ham, keep = _prep(table, y0, t_vals_signed, order, c_omega_heuristic, arith, device)
        n = int(yd.shape[0])
        traj = torch.empty((n, t.size, 6), dtype=torch.float64, device=device)
        ws = workspace(device)
        if jit:
            L.check(lib.hb_ham_symplectic_jit(ham, L.C.byref(o), None, n, yd.data_ptr(), None, tabd.data_ptr(),
                                              traj.data_ptr(), None, None, None, None, ws.data_ptr(),
                                              _stream_ptr(stream)), "hb_ham_symplectic_jit")
        else:
            L.check(lib.hb_ham_symplectic_dense(ham, L.C.byref(o), n, yd.data_ptr(), tabd.data_ptr(), traj.data_ptr(),
                                                ws.data_ptr(), _stream_ptr(stream)), "hb_ham_symplectic_dense")
        return traj.cpu().numpy() if host else traj


@dataclass
class SymplecticEventResult:
    hit: object      # [N] bool
    t_hit: object    # [N] time on the signed grid (t_vals[-1] where no event)
    y_hit: object    # [N, 6]
    n_rows: object   # [N] trajectory rows before the event (m where no event)
    traj: object     # [N, m, 6] or None; rows >= n_rows[i] are undefined


def integrate_symplectic_until_event(table, y0, t_vals_signed, order, event, *, c_omega_heuristic=20.0, arith="parity",
                                     want_trajectory=False, device=None, stream=None, jit=True):
    """_integrate_symplectic_until_event for a batch; `event` = (idx, offset, direction, xtol, gtol) of the plane event
    g = y[idx] - offset."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx, offset, direction, xtol, gtol = event
    if not 0 <= int(idx) < 6:
        raise ValueError("event index must be in 0..5")
    ev = L.HbEvent(int(idx), int(direction), float(offset), float(xtol), float(gtol))
    with torch.cuda.device(device):
        host, yd, t, tabd, o, ham, keep = _prep(table, y0, t_vals_signed, order, c_omega_heuristic, arith, device)
        n = int(yd.shape[0])
        td = torch.from_numpy(t).to(device)
        traj = torch.empty((n, t.size, 6), dtype=torch.float64, device=device) if want_trajectory else None
        hit = torch.zeros(n, dtype=torch.int32, device=device)
        n_rows = torch.zeros(n, dtype=torch.int32, device=device)
        t_hit = torch.zeros(n, dtype=torch.float64, device=device)
        y_hit = torch.zeros((n, 6), dtype=torch.float64, device=device)
        ws = workspace(device)
        fn = lib.hb_ham_symplectic_jit if jit else lib.hb_ham_symplectic_event
        L.check(fn(ham, L.C.byref(o), L.C.byref(ev), n, yd.data_ptr(), td.data_ptr(), tabd.data_ptr(),
                   traj.data_ptr() if traj is not None else None, hit.data_ptr(), t_hit.data_ptr(), y_hit.data_ptr(),
                   n_rows.data_ptr(), ws.data_ptr(), _stream_ptr(stream)),
                "hb_ham_symplectic_jit" if jit else "hb_ham_symplectic_event")
        if host:
            return SymplecticEventResult(hit.cpu().numpy().astype(bool), t_hit.cpu().numpy(), y_hit.cpu().numpy(),
                                         n_rows.cpu().numpy().astype(np.int64),
                                         traj.cpu().numpy() if traj is not None else None)
        return SymplecticEventResult(hit.bool(), t_hit, y_hit, n_rows, traj)


# ------------------------------------------------------------------------------------------------------------------
# Fixed-step RK classes on a polynomial Hamiltonian system (the `_ham` kernels, rk.py:592-656 / 722-757)
# ------------------------------------------------------------------------------------------------------------------
_RK_METHOD = {4: L.HB_RK4, 6: L.HB_RK6, 8: L.HB_RK8}


def _prep_rk(table, y0, t_vals, order, arith, device):
    host = not (isinstance(y0, torch.Tensor) and y0.is_cuda)
    yd = torch.from_numpy(np.ascontiguousarray(y0, dtype=np.float64)).to(device) if host else y0.contiguous()
    if yd.dim() != 2 or yd.shape[1] != 6:
        raise ValueError("y0 must have shape (N, 6) = [Q, P]")
    t = np.ascontiguousarray(t_vals, dtype=np.float64)
    if t.ndim != 1 or t.size < 2:
        raise ValueError("Must provide at least 2 time points")                 # integrators/base.py:183
    if int(order) not in _RK_METHOD:
        raise ValueError("RK order must be 4, 6, or 8")
    ham, keep = table.device_struct(device)
    return (host, yd, t, torch.from_numpy(t).to(device), _RK_METHOD[int(order)],
            {"parity": L.HB_ARITH_PARITY, "fast": L.HB_ARITH_FAST}[arith], ham, keep)


def integrate_rk_ham(table, y0, t_vals, order, *, arith="parity", want_derivatives=True, device=None, stream=None):
    """_FixedStepRK._integrate_fixed_rk_ham for a batch: y0 [N, 6] -> (states [N, m, 6], derivatives [N, m, 6] or None)."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        host, yd, t, td, method, ar, ham, keep = _prep_rk(table, y0, t_vals, order, arith, device)
        n = int(yd.shape[0])
        traj = torch.empty((n, t.size, 6), dtype=torch.float64, device=device)
        der = torch.empty((n, t.size, 6), dtype=torch.float64, device=device) if want_derivatives else None
        ws = workspace(device)
        L.check(lib.hb_ham_rk_dense(ham, method, ar, n, yd.data_ptr(), td.data_ptr(), int(t.size), traj.data_ptr(),
                                    der.data_ptr() if der is not None else None, ws.data_ptr(), _stream_ptr(stream)),
                "hb_ham_rk_dense")
        if host:
            return traj.cpu().numpy(), (der.cpu().numpy() if der is not None else None)
        return traj, der


def integrate_rk_ham_until_event(table, y0, t_vals, order, event, *, arith="parity", want_trajectory=False, device=None,
                                 stream=None):
    """_FixedStepRK._integrate_fixed_rk_until_event_ham for a batch; `event` = (idx, offset, direction, xtol, gtol)."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx, offset, direction, xtol, gtol = event
    if not 0 <= int(idx) < 6:
        raise ValueError("event index must be in 0..5")
    ev = L.HbEvent(int(idx), int(direction), float(offset), float(xtol), float(gtol))
    with torch.cuda.device(device):
        host, yd, t, td, method, ar, ham, keep = _prep_rk(table, y0, t_vals, order, arith, device)
        n = int(yd.shape[0])
        traj = torch.empty((n, t.size, 6), dtype=torch.float64, device=device) if want_trajectory else None
        hit = torch.zeros(n, dtype=torch.int32, device=device)
        n_rows = torch.zeros(n, dtype=torch.int32, device=device)
        t_hit = torch.zeros(n, dtype=torch.float64, device=device)
        y_hit = torch.zeros((n, 6), dtype=torch.float64, device=device)
        ws = workspace(device)
        L.check(lib.hb_ham_rk_event(ham, method, ar, L.C.byref(ev), n, yd.data_ptr(), td.data_ptr(), int(t.size),
                                    traj.data_ptr() if traj is not None else None, hit.data_ptr(), t_hit.data_ptr(),
                                    y_hit.data_ptr(), n_rows.data_ptr(), ws.data_ptr(), _stream_ptr(stream)),
                "hb_ham_rk_event")
        if host:
            return SymplecticEventResult(hit.cpu().numpy().astype(bool), t_hit.cpu().numpy(), y_hit.cpu().numpy(),
                                         n_rows.cpu().numpy().astype(np.int64),
                                         traj.cpu().numpy() if traj is not None else None)
        return SymplecticEventResult(hit.bool(), t_hit, y_hit, n_rows, traj)


# ------------------------------------------------------------------------------------------------------------------
# AdaptiveRK (DOP853 / RK45) on a polynomial Hamiltonian system (the `_ham` kernels, rk.py:2553-2676, 1403-1456, ...)
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class HamAdaptiveResult:
    states: object       # dense: [N, m, 6]; event: None
    derivatives: object  # dense: [N, m, 6] or None
    t_hit: object        # event: [N]
    y_hit: object        # event: [N, 6]
    n_acc: object
    n_rej: object
    status: object       # HB_TRAJ_* per trajectory (1 = event hit)


def _integ_struct(integ):
    from .propagate import make_integ
    return make_integ() if integ is None else integ


def integrate_adaptive_ham(table, y0, t_eval, *, integ=None, want_derivatives=True, device=None, stream=None):
    """_integrate_dop853_ham / _integrate_rk45_ham for a batch (integ.method = HB_DOP853 | HB_RK45)."""
    _require_cuda()
    lib = L.load()
    integ = _integ_struct(integ)
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        host = not (isinstance(y0, torch.Tensor) and y0.is_cuda)
        yd = torch.from_numpy(np.ascontiguousarray(y0, dtype=np.float64)).to(device) if host else y0.contiguous()
        if yd.dim() != 2 or yd.shape[1] != 6:
            raise ValueError("y0 must have shape (N, 6) = [Q, P]")
        t = np.ascontiguousarray(t_eval, dtype=np.float64)
        if t.ndim != 1 or t.size < 2:
            raise ValueError("Must provide at least 2 time points")
        td = torch.from_numpy(t).to(device)
        n = int(yd.shape[0])
        st = torch.empty((n, t.size, 6), dtype=torch.float64, device=device)
        der = torch.empty((n, t.size, 6), dtype=torch.float64, device=device) if want_derivatives else None
        na, nr, ss = (torch.zeros(n, dtype=torch.int32, device=device) for _ in range(3))
        ham, keep = table.device_struct(device)
        ws = workspace(device)
        L.check(lib.hb_ham_adaptive_dense(ham, L.C.byref(integ), n, yd.data_ptr(), td.data_ptr(), int(t.size),
                                          st.data_ptr(), der.data_ptr() if der is not None else None, na.data_ptr(),
                                          nr.data_ptr(), ss.data_ptr(), ws.data_ptr(), _stream_ptr(stream)),
                "hb_ham_adaptive_dense")
        if host:
            return HamAdaptiveResult(st.cpu().numpy(), der.cpu().numpy() if der is not None else None, None, None,
                                     na.cpu().numpy(), nr.cpu().numpy(), ss.cpu().numpy())
        return HamAdaptiveResult(st, der, None, None, na, nr, ss)


def integrate_adaptive_ham_until_event(table, y0, t0, tmax, event, *, integ=None, device=None, stream=None):
    """_integrate_dop853_until_event_ham / _integrate_rk45_until_event_ham for a batch."""
    _require_cuda()
    lib = L.load()
    integ = _integ_struct(integ)
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx, offset, direction, xtol, gtol = event
    if not 0 <= int(idx) < 6:
        raise ValueError("event index must be in 0..5")
    ev = L.HbEvent(int(idx), int(direction), float(offset), float(xtol), float(gtol))
    with torch.cuda.device(device):
        host = not (isinstance(y0, torch.Tensor) and y0.is_cuda)
        yd = torch.from_numpy(np.ascontiguousarray(y0, dtype=np.float64)).to(device) if host else y0.contiguous()
        if yd.dim() != 2 or yd.shape[1] != 6:
            raise ValueError("y0 must have shape (N, 6) = [Q, P]")
        n = int(yd.shape[0])
        th = torch.zeros(n, dtype=torch.float64, device=device)
        yh = torch.zeros((n, 6), dtype=torch.float64, device=device)
        na, nr, ss = (torch.zeros(n, dtype=torch.int32, device=device) for _ in range(3))
        ham, keep = table.device_struct(device)
        ws = workspace(device)
        L.check(lib.hb_ham_adaptive_event(ham, L.C.byref(integ), L.C.byref(ev), n, yd.data_ptr(), float(t0), float(tmax),
                                          th.data_ptr(), yh.data_ptr(), na.data_ptr(), nr.data_ptr(), ss.data_ptr(),
                                          ws.data_ptr(), _stream_ptr(stream)), "hb_ham_adaptive_event")
        if host:
            return HamAdaptiveResult(None, None, th.cpu().numpy(), yh.cpu().numpy(), na.cpu().numpy(), nr.cpu().numpy(),
                                     ss.cpu().numpy())
        return HamAdaptiveResult(None, None, th, yh, na, nr, ss)
