"""Manifold-tube host logic: initial conditions and the batched replacement of the fraction loop.

Reference: hiten/algorithms/types/services/manifold.py
  _compute_manifold_section  :470-537   (IC = orbit point + displacement along Phi*eigvec)
  _run_compute               :293-442   (serial `for fraction in ...` loop -> one batched GPU call here)
"""
import numpy as np


def manifold_initial_conditions(x_node, man, displacement):
    """x0W for every (node, displacement) pair.

    x_node[K,6]: states on the orbit; man[K,6]: direction * Phi @ eigvec (real); displacement: scalar or
    array[D].  Returns [K*D, 6] ordered displacement-major (all nodes for displacement 0, then 1, ...).
    Mirrors manifold.py:515-535: d = displacement / |man[0:3]| (1.0 if that norm < 1e-14),
    x0W = x + d * man, then z and vz are zeroed when |.| < 1e-15.
    """
    x_node = np.asarray(x_node, dtype=np.float64)
    man = np.asarray(man, dtype=np.float64)
    disp = np.atleast_1d(np.asarray(displacement, dtype=np.float64))
    mag = np.array([np.linalg.norm(m[0:3]) for m in man])
    mag = np.where(mag < 1e-14, 1.0, mag)
    out = np.empty((disp.size, x_node.shape[0], 6))
    for j, dj in enumerate(disp):
        d = dj / mag
        out[j] = x_node + d[:, None] * man
    out = out.reshape(-1, 6)
    out[np.abs(out[:, 2]) < 1.0e-15, 2] = 0.0
    out[np.abs(out[:, 5]) < 1.0e-15, 5] = 0.0
    return out


# ----------------------------------------------------------------------------------------------------------------
# device path: initial conditions from the dense STM and the trajectory filters, without leaving HBM
# ----------------------------------------------------------------------------------------------------------------
def _dev_f64(a, device):
    import torch
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(device)


def tube_initial_conditions(phi_dense, tt, period, eigvec, direction, fractions, displacements, *, device=None,
                            stream=None):
    """hb_manifold_ics: the tube's initial conditions as the SoA device array [6, D*K] the propagation entry points
    read, and the STM sample index of every fraction (int32 device tensor [K]).

    phi_dense[S,42] / tt[S]: the reference's PHI / times of `_compute_stm` (device tensors are used in place, e.g.
    `cr3bp_stm_dense(..., keep_on_device=True).states[0]`); eigvec: the (real-valued) eigenvector of
    manifold.py:357-370; direction +1/-1; trajectory i = j*K + k is (fraction k, displacement j).
    Mirrors `_compute_manifold_section` + `_totime` (manifold.py:470-573) bit for bit.
    """
    import torch
    from . import _lib as L
    from .propagate import _require_cuda, _stream_ptr
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ev = np.asarray(eigvec)
    if np.iscomplexobj(ev):
        if np.any(ev.imag != 0.0):
            raise ValueError("eigvec must be real-valued (get_real_eigenvectors output)")
        ev = ev.real
    if ev.shape != (6,):
        raise ValueError("eigvec must have 6 components")
    if int(direction) not in (1, -1):
        raise ValueError("direction must be +1 or -1")
    with torch.cuda.device(device):
        phi = _dev_f64(phi_dense, device)
        if phi.dim() != 2 or phi.shape[1] != 42:
            raise ValueError("phi_dense must be [S, 42]")
        t = _dev_f64(tt, device)
        if t.numel() != phi.shape[0]:
            raise ValueError("tt and phi_dense disagree on the number of samples")
        fr = _dev_f64(np.atleast_1d(fractions) if not isinstance(fractions, torch.Tensor) else fractions, device)
        dd = _dev_f64(np.atleast_1d(displacements) if not isinstance(displacements, torch.Tensor) else displacements,
                      device)
        evd = _dev_f64(ev, device)
        K, D = fr.numel(), dd.numel()
        x0 = torch.empty((6, K * D), dtype=torch.float64, device=device)
        idx = torch.empty(K, dtype=torch.int32, device=device)
        rc = lib.hb_manifold_ics(phi.data_ptr(), t.data_ptr(), int(t.numel()), float(period), evd.data_ptr(),
                                 int(direction), fr.data_ptr(), K, dd.data_ptr(), D, x0.data_ptr(), idx.data_ptr(),
                                 _stream_ptr(stream))
        L.check(rc, "hb_manifold_ics")
    return x0, idx


def tube_filter(states, mu, *, safe_r1=0.0, safe_r2=0.0, energy_tol=np.inf, device=None, stream=None):
    """hb_tube_filter on stored tubes states[N, m, 6] (device tensor used in place): returns
    (quantities[N,3] = min r1, min r2, max relative Jacobi drift; keep[N] int32) as device tensors.
    Mirrors manifold.py:412-432 + energy.py:27-76 bit for bit."""
    import torch
    from . import _lib as L
    from .propagate import _require_cuda, _stream_ptr
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        s = _dev_f64(states, device)
        if s.dim() != 3 or s.shape[2] != 6 or s.shape[1] < 1:
            raise ValueError("states must be [N, m, 6] with m >= 1")
        n, m = int(s.shape[0]), int(s.shape[1])
        out = torch.empty((n, 3), dtype=torch.float64, device=device)
        keep = torch.empty(n, dtype=torch.int32, device=device)
        opts = L.HbTubeFilterOpts(float(mu), float(safe_r1), float(safe_r2), float(energy_tol))
        rc = lib.hb_tube_filter(opts, n, s.data_ptr(), m, out.data_ptr(), keep.data_ptr(), _stream_ptr(stream))
        L.check(rc, "hb_tube_filter")
    return out, keep
