"""Batched CR3BP propagation on the GPU: host-side wrappers over the C ABI.

Each function takes either host numpy arrays in the reference's layout (states[N, 6]) or CUDA
torch tensors already resident in the kernels' struct-of-arrays layout ([6, N]) and returns the
same kind.  Reference call sites replaced (paths relative to hiten/):
  algorithms/dynamics/base.py:346 (_propagate_dynsys), algorithms/types/services/manifold.py:381-409
  (the serial fraction loop), algorithms/poincare/singlehit/backend.py:164-235 (event propagation).
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L


@dataclass
class BatchResult:
    """End states and per-trajectory bookkeeping of one batch call."""
    yf: "np.ndarray | torch.Tensor"          # [N, 6] numpy (host input) or [6, N] CUDA tensor (device input)
    n_acc: "np.ndarray | torch.Tensor"
    n_rej: "np.ndarray | torch.Tensor"
    status: "np.ndarray | torch.Tensor"
    t_hit: "np.ndarray | torch.Tensor | None" = None
    states: "np.ndarray | torch.Tensor | None" = None   # dense mode: [N, m, 6]


def _require_cuda():
    if not torch.cuda.is_available():
        raise L.HitenB200Error("hiten_b200 needs a CUDA device; there is no CPU fallback")


def default_min_step():
    return 10.0 * float(np.finfo(float).eps)      # rk.py:833-834


def make_sys(mu, forward=1, flip=None):
    lo, hi = (-1, -1) if flip is None else (int(flip[0]), int(flip[1]))
    return L.HbCr3bp(float(mu), 1 if forward >= 0 else -1, lo, hi, 0)


def make_integ(method=L.HB_DOP853, arith="parity", rtol=1e-12, atol=1e-12, max_step=1e4, min_step=None,
               max_attempts=0, n_fixed_steps=0, max_ctas=0):
    ar = {"parity": L.HB_ARITH_PARITY, "fast": L.HB_ARITH_FAST}[arith] if isinstance(arith, str) else int(arith)
    return L.HbInteg(int(method), ar, float(rtol), float(atol), float(max_step),
                     default_min_step() if min_step is None else float(min_step), int(max_attempts),
                     int(n_fixed_steps), int(max_ctas), None)


def with_order(integ, order):
    """Copy of `integ` whose persistent DOP853 launches hand out trajectory order[q] q-th (hb_integ.order).  `order`:
    int32 CUDA tensor holding a permutation of 0..n-1 (the returned struct holds a reference to it), or None."""
    out = L.HbInteg.from_buffer_copy(integ)
    if order is None:
        out.order = None
    else:
        if not (isinstance(order, torch.Tensor) and order.is_cuda and order.dtype == torch.int32 and order.dim() == 1
                and order.is_contiguous()):
            raise ValueError("order must be a contiguous 1-D int32 CUDA tensor")
        out.order = order.data_ptr()
        out._order_ref = order                               # the struct keeps the tensor alive
    return out


def check_order(order, n):
    """Raise unless `order` holds every index 0..n-1 exactly once (the kernels trust it: a wrong entry is an out-of-bounds
    access).  One small device pass + a host read; callers set an order once per batch shape, not per launch."""
    if order.numel() != n:
        raise ValueError(f"order must hold one entry per trajectory ({order.numel()} != {n})")
    if n and not bool((torch.bincount(order.clamp(0, n - 1).to(torch.int64), minlength=n) == 1).all().item()):
        raise ValueError("order is not a permutation of 0..n-1")
    if n and (int(order.min().item()) < 0 or int(order.max().item()) >= n):
        raise ValueError("order is not a permutation of 0..n-1")


def _check_integ_order(integ, n):
    ref = getattr(integ, "_order_ref", None)
    if ref is not None:
        check_order(ref, n)
    elif integ.order:
        raise ValueError("hb_integ.order set without hiten_b200.with_order (nothing keeps the device array alive)")


def cost_order(cost):
    """Launch order for `with_order`: most expensive first.  cost: CUDA tensor [n] (e.g. n_acc + n_rej of an earlier,
    similar batch).  Stable, so equal costs keep their index order."""
    key = cost if cost.is_floating_point() else cost.to(torch.int64)
    return torch.argsort(key, descending=True, stable=True).to(torch.int32).contiguous()


def _stream_ptr(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return L.vp(s.cuda_stream)


def _to_device_soa(y0, device):
    """Host [N, 6] (pinned staging, async H2D) or device [6, N] -> contiguous device [6, N] + host flag."""
    if isinstance(y0, torch.Tensor) and y0.is_cuda:
        if y0.dim() != 2 or y0.shape[0] != 6 or y0.dtype != torch.float64:
            raise ValueError("device input must be a float64 CUDA tensor of shape [6, N] (SoA)")
        return y0.contiguous(), False
    arr = np.ascontiguousarray(np.asarray(y0, dtype=np.float64))
    if arr.ndim != 2 or arr.shape[1] != 6:
        raise ValueError(f"Initial state array must have shape (N, 6), got {arr.shape}")
    soa = torch.from_numpy(np.ascontiguousarray(arr.T))
    return soa.to(device, non_blocking=False), True


def _alloc_out(n, device):
    i32 = dict(dtype=torch.int32, device=device)
    return (torch.empty((6, n), dtype=torch.float64, device=device), torch.empty(n, **i32),
            torch.empty(n, **i32), torch.empty(n, **i32))


def workspace(device):
    return torch.zeros(int(L.load().hb_workspace_bytes()) // 8, dtype=torch.int64, device=device)


def cr3bp_propagate(y0, mu, tf, *, t0=0.0, forward=1, flip=None, tf_per_traj=None, integ=None, device=None,
                    stream=None, ws=None):
    """Propagate a batch to tf; returns end states with the reference's _propagate_dynsys semantics."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        y0d, host = _to_device_soa(y0, device)
        n = y0d.shape[1]
        yf, nacc, nrej, status = _alloc_out(n, device)
        ws = workspace(device) if ws is None else ws
        integ = make_integ() if integ is None else integ
        _check_integ_order(integ, n)
        sys_ = make_sys(mu, forward, flip)
        tfp = None
        if tf_per_traj is not None:
            tfp = torch.as_tensor(np.asarray(tf_per_traj, dtype=np.float64)).to(device) \
                if not isinstance(tf_per_traj, torch.Tensor) else tf_per_traj
        rc = lib.hb_cr3bp_propagate(sys_, integ, n, y0d.data_ptr(), float(t0), float(tf),
                                    None if tfp is None else tfp.data_ptr(), 0, yf.data_ptr(), nacc.data_ptr(),
                                    nrej.data_ptr(), status.data_ptr(), ws.data_ptr(), _stream_ptr(stream))
        L.check(rc, "hb_cr3bp_propagate")
        if host:
            return BatchResult(yf.t().contiguous().cpu().numpy(), nacc.cpu().numpy(), nrej.cpu().numpy(),
                               status.cpu().numpy())
        return BatchResult(yf, nacc, nrej, status)


def cr3bp_dense(y0, mu, t_eval, *, forward=1, flip=None, integ=None, device=None, stream=None, ws=None,
                keep_on_device=False):
    """Propagate over t_eval[0]..t_eval[-1] (ascending) with dense output: states[N, m, 6]."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        y0d, host = _to_device_soa(y0, device)
        n = y0d.shape[1]
        te = t_eval if isinstance(t_eval, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64)))
        te = te.to(device).contiguous()
        m = te.numel()
        out = torch.empty((n, m, 6), dtype=torch.float64, device=device)
        _, nacc, nrej, status = _alloc_out(n, device)
        ws = workspace(device) if ws is None else ws
        integ = make_integ() if integ is None else integ
        _check_integ_order(integ, n)
        sys_ = make_sys(mu, forward, flip)
        rc = lib.hb_cr3bp_dense(sys_, integ, n, y0d.data_ptr(), te.data_ptr(), m, out.data_ptr(), nacc.data_ptr(),
                                nrej.data_ptr(), status.data_ptr(), ws.data_ptr(), _stream_ptr(stream))
        L.check(rc, "hb_cr3bp_dense")
        if host and not keep_on_device:
            return BatchResult(None, nacc.cpu().numpy(), nrej.cpu().numpy(), status.cpu().numpy(),
                               states=out.cpu().numpy())
        return BatchResult(None, nacc, nrej, status, states=out)


def cr3bp_event(y0, mu, tmax, event_idx, *, event_offset=0.0, direction=0, xtol=1e-12, gtol=1e-12, t0=0.0,
                forward=1, flip=None, tmax_per_traj=None, integ=None, device=None, stream=None, ws=None):
    """Propagate until the plane event y[idx] = offset (always terminal); refined by in-step bisection."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        y0d, host = _to_device_soa(y0, device)
        n = y0d.shape[1]
        yh, nacc, nrej, status = _alloc_out(n, device)
        th = torch.empty(n, dtype=torch.float64, device=device)
        ws = workspace(device) if ws is None else ws
        integ = make_integ() if integ is None else integ
        _check_integ_order(integ, n)
        sys_ = make_sys(mu, forward, flip)
        ev = L.HbEvent(int(event_idx), int(direction), float(event_offset), float(xtol), float(gtol))
        tmp = None
        if tmax_per_traj is not None:
            tmp = torch.as_tensor(np.asarray(tmax_per_traj, dtype=np.float64)).to(device) \
                if not isinstance(tmax_per_traj, torch.Tensor) else tmax_per_traj
        rc = lib.hb_cr3bp_event(sys_, integ, ev, n, y0d.data_ptr(), float(t0), float(tmax),
                                None if tmp is None else tmp.data_ptr(), th.data_ptr(), yh.data_ptr(),
                                nacc.data_ptr(), nrej.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                _stream_ptr(stream))
        L.check(rc, "hb_cr3bp_event")
        if host:
            return BatchResult(yh.t().contiguous().cpu().numpy(), nacc.cpu().numpy(), nrej.cpu().numpy(),
                               status.cpu().numpy(), t_hit=th.cpu().numpy())
        return BatchResult(yh, nacc, nrej, status, t_hit=th)


def cr3bp_stm(x0, mu, tf, *, t0=0.0, forward=1, flip=(36, 42), tf_per_traj=None, integ=None, device=None,
              stream=None, ws=None):
    """Batched _compute_stm end result: PHI rows [N, 42] at tf (Phi row-major, then the state).

    tf may differ per trajectory (tf_per_traj, e.g. the period of each family member)."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        x0d, host = _to_device_soa(x0, device)
        n = x0d.shape[1]
        _, nacc, nrej, status = _alloc_out(n, device)
        out = torch.empty((n, 42), dtype=torch.float64, device=device)
        ws = workspace(device) if ws is None else ws
        integ = make_integ() if integ is None else integ
        sys_ = make_sys(mu, forward, flip)
        tfp = None
        if tf_per_traj is not None:
            tfp = torch.as_tensor(np.asarray(tf_per_traj, dtype=np.float64)).to(device) \
                if not isinstance(tf_per_traj, torch.Tensor) else tf_per_traj
        rc = lib.hb_cr3bp_stm(sys_, integ, n, x0d.data_ptr(), float(t0), float(tf),
                              None if tfp is None else tfp.data_ptr(), out.data_ptr(), nacc.data_ptr(),
                              nrej.data_ptr(), status.data_ptr(), ws.data_ptr(), _stream_ptr(stream))
        L.check(rc, "hb_cr3bp_stm")
        if host:
            return BatchResult(None, nacc.cpu().numpy(), nrej.cpu().numpy(), status.cpu().numpy(),
                               states=out.cpu().numpy())
        return BatchResult(None, nacc, nrej, status, states=out)


def cr3bp_stm_dense(x0, mu, t_eval, *, forward=1, flip=(36, 42), integ=None, device=None, stream=None, ws=None,
                    keep_on_device=False):
    """Batched _compute_stm with dense output PHI[N, m, 42]; t_eval is [m] (shared) or [N, m]."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        x0d, host = _to_device_soa(x0, device)
        n = x0d.shape[1]
        te = t_eval if isinstance(t_eval, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64)))
        te = te.to(device).contiguous()
        per = 1 if te.dim() == 2 else 0
        m = int(te.shape[-1])
        if per and te.shape[0] != n:
            raise ValueError("per-trajectory t_eval must have shape [N, m]")
        out = torch.empty((n, m, 42), dtype=torch.float64, device=device)
        _, nacc, nrej, status = _alloc_out(n, device)
        ws = workspace(device) if ws is None else ws
        integ = make_integ() if integ is None else integ
        sys_ = make_sys(mu, forward, flip)
        rc = lib.hb_cr3bp_stm_dense(sys_, integ, n, x0d.data_ptr(), te.data_ptr(), m, per, out.data_ptr(),
                                    nacc.data_ptr(), nrej.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                    _stream_ptr(stream))
        L.check(rc, "hb_cr3bp_stm_dense")
        if host and not keep_on_device:
            return BatchResult(None, nacc.cpu().numpy(), nrej.cpu().numpy(), status.cpu().numpy(),
                               states=out.cpu().numpy())
        return BatchResult(None, nacc, nrej, status, states=out)


def dfma_peak(millis=50.0):
    """Measured FP64 FMA flop/s of the current device (roofline denominator)."""
    _require_cuda()
    out = L.C.c_double(0.0)
    L.check(L.load().hb_dfma_peak(float(millis), L.C.byref(out), _stream_ptr()), "hb_dfma_peak")
    return out.value
