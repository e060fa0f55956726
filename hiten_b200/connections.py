"""Connection search between two sets of section hits on the GPU (host wrapper over hb_connections).

Reference: hiten/algorithms/connections/backends.py (_ConnectionsBackend.run :425-540).  Inputs are what
SynodicMap sections deliver (2-D plane points, 6-states, trajectory indices) -- e.g. the hit buffers of
synodic.tube_section for the unstable and the stable manifold.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from .propagate import _require_cuda, _stream_ptr

CONN_DTYPE = np.dtype([("index_u", "<i8"), ("index_s", "<i8"), ("delta_v", "<f8"), ("point2d", "<f8", (2,)),
                       ("state_u", "<f8", (6,)), ("state_s", "<f8", (6,)), ("kind", "<i8")])
assert CONN_DTYPE.itemsize == 144
KINDS = ("ballistic", "impulsive")


@dataclass
class Connections:
    """Accepted connections in the reference's order (ascending delta_v, ties in pair order)."""
    kind: np.ndarray                 # [K] 0 ballistic, 1 impulsive
    delta_v: np.ndarray              # [K]
    point2d: np.ndarray              # [K, 2]
    state_u: np.ndarray              # [K, 6]
    state_s: np.ndarray              # [K, 6]
    index_u: np.ndarray              # [K]
    index_s: np.ndarray              # [K]
    trajectory_index_u: np.ndarray   # [K]
    trajectory_index_s: np.ndarray   # [K]
    pairs_considered: int


def _dev(a, device, cols):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    t = t.to(device=device, dtype=torch.float64).contiguous()
    if t.dim() != 2 or t.shape[1] != cols:
        raise ValueError(f"expected an array of shape (N, {cols})")
    return t


def find_connections(points_u, points_s, states_u, states_s, eps, dv_tol, bal_tol, *, traj_indices_u=None,
                     traj_indices_s=None, device=None, stream=None):
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        pu, ps = _dev(points_u, device, 2), _dev(points_s, device, 2)
        Xu, Xs = _dev(states_u, device, 6), _dev(states_s, device, 6)
        nu, ns = int(pu.shape[0]), int(ps.shape[0])
        if Xu.shape[0] != nu or Xs.shape[0] != ns:
            raise ValueError("points and states must have the same number of rows")
        cap = max(min(nu, ns), 1)
        out = torch.empty(cap * 18, dtype=torch.float64, device=device)
        nbytes = int(lib.hb_connections_scratch_bytes(nu, ns))
        scratch = torch.empty(max(nbytes, 8) // 8 + 1, dtype=torch.float64, device=device)
        n_out, n_drop, considered = L.C.c_int64(0), L.C.c_int64(0), L.C.c_int64(0)
        rc = lib.hb_connections(pu.data_ptr(), nu, ps.data_ptr(), ns, Xu.data_ptr(), Xs.data_ptr(), float(eps),
                                float(dv_tol), float(bal_tol), out.data_ptr(), cap, L.C.byref(n_out), L.C.byref(n_drop),
                                L.C.byref(considered), scratch.data_ptr(), scratch.numel() * 8, _stream_ptr(stream))
        L.check(rc, "hb_connections")
        k = int(n_out.value)
        rec = out[: k * 18].cpu().numpy().view(CONN_DTYPE) if k else np.empty(0, dtype=CONN_DTYPE)
    rec = rec[np.lexsort((rec["index_u"], rec["delta_v"]))]
    iu, js = rec["index_u"].copy(), rec["index_s"].copy()
    tu = np.asarray(traj_indices_u)[iu].astype(np.int64) if traj_indices_u is not None else np.zeros(len(iu), np.int64)
    ts = np.asarray(traj_indices_s)[js].astype(np.int64) if traj_indices_s is not None else np.zeros(len(js), np.int64)
    return Connections(rec["kind"].copy(), rec["delta_v"].copy(), rec["point2d"].copy(), rec["state_u"].copy(),
                       rec["state_s"].copy(), iu, js, tu, ts, int(considered.value))
