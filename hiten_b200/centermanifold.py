"""Centre-manifold Poincare map on the GPU (host wrapper over hb_cm_poincare_map).

Reference: hiten/algorithms/poincare/centermanifold/backend.py (_CenterManifoldBackend.run :404-466,
_poincare_map :314-382).  The polynomial tables arrive as the reference's jac_H / clmo_table and are reduced
once, on the host, to the sparse real term list the kernel keeps in shared memory.
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from .propagate import _require_cuda, _stream_ptr, workspace

SECTION = {"q2": 0, "p2": 1, "q3": 2, "p3": 3}
TERM_DTYPE = np.dtype([("coef", "<f8"), ("ex", "<u8")])


@dataclass
class PolyTable:
    """Sparse gradient table: terms of partial p are rows ptr[p]:ptr[p+1] (order of evaluation preserved)."""
    ptr: np.ndarray      # [7] int64
    deg: np.ndarray      # [T] int32
    coef: np.ndarray     # [T] float64
    exp: np.ndarray      # [T, 6] int32
    _dev: dict = None

    @property
    def max_deg(self):
        return int(self.exp.max()) if self.exp.size else 0

    @classmethod
    def from_reference(cls, jac_H, clmo):
        """jac_H[p][d] = packed complex coefficients of degree d of dH/dvar_p; clmo[d][i] = packed exponents
        (6 bits each for k1..k5, k0 = d - sum; hiten/algorithms/polynomial/base.py:222-259)."""
        ptr, deg, coef, exp = [0], [], [], []
        for p in range(6):
            for d in range(len(jac_H[p])):
                arr = np.asarray(jac_H[p][d])
                nz = np.nonzero(arr != 0)[0]
                if nz.size == 0:
                    continue
                packed = np.asarray(clmo[d])[nz].astype(np.int64)
                ks = np.stack([(packed >> s) & 0x3F for s in (0, 6, 12, 18, 24)], axis=1)
                k0 = d - ks.sum(axis=1)
                deg.append(np.full(nz.size, d, dtype=np.int32))
                coef.append(np.real(arr[nz]).astype(np.float64))
                exp.append(np.column_stack([k0, ks]).astype(np.int32))
            ptr.append(sum(len(x) for x in deg))
        if not deg:
            return cls(np.array(ptr, dtype=np.int64), np.empty(0, np.int32), np.empty(0), np.empty((0, 6), np.int32))
        return cls(np.array(ptr, dtype=np.int64), np.concatenate(deg), np.concatenate(coef), np.concatenate(exp))

    @classmethod
    def from_hamiltonian(cls, H_blocks, clmo):
        """The Hamiltonian ITSELF (hamsys.poly_H(): H_blocks[d] = packed coefficients of degree d) as polynomial 0
        of a table -- what lift_plane_points evaluates."""
        t = cls.from_reference([H_blocks] + [[]] * 5, clmo)
        return t

    @classmethod
    def single(cls, deg, coef, exp):
        """One polynomial given as sparse terms in evaluation order (polynomial 0 of the table)."""
        n = len(deg)
        return cls(np.array([0] + [n] * 6, dtype=np.int64), np.asarray(deg, dtype=np.int32),
                   np.asarray(coef, dtype=np.float64), np.asarray(exp, dtype=np.int32).reshape(-1, 6))

    def packed(self):
        rec = np.empty(self.coef.size, dtype=TERM_DTYPE)
        rec["coef"] = self.coef
        ex = np.zeros(self.coef.size, dtype=np.uint64)
        for v in range(6):
            ex |= self.exp[:, v].astype(np.uint64) << np.uint64(8 * v)
        ex |= self.deg.astype(np.uint64) << np.uint64(48)
        rec["ex"] = ex
        return rec

    def device_struct(self, device):
        key = str(device)
        if self._dev is None:
            self._dev = {}
        if key not in self._dev:
            rec = self.packed()
            buf = torch.from_numpy(rec.view(np.float64).copy() if rec.size else np.zeros(2)).to(device)
            self._dev[key] = buf
        buf = self._dev[key]
        return L.HbPolyHam(3, self.max_deg, (L.C.c_int64 * 7)(*[int(x) for x in self.ptr]), buf.data_ptr()), buf


def make_opts(dt=0.01, max_steps=2000, method="fixed", order=4, section_coord="q3", c_omega_heuristic=20.0,
              arith="parity"):
    """method/order as in IntegrationOptions + config.integration.method ("fixed" | "symplectic")."""
    if method == "adaptive":
        raise NotImplementedError("Adaptive integrator is not implemented in CM backend; use 'fixed' (RK) or 'symplectic'.")
    if method == "symplectic":
        m = L.HB_SYMPLECTIC
    elif method == "fixed":
        m = {4: L.HB_RK4, 6: L.HB_RK6, 8: L.HB_RK8}.get(int(order))
        if m is None:
            raise ValueError("RK order must be 4, 6, or 8")
    else:
        raise ValueError(f"unknown integration method {method!r}")
    o = L.HbCmOpts()
    o.dt, o.max_steps, o.method, o.order = float(dt), int(max_steps), m, int(order)
    o.section = SECTION[section_coord]
    o.arith = {"parity": L.HB_ARITH_PARITY, "fast": L.HB_ARITH_FAST}[arith]
    L.check(L.load().hb_cm_prepare(L.C.byref(o), float(c_omega_heuristic)), "hb_cm_prepare")
    return o


def jit_compile_host(table, opts, want_source=False):
    """Generate + compile the specialised kernel offline (no GPU): returns (cubin_bytes, source or None)."""
    lib = L.load()
    rec = table.packed()
    ptr = (L.C.c_int64 * 7)(*[int(x) for x in table.ptr])
    nbytes = L.C.c_int64(0)
    cap = 4 << 20
    buf = L.C.create_string_buffer(cap) if want_source else None
    rc = lib.hb_cm_jit_compile_host(rec.ctypes.data if rec.size else None, ptr, table.max_deg, opts, L.C.byref(nbytes),
                                    buf, cap if want_source else 0)
    if rc != 0 and buf is not None:
        raise L.HitenB200Error(f"hb_cm_jit_compile_host: {rc}: {buf.value.decode(errors='replace')[:2000]}")
    L.check(rc, "hb_cm_jit_compile_host")
    return int(nbytes.value), (buf.value.decode() if want_source else None)


def poincare_map(table, seeds, opts, *, device=None, stream=None, ws=None, jit=True):
    """_poincare_map on the GPU: seeds [N, 4] (host ndarray or CUDA tensor) -> (flags, states[N,4], times[N]).

    jit=True (default) runs the kernel specialised for this Hamiltonian (hb_cm_poincare_map_jit);
    jit=False the table-driven kernel (hb_cm_poincare_map).  Both are bit-identical in the parity variant."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        host = not (isinstance(seeds, torch.Tensor) and seeds.is_cuda)
        sd = torch.from_numpy(np.ascontiguousarray(seeds, dtype=np.float64)).to(device) if host else seeds.contiguous()
        if sd.dim() != 2 or sd.shape[1] != 4:
            raise ValueError("seeds must have shape (N, 4) = (q2, p2, q3, p3)")
        n = int(sd.shape[0])
        flags = torch.zeros(n, dtype=torch.int32, device=device)
        out = torch.zeros((n, 4), dtype=torch.float64, device=device)
        tt = torch.zeros(n, dtype=torch.float64, device=device)
        ws = workspace(device) if ws is None else ws
        ham, keep = table.device_struct(device)
        fn = lib.hb_cm_poincare_map_jit if jit else lib.hb_cm_poincare_map
        rc = fn(ham, opts, n, sd.data_ptr(), flags.data_ptr(), out.data_ptr(), tt.data_ptr(), ws.data_ptr(),
                _stream_ptr(stream))
        L.check(rc, "hb_cm_poincare_map_jit" if jit else "hb_cm_poincare_map")
        if host:
            return flags.cpu().numpy().astype(np.int64), out.cpu().numpy(), tt.cpu().numpy()
        return flags, out, tt


_STATE_INDEX = {"q2": 0, "p2": 1, "q3": 2, "p3": 3}      # columns of a (q2, p2, q3, p3) state row


def poincare_map_iterate(table, seeds, opts, n_iter, section_coord, *, device=None, stream=None, jit=True):
    """The iteration loop of `_CenterManifoldEngine.solve._worker` (centermanifold/engine.py:163-191) with the seeds
    resident in HBM: every pass maps the current seeds (`poincare_map`), drops the failed ones, zeroes the section
    coordinate (`enforce_section_coordinate`, interfaces.py:338-346), accumulates the hits and feeds them back as the
    next seeds.  Returns (states[M, 4], times[M], iteration[M]) as device tensors, in the reference's order (iteration
    by iteration, seed order inside an iteration)."""
    _require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    col = _STATE_INDEX[section_coord]
    with torch.cuda.device(device):
        cur = seeds if isinstance(seeds, torch.Tensor) and seeds.is_cuda else torch.from_numpy(
            np.ascontiguousarray(seeds, dtype=np.float64)).to(device)
        ws = workspace(device)
        st_acc, t_acc, it_acc = [], [], []
        for it in range(int(n_iter)):
            if cur.shape[0] == 0:
                break
            flags, out, tt = poincare_map(table, cur, opts, device=device, stream=stream, ws=ws, jit=jit)
            ok = flags == 1
            nxt = out[ok].clone()
            if nxt.shape[0] == 0:
                break
            nxt[:, col] = 0.0
            st_acc.append(nxt)
            t_acc.append(tt[ok])
            it_acc.append(torch.full((nxt.shape[0],), it, dtype=torch.int32, device=device))
            cur = nxt
        if not st_acc:
            z = torch.empty
            return (z((0, 4), dtype=torch.float64, device=device), z(0, dtype=torch.float64, device=device),
                    z(0, dtype=torch.int32, device=device))
        return torch.cat(st_acc), torch.cat(t_acc), torch.cat(it_acc)


def lift_plane_points(H_table, section_coord, plane_points, h0, *, initial_guess=1e-3, expand_factor=2.0, max_expand=40,
                      symmetric=False, xtol=1e-12, device=None, stream=None):
    """Batched _CenterManifoldInterface.lift_plane_point (interfaces.py:297-337): plane_points [N, 2] in the section's
    plane coordinates -> (ok[N] bool, states[N, 4] = (q2, p2, q3, p3)); ok is False where the reference returns None.
    H_table: PolyTable.from_hamiltonian(hamsys.poly_H(), hamsys.clmo_table).  Bit-identical to the reference."""
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        host = not (isinstance(plane_points, torch.Tensor) and plane_points.is_cuda)
        pts = torch.from_numpy(np.ascontiguousarray(plane_points, dtype=np.float64)).to(device) if host \
            else plane_points.contiguous()
        if pts.dim() != 2 or pts.shape[1] != 2:
            raise ValueError("plane_points must have shape (N, 2)")
        n = int(pts.shape[0])
        ok = torch.zeros(n, dtype=torch.int32, device=device)
        out = torch.zeros((n, 4), dtype=torch.float64, device=device)
        o = L.HbCmLiftOpts(float(h0), float(initial_guess), float(expand_factor), float(xtol), int(max_expand),
                           int(bool(symmetric)), SECTION[section_coord], 200)
        ham, keep = H_table.device_struct(device)
        L.check(lib.hb_cm_lift(ham, L.C.byref(o), n, pts.data_ptr(), out.data_ptr(), ok.data_ptr(), _stream_ptr(stream)),
                "hb_cm_lift")
        if host:
            return ok.cpu().numpy().astype(bool), out.cpu().numpy()
        return ok.bool(), out
