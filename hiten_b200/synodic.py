"""Synodic-section detection on the GPU (host wrapper over hb_synodic_detect).

Reference: hiten/algorithms/poincare/synodic/backend.py (_SynodicDetectionBackend.run :823-887).
Axis / plane-coordinate names follow _PlaneEvent._IDX_MAP (x, y, z, vx, vy, vz -> 0..5).
"""
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from .propagate import _require_cuda, _stream_ptr, workspace

IDX = {"x": 0, "y": 1, "z": 2, "vx": 3, "vy": 4, "vz": 5}
HIT_DTYPE = np.dtype([("traj", "<i8"), ("seq", "<i8"), ("t", "<f8"), ("state", "<f8", (6,))])
assert HIT_DTYPE.itemsize == 72


@dataclass
class SectionHits:
    """Hits in the reference's order: by trajectory, then by position along the trajectory."""
    trajectory_indices: np.ndarray   # [K] int64
    times: np.ndarray                # [K]
    states: np.ndarray               # [K, 6]
    points: np.ndarray               # [K, 2]
    hits_per_traj: np.ndarray        # [N] int32


def make_section(section_axis="x", section_offset=0.0, plane_coords=("y", "vy"), direction=None,
                 segment_refine=50, tol_on_surface=1e-6, dedup_time_tol=1e-9, dedup_point_tol=1e-6,
                 max_hits_per_traj=None):
    """hb_section with the reference's SynodicMap defaults (algorithms/types/services/maps.py:753-774)."""
    idx = IDX[section_axis.lower()] if isinstance(section_axis, str) else int(section_axis)
    pi, pj = (IDX[c.lower()] if isinstance(c, str) else int(c) for c in plane_coords)
    return L.HbSection(idx, 0 if direction is None else int(direction), float(section_offset), pi, pj,
                       int(segment_refine), 0 if max_hits_per_traj is None else int(max_hits_per_traj),
                       float(tol_on_surface), float(dedup_time_tol), float(dedup_point_tol))


def detect(states, times, section, *, offsets=None, hit_capacity=None, device=None, stream=None, ws=None,
           interp_kind="linear", newton_max_iter=4):
    """Detect section hits.

    states : CUDA tensor / ndarray [N, m, 6] (uniform) or [sum m_i, 6] with `offsets` [N + 1];
    times  : [m] shared signed times, [N, m], or concatenated [sum m_i].
    interp_kind : "linear" (what the shipped SynodicMap always requests) or "cubic" with `newton_max_iter` Newton steps
    (backend.py:695-701, 762; hb_synodic_detect_cubic).
    """
    if interp_kind not in ("linear", "cubic"):
        raise ValueError("interp_kind must be 'linear' or 'cubic'")
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        st = states if isinstance(states, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(states, dtype=np.float64))
        st = st.to(device).contiguous()
        tm = times if isinstance(times, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(times, dtype=np.float64))
        tm = tm.to(device).contiguous()
        if offsets is None:
            if st.dim() != 3 or st.shape[2] != 6:
                raise ValueError("states must have shape [N, m, 6] when offsets is None")
            n, m = int(st.shape[0]), int(st.shape[1])
            shared = 1 if tm.dim() == 1 and tm.numel() == m else 0
            if not shared and tm.numel() != n * m:
                raise ValueError("times must have m or N*m entries")
            off_t = None
        else:
            off_np = np.ascontiguousarray(offsets, dtype=np.int64)
            n, m, shared = off_np.size - 1, 0, 0
            off_t = torch.from_numpy(off_np).to(device)
        cap = int(hit_capacity) if hit_capacity is not None else max(1024, 8 * n)
        ws = workspace(device) if ws is None else ws
        while True:
            hits = torch.empty(cap * 9, dtype=torch.float64, device=device)      # 72-byte records
            per = torch.empty(max(n, 1), dtype=torch.int32, device=device)
            tail = (n, st.data_ptr(), tm.data_ptr(), None if off_t is None else off_t.data_ptr(), m, shared,
                    hits.data_ptr(), cap, per.data_ptr(), ws.data_ptr(), _stream_ptr(stream))
            if interp_kind == "cubic":
                rc = lib.hb_synodic_detect_cubic(section, int(newton_max_iter), *tail)
            else:
                rc = lib.hb_synodic_detect(section, *tail)
            L.check(rc, "hb_synodic_detect")
            nh, no = L.C.c_int64(0), L.C.c_int64(0)
            L.check(lib.hb_read_hit_count(ws.data_ptr(), L.C.byref(nh), L.C.byref(no), _stream_ptr(stream)),
                    "hb_read_hit_count")
            if no.value == 0:
                break
            cap = int(nh.value) + 16                                             # retry once with room for all
        k = int(nh.value)
        rec = hits[: k * 9].cpu().numpy().view(HIT_DTYPE) if k else np.empty(0, dtype=HIT_DTYPE)
        order = np.lexsort((rec["seq"], rec["traj"]))
        rec = rec[order]
        pts = np.column_stack((rec["state"][:, section.proj_i], rec["state"][:, section.proj_j])) if k else np.empty((0, 2))
        return SectionHits(rec["traj"].copy(), rec["t"].copy(), rec["state"].copy(), pts, per[:n].cpu().numpy())


def _auto_steps_capacity(n, device, want=160):
    """Accepted steps per trajectory the step scratch of hb_cr3bp_section2 can hold in ~60 % of the free HBM
    (512 B per step); 0 = use the fused kernel (tiny batches, or no room)."""
    if n < 256:
        return 0
    free, _ = torch.cuda.mem_get_info(device)
    cap = min(want, int(0.6 * free / (n * 512.0)))
    cap -= cap % 32
    return cap if cap >= 64 else 0


def _auto_plan(n, device, records, free_bytes=None):
    """(chunk, steps_capacity) for tube_section(steps_capacity="auto"): the kernel pipeline with a step scratch for `chunk`
    trajectories at a time -- the whole batch when its scratch fits half of the free HBM, else equal chunks that do (a
    B200 holds ~1.2e6 trajectories of sparse records per chunk; the chunks reuse one scratch).  (n, 0): tiny batch, fused
    kernel.  free_bytes: override of the free-memory probe (tests)."""
    if n < 256:
        return n, 0
    cap = 128 if records == "near" else 160
    lib = L.load()
    per_traj = int(lib.hb_section2_scratch_bytes(1024, cap)) / 1024.0 + 400.0     # + hits, end states, counters
    free = torch.cuda.mem_get_info(device)[0] if free_bytes is None else int(free_bytes)
    fit = int(0.5 * free / per_traj)
    if fit >= n:
        return n, cap
    if fit < 1024:
        return n, 0                                             # no room for a useful scratch: fused kernel, no scratch
    n_chunks = -(-n // fit)
    chunk = -(-n // n_chunks)
    return chunk + (-chunk) % 256, cap


def tube_section(y0, mu, t_eval, section, *, forward=1, flip=None, integ=None, hit_capacity=None, device=None,
                 stream=None, ws=None, sort=True, steps_capacity="auto", records="near", _free_bytes=None):
    """Manifold.compute() + SynodicMap.compute() in one call: propagate a batch over the t_eval grid and detect the
    section hits on the device, without storing the dense tube.  Returns (SectionHits, BatchResult with end states).

    steps_capacity: "auto" (default) picks the kernel pipeline hb_cr3bp_section2 when the batch is large enough -- in
    equal chunks that reuse one step scratch when the whole batch's scratch does not fit half of the free device memory
    (any batch size runs: 1e7 trajectories are ~9 chunks on a B200) -- else the fused kernel hb_cr3bp_section; an int
    forces the scratch size (0 = fused kernel).  All forms give the same hits, bit for bit."""
    from . import propagate as P
    _require_cuda()
    lib = L.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(device):
        y0d, host = P._to_device_soa(y0, device)
        n = y0d.shape[1]
        chunk = n
        if steps_capacity == "auto":
            chunk, steps_capacity = _auto_plan(n, device, records, _free_bytes) if sort else (n, 0)
        if steps_capacity and sort and n > 0 and chunk < n:
            return _tube_section_chunked(y0d, host, chunk, int(steps_capacity), mu, t_eval, section, forward, flip, integ,
                                         device, stream, records)
        if steps_capacity and sort and n > 0:
            run = TubeSectionRunner(n, mu, t_eval, section, forward=forward, flip=flip, integ=integ,
                                    hit_capacity=hit_capacity, device=device, steps_capacity=int(steps_capacity),
                                    records=records)
            run.launch(y0d, stream)
            h = run.sorted_hits(stream)
            if host:
                res = P.BatchResult(run.yf.t().contiguous().cpu().numpy(), run.nacc.cpu().numpy(),
                                    run.nrej.cpu().numpy(), run.status.cpu().numpy())
            else:
                res = P.BatchResult(run.yf, run.nacc, run.nrej, run.status)
            return h, res
        te = t_eval if isinstance(t_eval, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64)))
        te = te.to(device).contiguous()
        m = te.numel()
        yf, nacc, nrej, status = P._alloc_out(n, device)
        per = torch.zeros(max(n, 1), dtype=torch.int32, device=device)
        ws = workspace(device) if ws is None else ws
        integ = P.make_integ() if integ is None else integ
        sys_ = P.make_sys(mu, forward, flip)
        cap = int(hit_capacity) if hit_capacity is not None else max(1024, 8 * n)
        while True:
            hits = torch.empty(cap * 9, dtype=torch.float64, device=device)
            rc = lib.hb_cr3bp_section(sys_, integ, section, n, y0d.data_ptr(), te.data_ptr(), m, hits.data_ptr(), cap,
                                      per.data_ptr(), yf.data_ptr(), nacc.data_ptr(), nrej.data_ptr(),
                                      status.data_ptr(), ws.data_ptr(), _stream_ptr(stream))
            L.check(rc, "hb_cr3bp_section")
            nh, no = L.C.c_int64(0), L.C.c_int64(0)
            L.check(lib.hb_read_hit_count(ws.data_ptr(), L.C.byref(nh), L.C.byref(no), _stream_ptr(stream)),
                    "hb_read_hit_count")
            if no.value == 0:
                break
            cap = int(nh.value) + 16
        k = int(nh.value)
        if host:
            res = P.BatchResult(yf.t().contiguous().cpu().numpy(), nacc.cpu().numpy(), nrej.cpu().numpy(),
                                status.cpu().numpy())
        else:
            res = P.BatchResult(yf, nacc, nrej, status)
        if not sort:
            return (hits[: k * 9], per[:n], k), res
        rec = hits[: k * 9].cpu().numpy().view(HIT_DTYPE) if k else np.empty(0, dtype=HIT_DTYPE)
        rec = rec[np.lexsort((rec["seq"], rec["traj"]))]
        pts = np.column_stack((rec["state"][:, section.proj_i], rec["state"][:, section.proj_j])) if k else np.empty((0, 2))
        return SectionHits(rec["traj"].copy(), rec["t"].copy(), rec["state"].copy(), pts, per[:n].cpu().numpy()), res


def _tube_section_chunked(y0d, host, chunk, steps_capacity, mu, t_eval, section, forward, flip, integ, device, stream,
                          records):
    """tube_section over a batch whose step scratch does not fit: `chunk` trajectories at a time through one runner (and
    one scratch); the last, shorter chunk gets a runner of its own size on the same scratch.  Trajectory indices of the
    hits are those of the whole batch."""
    from . import propagate as P
    n = y0d.shape[1]
    yf, nacc, nrej, status = P._alloc_out(n, device)
    per = np.zeros(n, dtype=np.int32)
    parts = []
    run = None
    for a in range(0, n, chunk):
        b = min(a + chunk, n)
        if run is None or run.n != b - a:
            scratch = None if run is None else run.scratch
            run = TubeSectionRunner(b - a, mu, t_eval, section, forward=forward, flip=flip, integ=integ, device=device,
                                    steps_capacity=steps_capacity, records=records, scratch=scratch)
        run.launch(y0d[:, a:b].contiguous(), stream)
        h = run.sorted_hits(stream)
        parts.append((h.trajectory_indices + a, h.times, h.states, h.points))
        per[a:b] = h.hits_per_traj
        yf[:, a:b] = run.yf
        nacc[a:b], nrej[a:b], status[a:b] = run.nacc, run.nrej, run.status
    hits = SectionHits(np.concatenate([q[0] for q in parts]), np.concatenate([q[1] for q in parts]),
                       np.concatenate([q[2] for q in parts]), np.concatenate([q[3] for q in parts]), per)
    if host:
        res = P.BatchResult(yf.t().contiguous().cpu().numpy(), nacc.cpu().numpy(), nrej.cpu().numpy(), status.cpu().numpy())
    else:
        res = P.BatchResult(yf, nacc, nrej, status)
    return hits, res


class TubeSectionRunner:
    """Pre-allocated, repeatable form of tube_section for resident batches (what bench.py times): all device
    buffers are created once; launch() enqueues the kernels, hit_count() / sorted_hits() read the result."""

    def __init__(self, n, mu, t_eval, section, *, forward=1, flip=None, integ=None, hit_capacity=None, device=None,
                 steps_capacity=0, scratch=None, filters=None, pool_records=0, records="near"):
        """Three forms of the same step, same hits bit for bit:
        pool_records > 0 selects hb_cr3bp_section3: step records handed from propagating to scanning warps through
        shared memory inside one kernel; the scratch holds the candidate lists and a pool of `pool_records` step
        records per trajectory (512 B each; 8 is plenty for a tube) -- no per-trajectory step capacity;
        steps_capacity > 0 selects the kernel pipeline hb_cr3bp_section2 with an HBM scratch for that many accepted
        steps per trajectory (512 B per step); both 0: the fused kernel hb_cr3bp_section.  Trajectories that do not fit
        the scratch / pool are rerun with the fused kernel by hit_count() / sorted_hits(), so the result is the same.
        `scratch` lets several runners share one (large) scratch tensor.
        records (hb_cr3bp_section2 only): "near" (default) -- the propagation kernel records only the steps that can come
        near the section plane and their neighbours, steps_capacity counts those (a tube needs ~10-30; 64 is plenty);
        "all" -- every accepted step (what `filters` needs; steps_capacity = accepted steps per trajectory).
        `filters` = (safe_r1, safe_r2, energy_tol) applies Manifold.compute()'s trajectory filters
        (services/manifold.py:412-432) from the step records (hb_section2_filter, pipeline only): sorted_hits() then
        returns the hits of the kept trajectories only -- what SynodicMap sees after manifold.compute()."""
        from . import propagate as P
        _require_cuda()
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n, self.section = int(n), section
        te = t_eval if isinstance(t_eval, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64)))
        self.te = te.to(self.device).contiguous()
        self.yf, self.nacc, self.nrej, self.status = P._alloc_out(self.n, self.device)
        self.per = torch.zeros(max(self.n, 1), dtype=torch.int32, device=self.device)
        self.ws = workspace(self.device)
        self.integ = P.make_integ() if integ is None else integ
        self.mu, self.forward, self.flip = mu, forward, flip
        self.sys = P.make_sys(mu, forward, flip)
        self.cap = int(hit_capacity) if hit_capacity is not None else max(1024, 8 * self.n)
        self._fixed_cap = hit_capacity is not None           # an explicit capacity is a hard limit (overflow raises)
        self.hits = torch.empty(self.cap * 9, dtype=torch.float64, device=self.device)
        self.steps_capacity = int(steps_capacity)
        self.pool_records = int(pool_records)
        self.done_event = None
        if records not in ("near", "all"):
            raise ValueError("records must be 'near' or 'all'")
        self.records = L.HB_RECORDS_ALL if (records == "all" or filters is not None) else L.HB_RECORDS_NEAR_SECTION
        if self.pool_records > 0 and self.steps_capacity > 0:
            raise ValueError("choose pool_records (hb_cr3bp_section3) or steps_capacity (hb_cr3bp_section2), not both")
        self.stage_events = None        # set_stage_events(): caller-owned CUDA events around the pipeline's stages
        self.scratch = None
        self._y0 = None
        self._order, self._integ_ordered = None, None        # set_order(): launch order of the persistent propagation kernel
        self._extra = None          # (indices, SectionHits) of the trajectories rerun with the fused kernel
        self.filters = None
        if filters is not None:
            if self.steps_capacity <= 0:
                raise ValueError("filters are computed from the step records of hb_cr3bp_section2: use steps_capacity > 0")
            self.filters = L.HbTubeFilterOpts(float(mu), float(filters[0]), float(filters[1]), float(filters[2]))
            self.filt = torch.empty((max(self.n, 1), 3), dtype=torch.float64, device=self.device)
            self.keep = torch.empty(max(self.n, 1), dtype=torch.int32, device=self.device)
        if self.steps_capacity > 0 or self.pool_records > 0:
            nbytes = int(self.lib.hb_section2_scratch_bytes(self.n, self.steps_capacity)) if self.steps_capacity > 0 \
                else int(self.lib.hb_section3_scratch_bytes(self.n, self.pool_records))
            self._owns_scratch = scratch is None
            if scratch is not None:
                if scratch.numel() * scratch.element_size() < nbytes or scratch.device != self.device:
                    raise ValueError("scratch tensor too small or on another device")
                self.scratch = scratch
            else:
                self.scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=self.device)

    def set_order(self, order):
        """Scheduling hint for the pipeline's persistent propagation launch (hb_integ.order): `order` is an int32 CUDA
        tensor with a permutation of 0..n-1 -- trajectory order[q] is handed out q-th -- or None for 0, 1, 2, ...
        Results do not depend on it (outputs stay in the caller's indexing); longest expected trajectories first
        shortens the launch's tail.  Honoured by hb_cr3bp_section2 (steps_capacity > 0)."""
        from . import propagate as P
        if order is not None:
            P.check_order(order, self.n)                      # the kernel trusts it
        self._order = order                                   # keeps the tensor alive
        self._integ_ordered = None if order is None else P.with_order(self.integ, order)

    def order_by_cost(self, cost=None):
        """set_order(most expensive first).  cost: CUDA tensor [n]; default: the attempted steps (n_acc + n_rej) of this
        runner's LAST launch -- for callers whose next batch resembles the last one (the same tube cut by another
        section, the next displacement set of the same orbit, the next iteration of a continuation)."""
        from . import propagate as P
        if cost is None:
            cost = self.nacc[: self.n] + self.nrej[: self.n]
        self.set_order(P.cost_order(cost))

    def set_stage_events(self, events):
        """events: list of torch.cuda.Event(enable_timing=True) the library records on the launch stream around the
        pipeline's stages (5 for hb_cr3bp_section2, 4 for hb_cr3bp_section3), or None.  bench.py's roofline uses it."""
        if events is None:
            self.stage_events = None
            return
        for e in events:
            e.record()                                       # torch creates the CUDA event lazily, on first record
        arr = (L.C.c_void_p * len(events))(*[e.cuda_event for e in events])
        self.stage_events = (arr, list(events))

    def launch(self, y0_soa, stream=None):
        try:
            self._launch(y0_soa, stream)
        finally:
            # end of THIS pipeline on the launch stream: lets a caller wait for it while later launches keep running
            if self.done_event is None:
                self.done_event = torch.cuda.Event()
            self.done_event.record(torch.cuda.current_stream(self.device) if stream is None else stream)

    def _launch(self, y0_soa, stream=None):
        self._y0, self._extra = y0_soa, None
        sev = None if self.stage_events is None else self.stage_events[0]
        if self.scratch is not None and self.pool_records > 0:
            rc = self.lib.hb_cr3bp_section3(self.sys, self.integ, self.section, self.n, y0_soa.data_ptr(),
                                            self.te.data_ptr(), self.te.numel(), self.hits.data_ptr(), self.cap,
                                            self.per.data_ptr(), self.yf.data_ptr(), self.nacc.data_ptr(),
                                            self.nrej.data_ptr(), self.status.data_ptr(), self.scratch.data_ptr(),
                                            self.scratch.numel() * 8, self.ws.data_ptr(), _stream_ptr(stream), sev)
            L.check(rc, "hb_cr3bp_section3")
            if self.filters is not None:
                raise L.HitenB200Error("filters need the step records of hb_cr3bp_section2 (steps_capacity > 0)")
            return
        if self.scratch is not None:
            rc = self.lib.hb_cr3bp_section2(self.sys, self.integ if self._integ_ordered is None else self._integ_ordered,
                                            self.section, self.n, y0_soa.data_ptr(),
                                            self.te.data_ptr(), self.te.numel(), self.hits.data_ptr(), self.cap,
                                            self.per.data_ptr(), self.yf.data_ptr(), self.nacc.data_ptr(),
                                            self.nrej.data_ptr(), self.status.data_ptr(), self.scratch.data_ptr(),
                                            self.scratch.numel() * 8, self.ws.data_ptr(), _stream_ptr(stream), sev,
                                            self.records)
            L.check(rc, "hb_cr3bp_section2")
            if self.filters is not None:
                rc = self.lib.hb_section2_filter(self.sys, self.integ, self.filters, self.n, self.te.data_ptr(),
                                                 self.te.numel(), self.nacc.data_ptr(), self.status.data_ptr(),
                                                 self.scratch.data_ptr(), self.scratch.numel() * 8,
                                                 self.filt.data_ptr(), self.keep.data_ptr(), _stream_ptr(stream))
                L.check(rc, "hb_section2_filter")
            return
        rc = self.lib.hb_cr3bp_section(self.sys, self.integ, self.section, self.n, y0_soa.data_ptr(),
                                       self.te.data_ptr(), self.te.numel(), self.hits.data_ptr(), self.cap,
                                       self.per.data_ptr(), self.yf.data_ptr(), self.nacc.data_ptr(),
                                       self.nrej.data_ptr(), self.status.data_ptr(), self.ws.data_ptr(),
                                       _stream_ptr(stream))
        L.check(rc, "hb_cr3bp_section")

    def _rerun_overflowed(self, stream=None):
        """Trajectories the step scratch could not hold (status HB_TRAJ_RECORD_OVERFLOW): fused kernel, on the GPU."""
        if self.scratch is None or self._extra is not None:
            return
        nt = L.C.c_int64(0)
        L.check(self.lib.hb_read_record_overflow(self.ws.data_ptr(), L.C.byref(nt), _stream_ptr(stream)),
                "hb_read_record_overflow")
        if nt.value == 0:
            self._extra = (None, None)
            return
        idx = torch.nonzero(self.status[: self.n] == L.HB_TRAJ_RECORD_OVERFLOW).flatten()
        self._grow_scratch(int(nt.value))
        sub = TubeSectionRunner(idx.numel(), self.mu, self.te, self.section, forward=self.forward, flip=self.flip,
                                integ=self.integ, device=self.device)
        y0 = self._y0.view(6, self.n)[:, idx].contiguous()
        sub.launch(y0, stream)
        if self.filters is not None:                          # no records for these: judge them on stored tubes
            from . import manifold as M
            from . import propagate as P
            f = self.filters
            for a in range(0, idx.numel(), 4096):
                part = idx[a:a + 4096]
                tube = P.cr3bp_dense(self._y0.view(6, self.n)[:, part].contiguous(), self.mu, self.te,
                                     forward=self.forward, flip=self.flip, integ=self.integ, device=self.device,
                                     stream=stream, keep_on_device=True)
                q, k = M.tube_filter(tube.states, self.mu, safe_r1=f.safe_r1, safe_r2=f.safe_r2,
                                     energy_tol=f.energy_tol, device=self.device, stream=stream)
                self.filt[part], self.keep[part] = q, k
        h = sub.sorted_hits(stream)
        self.yf.view(6, self.n)[:, idx] = sub.yf.view(6, idx.numel())
        self.nacc[idx], self.nrej[idx], self.status[idx] = sub.nacc[: idx.numel()], sub.nrej[: idx.numel()], \
            sub.status[: idx.numel()]
        self.per[idx] = sub.per[: idx.numel()]
        self._extra = (idx.cpu().numpy(), h)

    def _grow_scratch(self, n_overflowed):
        """More than 2 % of the batch did not fit: size the scratch for the longest trajectory seen (the propagation
        kernel keeps counting steps past the capacity), if this runner owns its scratch and the memory is there --
        the NEXT launch then runs without reruns."""
        if not getattr(self, "_owns_scratch", False) or n_overflowed * 50 < self.n or self.steps_capacity <= 0:
            return
        need = (int(self.nacc[: self.n].max().item()) + 31) // 32 * 32
        if self.records == L.HB_RECORDS_NEAR_SECTION:         # only recorded steps count: double until it fits
            need = min(need, 2 * ((self.steps_capacity + 31) // 32 * 32))
        if need <= self.steps_capacity:
            return                                            # candidate-list overflow, not a step overflow
        nbytes = int(self.lib.hb_section2_scratch_bytes(self.n, need))
        have = self.scratch.numel() * 8
        free, _ = torch.cuda.mem_get_info(self.device)
        if nbytes > 0.8 * (free + have):
            return
        self.scratch = None
        torch.cuda.empty_cache()
        self.scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=self.device)
        self.steps_capacity = need

    def records_written(self):
        """hb_cr3bp_section2 only: step records per trajectory the last launch wrote to the scratch (int32 device tensor
        [n]) -- the accepted steps with records="all", the steps near the section plane and their neighbours with
        records="near".  Reads the scratch with the layout hb_cr3bp_section2 derives from its size."""
        if self.scratch is None or self.steps_capacity <= 0:
            raise ValueError("no step scratch")
        if self.records == L.HB_RECORDS_ALL:
            return self.nacc[: self.n].clamp(max=self._rec_cap())
        n, cap = self.n, self._rec_cap()
        off_ints = 2 * (n * cap * 64 + 2 * n * 32 * 8) + 2 * n               # past records, cand, desc, two counters
        return self.scratch.view(torch.int32)[off_ints: off_ints + n]

    def _rec_cap(self):
        fixed = (32 * 16 + 16 + 2) * 8                                        # HB_S2_FIXED_DOUBLES of hb_scan.cuh
        cap = ((self.scratch.numel() * 8 - 256) // self.n - fixed) // 512
        cap -= cap % 32
        return min(cap, 99968)

    def hit_count(self, stream=None):
        nh, no = L.C.c_int64(0), L.C.c_int64(0)
        L.check(self.lib.hb_read_hit_count(self.ws.data_ptr(), L.C.byref(nh), L.C.byref(no), _stream_ptr(stream)),
                "hb_read_hit_count")
        if no.value:
            # more hits than the buffer holds (the reference has no such cap): the workspace counted every hit, so
            # size the buffer for all of them and run the batch again -- like detect() and the fused path of tube_section
            if self._fixed_cap or self._y0 is None:
                raise L.HitenB200Error(f"hit buffer overflow: {no.value} hits dropped (capacity {self.cap})")
            self.cap = int(nh.value) + 16
            self.hits = torch.empty(self.cap * 9, dtype=torch.float64, device=self.device)
            self.launch(self._y0, stream)
            return self.hit_count(stream)
        self._main_hits = int(nh.value)
        self._rerun_overflowed(stream)
        extra = self._extra[1] if self._extra is not None else None
        return self._main_hits + (len(extra.times) if extra is not None else 0)

    def sorted_hits(self, stream=None):
        """SectionHits in the reference's order (by trajectory, then along the trajectory)."""
        self.hit_count(stream)
        k = self._main_hits
        rec = self.hits[: k * 9].cpu().numpy().view(HIT_DTYPE) if k else np.empty(0, dtype=HIT_DTYPE)
        traj, seq, t, state = rec["traj"], rec["seq"], rec["t"], rec["state"].reshape(-1, 6)
        if self._extra is not None and self._extra[1] is not None and len(self._extra[1].times):
            idx, h = self._extra
            traj = np.concatenate((traj, idx[h.trajectory_indices]))
            seq = np.concatenate((seq, _seq_within(h.trajectory_indices)))
            t = np.concatenate((t, h.times))
            state = np.concatenate((state, h.states))
        per = self.per[: self.n].cpu().numpy()
        if self.filters is not None:                          # manifold.py:412-432: discarded tubes contribute nothing
            kept = self.keep[: self.n].cpu().numpy() == 1
            sel = kept[traj]
            traj, seq, t, state = traj[sel], seq[sel], t[sel], state[sel]
            per = np.where(kept, per, 0).astype(per.dtype)
        order = np.lexsort((seq, traj))
        traj, t, state = traj[order], t[order], state[order]
        sec = self.section
        pts = np.column_stack((state[:, sec.proj_i], state[:, sec.proj_j])) if len(t) else np.empty((0, 2))
        return SectionHits(traj.copy(), t.copy(), state.copy(), pts, per)

    def filter_result(self, stream=None):
        """(quantities[n,3] = min r1, min r2, max relative Jacobi drift; keep[n]: 1 kept, 0 discarded, -1 failed
        propagation) as device tensors, once every trajectory has been judged."""
        if self.filters is None:
            raise ValueError("runner was created without filters")
        self.hit_count(stream)
        return self.filt[: self.n], self.keep[: self.n]


def _seq_within(traj):
    """0, 1, 2 ... within each run of equal (sorted) trajectory indices."""
    if len(traj) == 0:
        return np.empty(0, dtype=np.int64)
    start = np.r_[0, np.flatnonzero(np.diff(traj)) + 1]
    return np.arange(len(traj)) - np.repeat(start, np.diff(np.r_[start, len(traj)]))


@dataclass
class HostBatchResult:
    """One batch of TubeSectionStream, in pinned host memory (valid until two more batches have been drained)."""
    n_hits: int
    hits: np.ndarray            # [K] HIT_DTYPE records in the reference's order (by trajectory, then along it)
    end_states: np.ndarray      # [N, 6]
    n_acc: np.ndarray           # [N]
    n_rej: np.ndarray           # [N]
    status: np.ndarray          # [N]
    hits_per_traj: np.ndarray   # [N]


class TubeSectionStream:
    """Sweep of many equally sized host batches through tube + section (BASELINE config 5: connection sweeps):
    double-buffered so that the host->device copy of batch i+1 and the device->host copy of batch i-1 run on
    their own streams while batch i computes.  Inputs are [N, 6] host arrays (pinned for full overlap); results
    come back in pinned host buffers.  The two buffer sets share one step scratch (compute is serial anyway)."""

    def __init__(self, n, mu, t_eval, section, *, forward=1, flip=None, integ=None, hit_capacity=None, device=None,
                 steps_capacity=0, scratch=None, pool_records=8, ordered=True, records="near"):
        """Default: hb_cr3bp_section3 (pool_records step records per trajectory in the scratch); steps_capacity > 0
        (with pool_records = 0) selects hb_cr3bp_section2.  ordered=True returns every batch's hits in the reference's
        order (sorted on the device before the copy to the host)."""
        self.ordered = bool(ordered)
        _require_cuda()
        if steps_capacity > 0:
            pool_records = 0
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n = int(n)
        with torch.cuda.device(self.device):
            r0 = TubeSectionRunner(n, mu, t_eval, section, forward=forward, flip=flip, integ=integ,
                                   hit_capacity=hit_capacity, device=self.device, steps_capacity=steps_capacity,
                                   scratch=scratch, pool_records=pool_records, records=records)
            r1 = TubeSectionRunner(n, mu, t_eval, section, forward=forward, flip=flip, integ=integ,
                                   hit_capacity=hit_capacity, device=self.device, steps_capacity=steps_capacity,
                                   scratch=r0.scratch, pool_records=pool_records, records=records)
            self.runners = (r0, r1)
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
            self.d_in = [torch.empty((self.n, 6), dtype=torch.float64, device=self.device) for _ in range(2)]
            self.d_soa = [torch.empty((6, self.n), dtype=torch.float64, device=self.device) for _ in range(2)]
            self.d_yf = [torch.empty((self.n, 6), dtype=torch.float64, device=self.device) for _ in range(2)]
            pin = dict(pin_memory=True)
            self.h_hits = [torch.empty(r0.cap * 9, dtype=torch.float64, **pin) for _ in range(2)]
            self.h_yf = [torch.empty((self.n, 6), dtype=torch.float64, **pin) for _ in range(2)]
            self.h_i32 = [torch.empty((4, self.n), dtype=torch.int32, **pin) for _ in range(2)]
            self.ev_in = [torch.cuda.Event() for _ in range(2)]
            self.ev_run = [torch.cuda.Event() for _ in range(2)]
            self.ev_free = [torch.cuda.Event() for _ in range(2)]
            for e in self.ev_free:
                e.record(self.s_out)
        self.h2d_bytes = self.n * 48
        self.d2h_bytes_fixed = self.n * (48 + 16)

    def _submit(self, slot, host_batch):
        hb = host_batch if isinstance(host_batch, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(host_batch, dtype=np.float64))
        if tuple(hb.shape) != (self.n, 6):
            raise ValueError(f"batch must have shape ({self.n}, 6)")
        run = self.runners[slot]
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_run[slot])          # the previous use of this input buffer has been consumed
            self.d_in[slot].copy_(hb, non_blocking=True)
            self.ev_in[slot].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_in[slot])
            self.s_run.wait_event(self.ev_free[slot])        # outputs of this slot have left for the host
            self.d_soa[slot].copy_(self.d_in[slot].t())      # AoS -> SoA on the device
            run.launch(self.d_soa[slot], self.s_run)
            self.d_yf[slot].copy_(run.yf.view(6, self.n).t())
            self.ev_run[slot].record(self.s_run)

    def _drain(self, slot):
        run = self.runners[slot]
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_run[slot])
            self.h_yf[slot].copy_(self.d_yf[slot], non_blocking=True)
            h32 = self.h_i32[slot]
            h32[0].copy_(run.nacc[: self.n], non_blocking=True)
            h32[1].copy_(run.nrej[: self.n], non_blocking=True)
            k = run.hit_count(self.s_out)                    # waits for this batch only (s_out); reruns overflows
            h32[2].copy_(run.status[: self.n], non_blocking=True)
            h32[3].copy_(run.per[: self.n], non_blocking=True)
            km = run._main_hits
            if self.ordered and km:
                # the reference's hit order (by trajectory, then along the trajectory), restored on the device: one
                # 64-bit key sort instead of a host lexsort of the records
                rec = run.hits[: km * 9].view(km, 9)
                key = rec[:, 0].view(torch.int64) * 65536 + rec[:, 1].view(torch.int64)
                self.h_hits[slot][: km * 9].view(km, 9).copy_(rec[torch.argsort(key)], non_blocking=True)
            else:
                self.h_hits[slot][: km * 9].copy_(run.hits[: km * 9], non_blocking=True)
            self.s_out.synchronize()
            self.ev_free[slot].record(self.s_out)
        rec = self.h_hits[slot][: km * 9].numpy().view(HIT_DTYPE)
        if k != km:                                          # trajectories rerun with the fused kernel (rare)
            idx, h = run._extra
            extra = np.empty(len(h.times), dtype=HIT_DTYPE)
            extra["traj"], extra["seq"] = idx[h.trajectory_indices], _seq_within(h.trajectory_indices)
            extra["t"], extra["state"] = h.times, h.states
            rec = np.concatenate((rec, extra))
            if self.ordered:
                rec = rec[np.lexsort((rec["seq"], rec["traj"]))]
            self.h_yf[slot].copy_(run.yf.view(6, self.n).t())
        return HostBatchResult(k, rec, self.h_yf[slot].numpy(), h32[0].numpy(), h32[1].numpy(), h32[2].numpy(),
                               h32[3].numpy())

    def run(self, batches):
        """Generator: yields one HostBatchResult per input batch, in order."""
        pending = None
        for i, hb in enumerate(batches):
            slot = i & 1
            self._submit(slot, hb)
            if pending is not None:
                yield self._drain(pending)
            pending = slot
        if pending is not None:
            yield self._drain(pending)
