"""Tube + section over several GPUs: trajectories shard by index, no data-path collective, one final
gather of section hits and end states.

The reference has no multi-device notion; its natural seams are the worker pools of the synodic / centre-manifold
engines (algorithms/poincare/synodic/engine.py:92-139, centermanifold/engine.py:163-200), which split the batch into
chunks exactly like this.  Two forms:

  ShardedTubeSection       one process driving N devices (per-device runner, workspace and stream; peer copies to
                           device 0 at the end) -- what a user of the drop-in gets on a multi-GPU box;
  DistributedTubeSection   one process per GPU under torch.distributed (NCCL), the launch bench.py is run with:
                           rank r owns trajectories r, r + W, r + 2W, ...; the final exchange writes every rank's hit
                           records and end states straight into rank 0's receive buffer over NVLink peer memory with the
                           COPY ENGINES (PeerExchange: symmetric memory, no SM work, so the transfer of one tube runs
                           under the propagation of the next) or, for small shards, with the hb_peer_put KERNEL, which
                           reads the hit count on the device and needs no host wait between pipeline and exchange;
                           without peer memory it is a padded NCCL gather after an all-gather of the counts.

Sharding is interleaved (i mod W) so that a tube's phase-dependent cost spreads evenly (8e).  Both return hits with
GLOBAL trajectory indices in the reference's order (by trajectory, then along the trajectory).
"""
import numpy as np
import torch

from . import _lib as _L
from . import synodic as _syn


def shard_indices(n, world, rank):
    """Global trajectory indices of one shard: rank, rank + world, ..."""
    return np.arange(rank, n, world)


def merge_shard_hits(parts, world, section):
    """parts[r] = (local trajectory index [K_r], seq [K_r], t [K_r], state [K_r, 6]) of shard r ->
    SectionHits-like tuple with global indices, ordered by (trajectory, seq)."""
    traj = np.concatenate([np.asarray(p[0], dtype=np.int64) * world + r for r, p in enumerate(parts)])
    seq = np.concatenate([np.asarray(p[1], dtype=np.int64) for p in parts])
    t = np.concatenate([np.asarray(p[2], dtype=np.float64) for p in parts])
    state = np.concatenate([np.asarray(p[3], dtype=np.float64).reshape(-1, 6) for p in parts])
    order = np.lexsort((seq, traj))
    traj, t, state = traj[order], t[order], state[order]
    pts = np.column_stack((state[:, section.proj_i], state[:, section.proj_j])) if len(t) else np.empty((0, 2))
    return traj, t, state, pts


class ShardedTubeSection:
    """One process, several devices.  launch() enqueues every device's pipeline on that device's own stream and returns
    at once; gather() waits, rerun-completes each shard and merges the results on the host (hit records and end states
    come over PCIe from every device in parallel)."""

    def __init__(self, n, mu, t_eval, section, *, forward=1, flip=None, integ=None, devices=None, steps_capacity=192,
                 pool_records=0, hit_capacity=None):
        _syn._require_cuda()
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        self.devices = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in devices]
        self.world = len(self.devices)
        if self.world < 1:
            raise ValueError("no devices")
        self.n, self.section = int(n), section
        self.index = [shard_indices(self.n, self.world, r) for r in range(self.world)]
        self.runners, self.streams = [], []
        for r, dev in enumerate(self.devices):
            with torch.cuda.device(dev):
                self.runners.append(_syn.TubeSectionRunner(len(self.index[r]), mu, t_eval, section, forward=forward,
                                                           flip=flip, integ=integ, device=dev,
                                                           steps_capacity=steps_capacity, pool_records=pool_records,
                                                           hit_capacity=hit_capacity))
                self.streams.append(torch.cuda.Stream(dev))
        self._y0 = [None] * self.world

    def launch(self, y0):
        """y0: host array [N, 6] (each shard is copied to its device) or a list of per-device SoA tensors [6, n_r]."""
        for r, dev in enumerate(self.devices):
            with torch.cuda.device(dev), torch.cuda.stream(self.streams[r]):
                if isinstance(y0, (list, tuple)):
                    y = y0[r]
                else:
                    part = np.ascontiguousarray(np.asarray(y0, dtype=np.float64)[self.index[r]].T)
                    y = torch.from_numpy(part).to(dev, non_blocking=True)
                self._y0[r] = y
                self.runners[r].launch(y, self.streams[r])

    def gather(self):
        """-> (SectionHits with global trajectory indices, end states [N, 6], status [N])."""
        parts, yf, status, per = [], np.empty((self.n, 6)), np.empty(self.n, dtype=np.int32), \
            np.zeros(self.n, dtype=np.int32)
        for r, dev in enumerate(self.devices):
            with torch.cuda.device(dev), torch.cuda.stream(self.streams[r]):
                run = self.runners[r]
                h = run.sorted_hits(self.streams[r])
                parts.append((h.trajectory_indices, _syn._seq_within(h.trajectory_indices), h.times, h.states))
                yf[self.index[r]] = run.yf.t().cpu().numpy()
                status[self.index[r]] = run.status.cpu().numpy()
                per[self.index[r]] = h.hits_per_traj[: len(self.index[r])]
        traj, t, state, pts = merge_shard_hits(parts, self.world, self.section)
        return _syn.SectionHits(traj, t, state, pts, per), yf, status


class PeerExchange:
    """Rank 0's receive buffer, mapped into every rank's address space (torch symmetric memory over NVLink / NVSwitch).
    Rank r owns slot r: [8 doubles of header | hit records | end states].  A rank fills its slot with plain device-to-
    device copies on its own copy stream -- cudaMemcpyAsync on a peer-mapped pointer, i.e. the copy engines: no SM is
    needed, so the copies of one tube overlap the persistent propagation kernel of the next, which owns every register
    of every SM.  finish() is the only synchronisation: a stream-ordered barrier over the group's signal pads.  Every
    put() must be followed by finish() on all ranks before the next put(); the tensors received() returns are views of
    the receive buffer and are overwritten by the next exchange."""

    HEADER = 8

    def __init__(self, hit_slots, n_max, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.hit_slots, self.n_max = int(hit_slots) + int(hit_slots) % 2, int(n_max)     # even: 16-byte aligned slots
        self.slot = self.HEADER + 9 * self.hit_slots + 6 * self.n_max
        with torch.cuda.device(device):
            self.buf = symm.empty(self.world * self.slot, dtype=torch.float64, device=device)
            self.hdl = symm.rendezvous(self.buf, self.group)
            self.mine = self.hdl.get_buffer(0, (self.slot,), torch.float64, self.rank * self.slot)   # my slot on rank 0
            self.stream = torch.cuda.Stream(device)
            self.hdr_host = torch.zeros(self.HEADER, dtype=torch.float64).pin_memory()
            self.hdr_dev = torch.zeros(self.HEADER, dtype=torch.float64, device=device)
            self.cnt_host = torch.zeros(4, dtype=torch.int64).pin_memory()
            # my slot in EVERY rank's buffer (hb_peer_put writes its header to all of them)
            self.mine_all = [self.hdl.get_buffer(r, (self.slot,), torch.float64, self.rank * self.slot)
                             for r in range(self.world)]
            self.slot_ptrs = (_L.C.c_void_p * self.world)(*[t.data_ptr() for t in self.mine_all])
            self.hdr_all_host = torch.zeros((self.world, self.HEADER), dtype=torch.float64).pin_memory()
            self.hdr_all_dev = torch.zeros((self.world, self.HEADER), dtype=torch.float64, device=device)

    def put_device(self, hits, yf_flat, n_local, ws, stream):
        """The same transfer as put() done by a kernel on `stream` (hb_peer_put): the hit count is read on the device, so
        no host wait is needed between the pipeline and the exchange; the closing barrier and the read-back of the
        headers all shards wrote into THIS rank's buffer are enqueued behind it.  Small shards (see
        DistributedTubeSection)."""
        lib = _L.load()
        _L.check(lib.hb_peer_put(self.slot_ptrs, self.world, 0, hits.data_ptr(), self.hit_slots, yf_flat.data_ptr(),
                                 int(n_local), ws.data_ptr(), _L.vp(stream.cuda_stream)), "hb_peer_put")
        with torch.cuda.stream(stream):
            self.hdl.barrier(channel=0)
            self.hdr_all_dev.copy_(self.buf.view(self.world, self.slot)[:, : self.HEADER])
            self.hdr_all_host.copy_(self.hdr_all_dev, non_blocking=True)

    def finish_device(self, stream):
        """Closes an exchange started with put_device on `stream` (the one host wait of that form).  -> host tensor
        [world, 8] on every rank: {hits, n_local, dropped, record overflows, sendable, ...} of every shard."""
        stream.synchronize()
        return self.hdr_all_host

    def put(self, k, hits, yf_flat):
        """Enqueue the copies of this rank's slot on the copy stream (k hit records, the end states)."""
        with torch.cuda.stream(self.stream):
            self.hdr_host[0], self.hdr_host[1] = float(k), float(yf_flat.numel() // 6)
            self.hdr_dev.copy_(self.hdr_host, non_blocking=True)
            self.mine[: self.HEADER].copy_(self.hdr_dev, non_blocking=True)
            if k:
                self.mine[self.HEADER: self.HEADER + 9 * k].copy_(hits[: 9 * k], non_blocking=True)
            o = self.HEADER + 9 * self.hit_slots
            self.mine[o: o + yf_flat.numel()].copy_(yf_flat, non_blocking=True)

    def finish(self):
        with torch.cuda.stream(self.stream):
            self.hdl.barrier(channel=0)
        self.stream.synchronize()

    def received(self, hdr=None):
        """Rank 0: (hit record tensors per rank, counts, end-state tensors per rank) as views of the receive buffer."""
        rows = self.buf.view(self.world, self.slot)
        hdr = rows[:, : self.HEADER].cpu() if hdr is None else hdr
        hits, yfs, counts = [], [], []
        o = self.HEADER + 9 * self.hit_slots
        for r in range(self.world):
            k, nl = int(hdr[r, 0].item()), int(hdr[r, 1].item())
            counts.append(k)
            hits.append(rows[r, self.HEADER: self.HEADER + 9 * k])
            yfs.append(rows[r, o: o + 6 * nl].view(6, nl))
        return hits, torch.tensor(counts, dtype=torch.int64), yfs


class DistributedTubeSection:
    """One process per GPU (torch.distributed initialised by the caller, NCCL on GPUs / gloo in the CPU tests).
    `runner_factory(n_local)` builds the rank's runner -- by default a TubeSectionRunner on the current device."""

    def __init__(self, n_global, mu, t_eval, section, *, forward=1, flip=None, integ=None, steps_capacity=192,
                 pool_records=0, hit_capacity=None, runner_factory=None, group=None, exchange="auto",
                 peer_hit_slots=None, device_put="auto"):
        """exchange: "peer" (PeerExchange), "nccl" (padded gather) or "auto" (peer when the process group runs on NCCL
        and symmetric memory can be set up, else nccl; HITEN_B200_EXCHANGE overrides).  peer_hit_slots: hit records per
        rank the receive buffer holds (default 4 per trajectory; more hits than that raise).
        device_put (peer exchange only): True -- the shard is written into rank 0's buffer by a kernel that reads the hit
        count on the device (hb_peer_put: no host wait between pipeline and exchange); False -- by the copy engines after
        the host has read the count (overlaps a later persistent launch that owns the SMs); "auto" -- the kernel for
        shards below 8 trajectories per lane of a full-device launch (where the host round trips are a visible share of
        the step and the SMs are idle when the pipeline ends), HITEN_B200_PEER_PUT=kernel|copy overrides."""
        import os
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self._exchange = os.environ.get("HITEN_B200_EXCHANGE", exchange)
        self.px = None
        self._device_put, self._pending = os.environ.get("HITEN_B200_PEER_PUT", device_put), None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_global, self.section = int(n_global), section
        self.index = shard_indices(self.n_global, self.world, self.rank)
        if runner_factory is None:
            def runner_factory(n_local):
                return _syn.TubeSectionRunner(n_local, mu, t_eval, section, forward=forward, flip=flip, integ=integ,
                                              steps_capacity=steps_capacity, pool_records=pool_records,
                                              hit_capacity=hit_capacity)
        self.runner = runner_factory(len(self.index))
        # shards differ by at most one trajectory: pad the end-state exchange to the largest
        self.n_max = (self.n_global + self.world - 1) // self.world
        if self.world > 1 and self._exchange in ("auto", "peer") and hasattr(self.runner, "yf") \
                and getattr(self.runner.yf, "is_cuda", False):
            try:
                slots = int(peer_hit_slots) if peer_hit_slots else max(1024, 4 * self.n_max)
                self.px = PeerExchange(slots, self.n_max, self.runner.yf.device, group)
            except Exception as exc:                       # no peer mapping on this box / backend: NCCL gather instead
                if self._exchange == "peer":
                    raise
                self.px = None
                self.peer_error = repr(exc)
            # the exchange is collective: every rank must have come to the same choice
            ok = torch.tensor([1 if self.px is not None else 0], dtype=torch.int32, device=self.runner.yf.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.px = None
            if self.px is not None:
                dp = self._device_put
                if dp in ("kernel", "copy"):
                    dp = dp == "kernel"
                elif dp == "auto":
                    sms = torch.cuda.get_device_properties(self.runner.yf.device).multi_processor_count
                    dp = self.n_max < 8 * sms * 256
                self._device_put = bool(dp) and self.world <= 16

    def launch(self, y0_soa_local, stream=None):
        self.runner.launch(y0_soa_local, stream)

    def start_gather(self, stream=None, device_put=None):
        """Peer form of the exchange, first half.  Copy-engine form: wait (host) for THIS runner's pipeline only -- later
        launches on the same stream keep running --, read its hit counter, and enqueue the copies into rank 0's receive
        buffer on the copy stream.  Kernel form (device_put, small shards): enqueue hb_peer_put behind the pipeline on its
        own stream; nothing waits.  Returns False when the peer path is not available (use gather_device)."""
        if self.px is None:
            return False
        run, px = self.runner, self.px
        main = torch.cuda.current_stream(run.yf.device) if stream is None else stream
        if self._device_put if device_put is None else device_put:
            px.put_device(run.hits, run.yf.view(-1), len(self.index), run.ws, main)
            self._pending = main
            return True
        ev = getattr(run, "done_event", None)
        if ev is None:
            ev = torch.cuda.Event()
            ev.record(main)
        px.stream.wait_event(ev)
        with torch.cuda.stream(px.stream):
            px.cnt_host.copy_(run.ws[:4], non_blocking=True)      # cursor, hit_count, overflow, rec_overflow
        px.stream.synchronize()
        k, dropped, rec_over = int(px.cnt_host[1]), int(px.cnt_host[2]), int(px.cnt_host[3])
        hits = run.hits
        if dropped or rec_over:
            # rare: the hit buffer was too small (regrow + relaunch) or trajectories outgrew the step scratch (rerun with
            # the fused kernel) -- completed on the main stream; the rerun trajectories' records travel with the rest
            with torch.cuda.stream(main):
                k, hits = self._all_records(main)
            px.stream.wait_stream(main)
        else:
            run._main_hits, run._extra = k, (None, None)
        if k > px.hit_slots:
            from ._lib import HitenB200Error
            raise HitenB200Error(f"peer exchange: {k} hits on rank {self.rank} exceed the {px.hit_slots} record slots of its "
                                 "share of rank 0's receive buffer; pass a larger peer_hit_slots")
        px.put(k, hits, run.yf.view(-1)[: 6 * len(self.index)])
        return True

    def _all_records(self, stream=None):
        """(k, device tensor holding k 72-byte records): this shard's hits including those of the trajectories that were
        rerun with the fused kernel (kept on the host by the runner)."""
        run = self.runner
        k = run.hit_count(stream)
        km = run._main_hits
        extra = run._extra if run._extra is not None else (None, None)
        if extra[1] is None or not len(extra[1].times):
            return km, run.hits
        idx, h = extra
        rec = np.empty(len(h.times), dtype=_syn.HIT_DTYPE)
        rec["traj"], rec["seq"] = idx[h.trajectory_indices], _syn._seq_within(h.trajectory_indices)
        rec["t"], rec["state"] = h.times, h.states
        ext = torch.from_numpy(rec.view(np.float64).copy()).to(run.hits.device)
        return k, torch.cat((run.hits[: km * 9], ext))

    def finish_gather(self):
        """Second half: all ranks' copies have landed on rank 0.  -> (hit record tensors per rank | None, counts | None,
        end-state tensors per rank | None)."""
        hdr = None
        if self._pending is not None:
            main, self._pending = self._pending, None
            hdr = self.px.finish_device(main)               # [world, 8] on every rank: all take the same branch below
            if bool((hdr[:, 4] != 1.0).any()):
                # some shard dropped hits, overflowed its step scratch or its slot: second, host-sized round (reruns
                # and regrown buffers are completed there; too many hits for the slot raise as in the copy form)
                with torch.cuda.stream(main):
                    self.start_gather(main, device_put=False)
                self.px.finish()
                hdr = None
            else:
                self.runner._main_hits, self.runner._extra = int(hdr[self.rank, 0]), (None, None)
        else:
            self.px.finish()
        if self.rank != 0:
            return None, None, None
        return self.px.received(hdr)

    def gather_device(self):
        """The one exchange of the path, on the device, stream-ordered after launch(): end states + per-trajectory hit
        counts (fixed size) and the hit records to rank 0 -- over peer memory when available (start_gather +
        finish_gather), else an all-gather of the counts and a gather padded to the largest shard.
        Returns (hit record tensors per rank | None, counts, end-state tensors per rank | None) without host work."""
        if self.world > 1 and self.start_gather():
            return self.finish_gather()
        dist, run = self.dist, self.runner
        dev = run.yf.device
        kk, recs = self._all_records()                                          # also completes overflow reruns
        k = torch.tensor([kk], dtype=torch.int64, device=dev)
        if self.world == 1:
            return [recs[: kk * 9]], k, [run.yf]
        counts = [torch.zeros_like(k) for _ in range(self.world)]
        dist.all_gather(counts, k, group=self.group)
        kmax = int(max(int(c.item()) for c in counts))
        send = torch.zeros(max(kmax, 1) * 9, dtype=torch.float64, device=dev)
        send[: kk * 9] = recs[: kk * 9]
        yf = torch.zeros((6, self.n_max), dtype=torch.float64, device=dev)
        yf[:, : len(self.index)] = run.yf.view(6, -1)
        if self.rank == 0:
            hit_bufs = [torch.empty_like(send) for _ in range(self.world)]
            yf_bufs = [torch.empty_like(yf) for _ in range(self.world)]
        else:
            hit_bufs = yf_bufs = None
        dist.gather(send, hit_bufs, dst=0, group=self.group)
        dist.gather(yf, yf_bufs, dst=0, group=self.group)
        return hit_bufs, torch.cat(counts), yf_bufs

    def gather(self):
        """Host form: rank 0 gets (SectionHits with global indices, end states [N, 6]); other ranks (None, None)."""
        run = self.runner
        h = run.sorted_hits()                                                    # local, ordered, reruns included
        local = (h.trajectory_indices, _syn._seq_within(h.trajectory_indices), h.times, h.states)
        yf_local = run.yf.t().cpu().numpy() if hasattr(run.yf, "cpu") else np.asarray(run.yf)
        if self.world == 1:
            traj, t, state, pts = merge_shard_hits([local], 1, self.section)
            return _syn.SectionHits(traj, t, state, pts, h.hits_per_traj), yf_local
        gathered = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object((local, yf_local), gathered, dst=0, group=self.group)
        if self.rank != 0:
            return None, None
        traj, t, state, pts = merge_shard_hits([g[0] for g in gathered], self.world, self.section)
        yf = np.empty((self.n_global, 6))
        per = np.zeros(self.n_global, dtype=np.int32)
        for r, g in enumerate(gathered):
            idx = shard_indices(self.n_global, self.world, r)
            yf[idx] = g[1]
        np.add.at(per, traj, 1)
        return _syn.SectionHits(traj, t, state, pts, per), yf
