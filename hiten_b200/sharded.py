"""Tube + section over several GPUs: trajectories shard by index, no data-path collective, one final
gather of section hits and end states.

The reference has no multi-device notion; its natural seams are the worker pools of the synodic / centre-manifold
engines (algorithms/poincare/synodic/engine.py:92-139, centermanifold/engine.py:163-200), which split the batch into
chunks exactly like this.  Two forms:

  ShardedTubeSection       one process driving N devices (per-device runner, workspace and stream; peer copies to
                           device 0 at the end) -- what a user of the drop-in gets on a multi-GPU box;
  DistributedTubeSection   one process per GPU under torch.distributed (NCCL), the launch bench.py is run with:
                           rank r owns trajectories r, r + W, r + 2W, ...; hits are gathered to rank 0 with a padded
                           NCCL gather after an all-gather of the counts.

Sharding is interleaved (i mod W) so that a tube's phase-dependent cost spreads evenly (8e).  Both return hits with
GLOBAL trajectory indices in the reference's order (by trajectory, then along the trajectory).
"""
import numpy as np
import torch

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


class DistributedTubeSection:
    """One process per GPU (torch.distributed initialised by the caller, NCCL on GPUs / gloo in the CPU tests).
    `runner_factory(n_local)` builds the rank's runner -- by default a TubeSectionRunner on the current device."""

    def __init__(self, n_global, mu, t_eval, section, *, forward=1, flip=None, integ=None, steps_capacity=192,
                 pool_records=0, hit_capacity=None, runner_factory=None, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
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

    def launch(self, y0_soa_local, stream=None):
        self.runner.launch(y0_soa_local, stream)

    def gather_device(self):
        """The one exchange of the path, on the device, stream-ordered after launch(): end states + per-trajectory hit
        counts (fixed size) and the hit records (all-gather of the counts, then a gather padded to the largest) to rank 0.
        Returns (hit record tensors per rank | None, counts, end-state tensors per rank | None) without host work."""
        dist, run = self.dist, self.runner
        dev = run.yf.device
        k = torch.tensor([run.hit_count()], dtype=torch.int64, device=dev)      # also completes overflow reruns
        if self.world == 1:
            return [run.hits[: int(k.item()) * 9]], k, [run.yf]
        counts = [torch.zeros_like(k) for _ in range(self.world)]
        dist.all_gather(counts, k, group=self.group)
        kmax = int(max(int(c.item()) for c in counts))
        send = torch.zeros(max(kmax, 1) * 9, dtype=torch.float64, device=dev)
        km = min(int(k.item()), run._main_hits)
        send[: km * 9] = run.hits[: km * 9]
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
