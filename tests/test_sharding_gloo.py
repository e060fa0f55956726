"""CPU, world_size 2 over gloo: the index sharding + final gather used by bench.py --gpus N.

The data path has no collective (trajectories are independent); the only exchange is the gather of end states
and per-trajectory hit counts to rank 0.  Here each rank propagates its shard with the CPU oracle (the GPU kernels
are covered by the -m gpu tests) and rank 0 checks that the gathered, re-interleaved result equals the unsharded run.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hiten_b200 import workloads as W
    import oracle_lib as O
    n = 24
    ics, mu = W.c1_tube_batch(n, rank, world)                     # ics_all[rank::world][:n]
    s = O.system(O.SYS_CR3BP6, mu, fwd=-1, flip=(0, 6))
    yf, counts = O.batch_final(s, O.DOP853, O.default_tol(), ics, 0.0, 1.0, 1)
    t_yf = torch.from_numpy(np.ascontiguousarray(yf.T))            # [6, n] like the kernels' SoA output
    t_steps = torch.from_numpy(counts.sum(axis=1).astype(np.int32))
    g_yf = [torch.empty_like(t_yf) for _ in range(world)] if rank == 0 else None
    g_st = [torch.empty_like(t_steps) for _ in range(world)] if rank == 0 else None
    dist.gather(t_yf, g_yf, dst=0)
    dist.gather(t_steps, g_st, dst=0)
    tt = torch.tensor([0.5 + rank, float(t_steps.sum())], dtype=torch.float64)
    tmax, tsum = tt.clone(), tt.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)                    # time: max over ranks
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)                    # work: sum over ranks
    if rank == 0:
        full = np.empty((n * world, 6))
        for r in range(world):
            full[r::world] = g_yf[r].numpy().T
        np.savez(out_path, yf=full, steps=int(tsum[1].item()), tmax=float(tmax[0].item()))
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    sys.path.insert(0, REPO)
    from hiten_b200 import workloads as W
    import oracle_lib as O
    O.build()
    world, n = 2, 24
    out = str(tmp_path / "gathered.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    ics, mu = W.c1_tube_batch(n * world)                           # the unsharded batch
    s = O.system(O.SYS_CR3BP6, mu, fwd=-1, flip=(0, 6))
    yf, counts = O.batch_final(s, O.DOP853, O.default_tol(), ics, 0.0, 1.0, 1)
    assert np.array_equal(got["yf"], yf)                           # shards re-interleave to the unsharded result
    assert int(got["steps"]) == int(counts.sum())
    assert float(got["tmax"]) == 1.5                               # MAX over ranks, not rank 0's time


def test_shards_partition_the_batch():
    sys.path.insert(0, REPO)
    from hiten_b200 import workloads as W
    full, _ = W.c1_tube_batch(64)
    parts = [W.c1_tube_batch(16, r, 4)[0] for r in range(4)]
    rebuilt = np.empty_like(full)
    for r in range(4):
        rebuilt[r::4] = parts[r]
    assert np.array_equal(rebuilt, full)


# ---- the product API (hiten_b200.sharded.DistributedTubeSection) with an oracle-backed runner, world_size 2 ----------
class _OracleRunner:
    """Stands in for synodic.TubeSectionRunner on a box without a GPU: same launch() / sorted_hits() / yf surface,
    computed by the CPU oracle.  Only the sharding / gathering logic of hiten_b200.sharded is under test here."""

    def __init__(self, n, mu, t_eval, sec_args, forward):
        self.n, self.mu, self.t_eval, self.sec_args, self.forward = n, mu, t_eval, sec_args, forward

    def launch(self, y0_soa, stream=None):
        import oracle_lib as O
        from hiten_b200.synodic import SectionHits
        x0 = np.ascontiguousarray(y0_soa.numpy().T)
        s = O.system(O.SYS_CR3BP6, self.mu, fwd=self.forward, flip=(0, 6))
        dense, _ = O.batch_dense(s, O.DOP853, O.default_tol(), x0, self.t_eval, 1)
        self.yf = torch.from_numpy(np.ascontiguousarray(dense[:, -1, :].T))
        idx, off, direction, proj = self.sec_args
        tr, tt, xx = [], [], []
        for i in range(len(dense)):
            t, x = O.synodic_detect(self.forward * self.t_eval, dense[i], idx, off, direction, proj)
            tr += [i] * len(t); tt += list(t); xx += list(x)
        st = np.array(xx).reshape(-1, 6)
        per = np.bincount(np.array(tr, dtype=np.int64), minlength=self.n).astype(np.int32)
        self._hits = SectionHits(np.array(tr, dtype=np.int64), np.array(tt), st, st[:, list(proj)], per)

    def sorted_hits(self, stream=None):
        return self._hits


def _sharded_worker(rank, world, port, out_path):
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hiten_b200 import sharded, workloads as W
    from hiten_b200._lib import HbSection
    n = 15                                                        # odd: the shards differ in size
    ics, mu = W.c5_batch(2 * n)
    x0 = ics["l1"]
    t_eval = W.c5_grid("l1")
    sec = HbSection(1, 0, 0.0, 0, 2, 50, 0, 1e-6, 1e-9, 1e-6)        # y = 0, (x, z), both directions
    d = sharded.DistributedTubeSection(n, mu, t_eval, sec, forward=-1, flip=(0, 6),
                                       runner_factory=lambda nl: _OracleRunner(nl, mu, t_eval, (1, 0.0, 0, (0, 2)), -1))
    assert np.array_equal(d.index, np.arange(rank, n, world))
    d.launch(torch.from_numpy(np.ascontiguousarray(x0[d.index].T)))
    hits, yf = d.gather()
    if rank == 0:
        np.savez(out_path, traj=hits.trajectory_indices, t=hits.times, state=hits.states, per=hits.hits_per_traj, yf=yf)
    else:
        assert hits is None and yf is None
    dist.destroy_process_group()


def test_distributed_tube_section_two_ranks_equals_one(tmp_path):
    """hits gathered from two interleaved shards, with global trajectory indices in the reference's order, equal the
    unsharded run; so do the end states."""
    sys.path.insert(0, REPO)
    import oracle_lib as O
    from hiten_b200 import workloads as W
    O.build()
    out = str(tmp_path / "sharded.npz")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_sharded_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    n = 15
    ics, mu = W.c5_batch(2 * n)
    t_eval = W.c5_grid("l1")
    one = _OracleRunner(n, mu, t_eval, (1, 0.0, 0, (0, 2)), -1)
    one.launch(torch.from_numpy(np.ascontiguousarray(ics["l1"].T)))
    h = one.sorted_hits()
    assert len(h.times) > 0
    assert np.array_equal(got["traj"], h.trajectory_indices) and np.array_equal(got["t"], h.times)
    assert np.array_equal(got["state"], h.states) and np.array_equal(got["per"], h.hits_per_traj)
    assert np.array_equal(got["yf"], one.yf.numpy().T)
