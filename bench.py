#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native HITEN propagation hot path.

Workload = BASELINE.json configs[4] as the reference runs it (examples/heteroclinic_connection.py:29-63):
the stable-manifold tube of the Earth-Moon L1 halo (Az = 0.5 S; backward, 0.9 x 2 pi) and the unstable-manifold tube of
the L2 halo (Az = 0.3663368 N; forward, 2 pi), each 2000 orbit nodes x displacements log-spaced in [1e-7, 1e-5], every
trajectory propagated with DOP853 at rtol = atol = 1e-12, dense samples on Manifold.compute()'s dt = 1e-3 grid streamed
through the synodic detector for the section x = 1 - mu / (y, z) / direction -1 (+1 for the stable tube, the flip of
connections/interfaces.py:350).  1e6 trajectories per GPU and step (5e5 per tube); hits + end states out.
Parity with the reference on this geometry: tests/test_gpu_c5.py (bit for bit against tests/golden/c5_connection.npz).

A "step" = one pass of the hot path over the per-GPU batch (both tubes).  metric = fp64 CR3BP RK steps/s (attempted
DOP853 steps, accepted + rejected, whole job); crossings/s is reported beside it.  N > 1: one rank per GPU, interleaved
index shards (weak scaling: 1e6 per GPU), no data-path collective, one NCCL gather of hit records, hit counts and end
states per tube at the end of a step; extra.strong_scaling times configs[4]'s 1e6 TOTAL over the N GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--arith parity|fast] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "tests", "golden")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

FLOP_PER_STEP = 1350.0          # algorithmic flop per attempted 6-state DOP853 step
FLOP_PER_SEGMENT = 1050.0       # dense-output cache of one accepted step (3 RHS + D/A_ext rows)
N_PER_GPU = 1_000_000           # BASELINE configs[4]: 1e6 manifold trajectories (both tubes together)
STEPS_CAPACITY = 128            # RECORDED steps per trajectory the hb_cr3bp_section2 scratch holds (sparse records:
                                # a C5 trajectory writes ~11 of its ~100 accepted steps; 183 with --records all)
TUBES = ("l1", "l2")
WORKLOAD = ("C5 = BASELINE configs[4] as examples/heteroclinic_connection.py runs it: EM L1 halo (Az=0.5 S) stable tube "
            "(backward, 0.9*2pi, 5655 samples) + EM L2 halo (Az=0.3663368 N) unstable tube (forward, 2pi, 6284 samples), "
            "2000 orbit nodes x log-spaced displacements [1e-7,1e-5] each, DOP853 rtol=atol=1e-12, dense samples on the "
            "dt=1e-3 grid streamed through the synodic detector x=1-mu / (y,z) / direction -1 (+1 for the stable tube), "
            "segment_refine=50; hits + end states out")


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}




# ---------------------------------------------------------------------------------------------------------------------
# CPU arms: the oracle port (all host threads) and the reference's own code (oracle/_ref, when it shipped)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_step(ics, mu, n_threads):
    """The same step on the CPU oracle (C restatement): dense tube on the dt = 1e-3 grid + section detection, both
    tubes.  Returns (RK steps, hits)."""
    import oracle_lib as O
    from hiten_b200 import workloads as W
    steps, hits = 0, 0
    for key in TUBES:
        fwd = W.C5_TUBES[key]["forward"]
        s = O.system(O.SYS_CR3BP6, mu, fwd=fwd, flip=(0, 6))
        t_eval = W.c5_grid(key)
        chunk = 512                                            # 512 x 6284 x 48 B = 154 MB of dense output at a time
        x = ics[key]
        for a in range(0, len(x), chunk):
            dense, c = O.batch_dense(s, O.DOP853, O.default_tol(), x[a:a + chunk], t_eval, n_threads)
            steps += int(c.sum())
            hits += O.batch_synodic_count(fwd * t_eval, dense, 0, 1.0 - mu, W.C5_TUBES[key]["direction"], (1, 2), 50,
                                          1e-6, 1e-9, 1e-6, n_threads)
    return steps, hits


def cpu_port_throughput(ics, mu, n_threads, seconds_target=10.0):
    """Bounded sample of the step on the oracle port, all host threads."""
    take = lambda k: {key: ics[key][:k] for key in TUBES}
    t0 = time.perf_counter()
    cpu_port_step(take(128), mu, n_threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    n = int(min(len(ics["l1"]), max(256, 128 / dt * seconds_target)))
    t0 = time.perf_counter()
    steps, _ = cpu_port_step(take(n), mu, n_threads)
    dt = time.perf_counter() - t0
    return steps / dt, 2 * n, dt


class ReferenceArm:
    """The reference's OWN code on this step: `_propagate_dynsys` per initial condition, exactly as the fraction loop of
    `_ManifoldDynamicsService._run_compute` calls it (services/manifold.py:399-409), then `_SynodicDetectionBackend.run`
    on the tube (synodic/backend.py:823) with the request SynodicMap.compute() builds.  The package is imported from
    oracle/_ref (unmodified copy, oracle/build_ref.sh) -- /root/reference does not exist on the GPU box.  The reference
    has no parallel driver for the propagation loop and its detection is a Python loop under the GIL: 1 core."""

    def __init__(self):
        import _refenv
        self.src = _refenv.enable()
        from hiten import System
        from hiten.algorithms.dynamics.base import _propagate_dynsys
        from hiten.algorithms.poincare.synodic.backend import _SynodicDetectionBackend
        from hiten.algorithms.poincare.synodic.types import SynodicBackendRequest
        self.system = System.from_bodies("earth", "moon")
        self.prop, self.backend, self.Request = _propagate_dynsys, _SynodicDetectionBackend(), SynodicBackendRequest

    def step(self, ics, mu):
        """-> (number of hits, seconds).  RK step counts come from the oracle (bit-identical controller)."""
        from hiten_b200 import workloads as W
        hits = 0
        t0 = time.perf_counter()
        for key in TUBES:
            tube = W.C5_TUBES[key]
            t_eval = W.c5_grid(key)
            trajs = []
            for x0 in ics[key]:
                sol = self.prop(dynsys=self.system.dynsys, state0=x0, t0=0.0, tf=float(t_eval[-1]), forward=tube["forward"],
                                steps=len(t_eval), method="adaptive", order=8, flip_indices=slice(0, 6))
                trajs.append((sol.times, sol.states))
            normal = np.zeros(6)
            normal[0] = 1.0
            req = self.Request(trajectories=trajs, normal=normal, trajectory_indices=list(range(len(trajs))),
                               offset=1.0 - mu, plane_coords=("y", "z"), interp_kind="linear", segment_refine=50,
                               tol_on_surface=1e-6, dedup_time_tol=1e-9, dedup_point_tol=1e-6, max_hits_per_traj=None,
                               newton_max_iter=10, direction=tube["direction"])
            hits += len(self.backend.run(req).points)
        return hits, time.perf_counter() - t0


_FORK_ARM = None          # ReferenceArm of the parent, inherited by forked workers (their Numba code is already compiled)


def _ref_worker(job):
    shard, mu = job
    return _FORK_ARM.step(shard, mu)[0]


def reference_all_cores(arm, ics_all, mu, per_tube_per_worker, steps, n_procs):
    """The reference's own code on every host core: it has no parallel driver for this loop (and its @njit kernels hold the
    GIL), so the batch is sharded by index over `n_procs` FORKED worker processes -- each runs exactly ReferenceArm.step
    on its shard ("run warm worker processes over index shards and state the core count").  Forked after the
    parent's warm-up call, so no worker compiles anything.  -> (RK steps/s, hits/s, trajectories per step, seconds)."""
    global _FORK_ARM
    import multiprocessing as mp
    _FORK_ARM = arm
    k = per_tube_per_worker * n_procs
    sample = {key: ics_all[key][:: max(1, len(ics_all[key]) // k)][:k] for key in TUBES}
    shards = [({key: sample[key][w::n_procs] for key in TUBES}, mu) for w in range(n_procs)]
    rk_steps = oracle_step_count(sample, mu)
    with mp.get_context("fork").Pool(n_procs) as pool:
        pool.map(_ref_worker, [({key: sample[key][:1] for key in TUBES}, mu)] * n_procs)      # every worker is up
        t0 = time.perf_counter()
        hits = 0
        for _ in range(steps):
            hits += sum(pool.map(_ref_worker, shards, chunksize=1))
        dt = time.perf_counter() - t0
    return rk_steps * steps / dt, hits / dt, 2 * k, dt


def oracle_step_count(ics, mu):
    import oracle_lib as O
    from hiten_b200 import workloads as W
    total = 0
    for key in TUBES:
        s = O.system(O.SYS_CR3BP6, mu, fwd=W.C5_TUBES[key]["forward"], flip=(0, 6))
        _, c = O.batch_final(s, O.DOP853, O.default_tol(), ics[key], 0.0, float(W.c5_grid(key)[-1]), 4)
        total += int(c.sum())
    return total


def reference_available():
    import _refenv
    return _refenv.available() is not None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step (oracle/_ref) on a bounded sample of the
    same batch; the oracle port (all host threads) when the reference package did not ship."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O
    from hiten_b200 import workloads as W
    ics_all, mu = W.c5_batch(N_PER_GPU)
    n_threads = O.lib().ho_max_threads()
    if reference_available():
        arm = ReferenceArm()
        per_tube = 16                                             # ~80 ms per trajectory: ~2.6 s per step
        pick = lambda k: {key: ics_all[key][:: max(1, len(ics_all[key]) // k)][:k] for key in TUBES}
        sample = pick(per_tube)
        for _ in range(max(args.warmup, 1)):
            arm.step(pick(1), mu)                                 # Numba compiles on the first call
        rk_steps = oracle_step_count(sample, mu)
        hits_total, t_total = 0, 0.0
        for _ in range(args.steps):
            h, dt = arm.step(sample, mu)
            hits_total += h
            t_total += dt
        single = {"value": rk_steps * args.steps / t_total, "unit": "RK steps/s", "cores": 1, "kind": "reference",
                  "crossings_per_s": hits_total / t_total,
                  "sample": f"{2 * per_tube} trajectories per step in ONE process (the reference as shipped: no parallel "
                            "driver for this loop)"}
        n_procs = max(1, min(os.cpu_count() or 1, 64))
        val, hits_per_s, n_traj, t_all = reference_all_cores(arm, ics_all, mu, per_tube, args.steps, n_procs)
        hits_total, t_total = hits_per_s * t_all, t_all
        kind, cores = "reference", n_procs
        what = (f"{n_traj} trajectories per step ({per_tube} per tube and worker, strided through the batch), sharded by "
                f"index over {n_procs} forked worker processes, each running the reference's own code from "
                f"oracle/_ref ({os.path.basename(arm.src.rstrip('/'))}): _propagate_dynsys per initial condition + "
                "_SynodicDetectionBackend.run")
        port_v, port_n, port_dt = cpu_port_throughput(ics_all, mu, n_threads, seconds_target=5.0)
        port = {"value": port_v, "unit": "RK steps/s", "cores": n_threads, "kind": "port",
                "sample": f"{port_n} trajectories, {port_dt:.1f} s"}
    else:
        sample = {key: ics_all[key][:4096] for key in TUBES}
        for _ in range(args.warmup):
            cpu_port_step({key: sample[key][:256] for key in TUBES}, mu, n_threads)
        steps_total, hits_total, t_total = 0.0, 0.0, 0.0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            st, hi = cpu_port_step(sample, mu, n_threads)
            t_total += time.perf_counter() - t0
            steps_total += st
            hits_total += hi
        val = steps_total / t_total
        kind, cores, port = "port", n_threads, None
        what = "8192 trajectories per step (4096 per tube), CPU oracle port (oracle/_ref absent)"
    line = {
        "impl": "reference", "metric": "fp64 CR3BP RK steps/s", "value": val, "unit": "RK steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "crossings_per_s": hits_total / t_total,
        "config": {"workload": WORKLOAD + "; bounded sample: " + what},
        "cpu_baseline": {"value": val, "unit": "RK steps/s", "cores": cores, "kind": kind, "sample": what},
        "e2e": {"value": val, "unit": "RK steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if port is not None:
        line["cpu_baseline_port"] = port
    if kind == "reference":
        line["cpu_baseline_single_core"] = single
    print(json.dumps(line))


def time_steps(fn, steps, flush, barrier, torch):
    """K timed iterations with an L2 flush between them; returns summed CUDA-event seconds."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for i in range(steps):
        flush.fill_(float(i))                                            # L2 flush between timed iterations
        ev[i][0].record()
        fn()
        ev[i][1].record()
    barrier()
    return sum(a.elapsed_time(b) for a, b in ev) * 1e-3


def secondary_configs(hb, torch, steps, flush, barrier):
    """BASELINE configs[2] (centre-manifold map, 1e5 seeds, Tao symplectic order 4) and configs[3] (42-state STM of
    the 100-member halo family, x128 replicas) -- reported beside the headline, parity variant."""
    from hiten_b200 import centermanifold as cm
    out = {}
    g = np.load(os.path.join(REPO, "tests", "golden", "cm_map.npz"))
    tab = cm.PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
    rng = np.random.default_rng(1)
    opts = cm.make_opts(0.01, 2000, "symplectic", 4, "p3", 20.0, "parity")
    hold = {}
    # configs[2]'s seeds: uniform in +-0.9 x the turning-point box on (q2, p2), lifted onto the energy
    # surface with hb_cm_lift (the reference's lift_plane_point), the first n valid ones -- all distinct
    Hs = cm.PolyTable.single(g["H_deg"], g["H_coef"], g["H_exp"])
    box = 0.9 * g["turning_q2_p2"]

    def lifted_seeds(n_seeds):
        got, have = [], 0
        while have < n_seeds:
            pts = torch.from_numpy(rng.uniform(-1.0, 1.0, (2 * n_seeds, 2)) * box).cuda()
            ok, st = cm.lift_plane_points(Hs, "p3", pts, float(g["energy"]))
            got.append(st[ok])
            have += int(ok.sum().item())
        return torch.cat(got)[:n_seeds].contiguous()

    for n_seeds, label in ((100_000, "cm_map_tao4_1e5_seeds"), (1_000_000, "cm_map_tao4_1e6_seeds")):
        seeds = lifted_seeds(n_seeds)

        def run_cm():
            hold["r"] = cm.poincare_map(tab, seeds, opts)

        for _ in range(3):
            run_cm()                                 # first call compiles the specialised kernel (NVRTC)
        t = time_steps(run_cm, steps, flush, barrier, torch)
        f, _, tt = hold["r"]
        cm_steps = float((tt / 0.01).ceil().sum().item())
        out[label] = {"distinct_lifted_seeds": int(torch.unique(seeds, dim=0).shape[0]),
                      "steps_per_s": cm_steps * steps / t, "crossings_per_s": int(f.sum().item()) * steps / t,
                      "ms_per_return": 1e3 * t / steps, "tflops": cm_steps * steps * 8100.0 / t / 1e12}
    out["cm_map_tao4_1e5_seeds"]["note"] = ("critical-path bound: the slowest seed needs ~1300 sequential steps; the late "
                                            "rounds (short work lists) run with four warps per 32 seeds (cm_map_split: "
                                            "~5.6 us per step instead of ~9.2); 1e6 seeds fill the machine")
    # the Tao integrator CLASS over a time grid (_ExtendedSymplectic.integrate): 1e5 trajectories x 100 grid intervals
    from hiten_b200 import symplectic as symp
    y6 = torch.zeros((100_000, 6), dtype=torch.float64, device="cuda")
    sd = torch.from_numpy(g["seeds_p3"][rng.integers(0, 512, 100_000)]).cuda()
    y6[:, 1], y6[:, 4], y6[:, 2], y6[:, 5] = sd[:, 0], sd[:, 1], sd[:, 2], sd[:, 3]
    tg = np.linspace(0.0, 1.0, 101)

    def run_symp():
        hold["s"] = symp.integrate_symplectic(tab, y6, tg, 4)

    for _ in range(3):
        run_symp()
    t = time_steps(run_symp, steps, flush, barrier, torch)
    out["tao4_grid_1e5_trajectories_x_100_intervals"] = {
        "steps_per_s": 1e7 * steps / t, "ms": 1e3 * t / steps, "tflops": 1e7 * steps * 12 * 610.0 / t / 1e12,
        "note": "_ExtendedSymplectic.integrate through hb_ham_symplectic_jit (run-time specialised gradient); 12 gradient "
                "evaluations of ~610 flop per Tao-4 step (SURVEY 8d), 48 B written per step"}
    # BASELINE configs[0] / [1] as they are (50 / 200 trajectories): small-batch latency, not throughput
    from hiten_b200 import synodic as syn
    c1 = np.load(os.path.join(REPO, "tests", "golden", "c1_manifold.npz"))
    te1 = np.linspace(0.0, float(c1["tf"]), 4713)
    x1 = torch.from_numpy(np.ascontiguousarray(c1["x0W"].T)).cuda()
    c2 = np.load(os.path.join(REPO, "tests", "golden", "synodic_c2.npz"))
    te2 = np.linspace(0.0, float(c2["tf"]), int(c2["steps"]))
    x2 = torch.from_numpy(np.ascontiguousarray(c2["x0W"].T)).cuda()
    sec2 = syn.make_section("y", 0.0, ("x", "z"), -1)
    r2 = syn.TubeSectionRunner(x2.shape[1], float(c2["mu"]), te2, sec2, forward=int(c2["forward"]), flip=(0, 6))

    def run_c1():
        hold["c1"] = hb.cr3bp_dense(x1, float(c1["mu"]), te1, forward=-1, flip=(0, 6), keep_on_device=True)

    def run_c2():
        r2.launch(x2)

    for fn in (run_c1, run_c2):
        for _ in range(3):
            fn()
    out["config1_manifold_50_trajectories_dense_4713"] = {"ms": 1e3 * time_steps(run_c1, steps, flush, barrier, torch) / steps,
                                                          "reference": "3.95 ms per trajectory (BASELINE.md) = ~200 ms"}
    out["config2_tube_200_trajectories_plus_section"] = {"ms": 1e3 * time_steps(run_c2, steps, flush, barrier, torch) / steps,
                                                         "crossings": r2.hit_count(),
                                                         "reference": "~0.8 s propagation + 21 s detection (SURVEY 8a17)"}
    # seed lifting for the CM map (1e6 plane points -> states on the energy surface)
    gl = np.load(os.path.join(REPO, "tests", "golden", "cm_lift.npz"))
    Ht = cm.PolyTable.single(gl["H_deg"], gl["H_coef"], gl["H_exp"])
    pts = torch.from_numpy(np.column_stack((rng.uniform(-1.1, 1.1, 1_000_000) * gl["turning"][0],
                                            rng.uniform(-1.1, 1.1, 1_000_000) * gl["turning"][1]))).cuda()

    def run_lift():
        hold["l"] = cm.lift_plane_points(Ht, "p3", pts, float(gl["energy"]))

    for _ in range(3):
        run_lift()
    t = time_steps(run_lift, steps, flush, barrier, torch)
    out["cm_lift_1e6_plane_points"] = {"lifts_per_s": 1e6 * steps / t, "ms_per_batch": 1e3 * t / steps,
                                       "liftable_fraction": float(hold["l"][0].float().mean().item()),
                                       "reference": "~1e3 lifts/s (one Python Brent solve per point)"}
    # connection search between two sets of 2e6 section hits (what a config-5 sweep produces)
    from hiten_b200 import connections as cn
    nn, eps = 2_000_000, 1.5e-4
    pu = rng.uniform(-0.4, 0.4, (nn, 2))
    ps = np.vstack((pu[rng.choice(nn, nn // 2, replace=False)] + rng.uniform(-1, 1, (nn // 2, 2)) * eps * 0.8,
                    rng.uniform(-0.4, 0.4, (nn - nn // 2, 2))))
    dd = [torch.from_numpy(a).cuda() for a in (pu, ps, rng.normal(0, 0.2, (nn, 6)), rng.normal(0, 0.2, (nn, 6)))]
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = cn.find_connections(*dd, eps, 0.5, 1e-3)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    out["connections_2e6_x_2e6_hits"] = {"ms": 1e3 * best, "pairs_considered": rc.pairs_considered,
                                         "accepted": int(len(rc.delta_v)),
                                         "note": "wall time incl. result D2H and ordering; the reference's pairing is an "
                                                 "O(N*M) = 4e12 double loop plus Python dicts"}
    del dd
    # tube initial conditions from a dense STM (2000 nodes x 500 displacements) and the two trajectory
    # filters on stored tubes (16384 x 4713 samples = 3.7 GB read once: HBM-bound)
    from hiten_b200 import manifold as mf
    gm = np.load(os.path.join(REPO, "tests", "golden", "manifold_ics.npz"))
    phi = np.zeros((gm["sp_tt"].size, 42))
    phi[gm["sp_rows"]] = gm["sp_phi_rows"]
    dphi, dtt = torch.from_numpy(phi).cuda(), torch.from_numpy(gm["sp_tt"]).cuda()
    dfr = torch.from_numpy(gm["fractions"][np.arange(2000) % gm["fractions"].size]).cuda()
    ddisp = torch.from_numpy(np.logspace(-7, -5, 500)).cuda()

    def run_ics():
        hold["ics"] = mf.tube_initial_conditions(dphi, dtt, float(gm["period"]), gm["sp_eigvec"], 1, dfr, ddisp)

    for _ in range(3):
        run_ics()
    t = time_steps(run_ics, steps, flush, barrier, torch)
    out["manifold_ics_1e6"] = {"ms": 1e3 * t / steps, "ics_per_s": 1e6 * steps / t}
    tube = torch.randn((16384, 4713, 6), dtype=torch.float64, device="cuda")

    def run_filter():
        hold["flt"] = mf.tube_filter(tube, float(gm["mu"]), safe_r1=3.3e-5, safe_r2=9e-6, energy_tol=1e-6)

    for _ in range(3):
        run_filter()
    t = time_steps(run_filter, steps, flush, barrier, torch)
    gbs = tube.numel() * 8 * steps / t / 1e9
    out["tube_filter_16384x4713"] = {"ms": 1e3 * t / steps, "hbm_read_gbs": gbs, "samples_per_s": 16384 * 4713 * steps / t,
                                     "note": "algorithmic bytes = 48 B per sample, read once"}
    del tube
    hold.pop("flt", None)
    # batched differential correction, 1e5 perturbed halo guesses in one lock-step batch
    from hiten_b200 import corrector as cr
    gc = np.load(os.path.join(REPO, "tests", "golden", "correction.npz"))
    for fam, nb in (("halo", 100_000), ("lyapunov", 100_000)):
        base = gc[f"{fam}_x0"][gc[f"{fam}_iters"] >= 0]
        xg = base[rng.integers(0, len(base), nb)].copy()
        xg[:, cr.FAMILIES[fam][0]] += 1e-4 * rng.standard_normal((nb, 2))
        xgd = torch.from_numpy(np.ascontiguousarray(xg.T)).cuda()
        copts = cr.make_opts(fam)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rcorr = cr.correct_orbits(xgd, float(gc["mu"]), copts)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out[f"correct_{fam}_1e5_orbits"] = {
            "ms": 1e3 * best, "orbits_per_s": nb / best, "converged_fraction": float((rcorr.status == 0).float().mean().item()),
            "mean_newton_iterations": float(rcorr.iterations.float().mean().item()),
            "rk_steps_per_s": (rcorr.rk_steps6 + rcorr.rk_steps42) / best,
            "note": "wall time of the whole Newton + Armijo loop (every event / STM propagation, solve and line-search "
                    "decision on the GPU; the host reads 4 bytes between launches)"}
    del xgd
    s = np.load(os.path.join(REPO, "tests", "golden", "stm_family.npz"))
    x0 = torch.from_numpy(np.ascontiguousarray(np.tile(s["x0"], (128, 1)).T)).cuda()
    T = torch.from_numpy(np.tile(s["period"], 128)).cuda()
    integ = hb.make_integ(arith="parity")

    def run_stm():
        hold["s"] = hb.cr3bp_stm(x0, float(s["mu"]), 0.0, tf_per_traj=T, integ=integ)

    for _ in range(3):
        run_stm()
    t = time_steps(run_stm, steps, flush, barrier, torch)
    st = int((hold["s"].n_acc.sum() + hold["s"].n_rej.sum()).item())
    out["stm42_family_x128"] = {"rk_steps_per_s": st * steps / t, "trajectories": int(x0.shape[1]),
                                "tflops": st * steps * 9500.0 / t / 1e12}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--concurrent-tubes", type=int, default=1,
                    help="1: the two tubes' pipelines on two streams with a step scratch each; 0: one after the other")
    ap.add_argument("--arith", default="parity", choices=["parity", "fast"])
    ap.add_argument("--max-ctas", default="auto",
                    help="auto: when the two tubes run on two streams and a tube has fewer than 8 trajectories per lane of a "
                         "full-device persistent launch, each tube's launch is capped at half the SMs so that the two run "
                         "side by side (the strong-scaling shards); 0: every launch uses every SM; N: cap at N CTAs")
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--pipeline", default="section2", choices=["section2", "section3", "fused"],
                    help="section2: step records through an HBM scratch (default, fastest); section3: records handed over "
                         "in shared memory inside one kernel (no step scratch); fused: hb_cr3bp_section")
    ap.add_argument("--steps-capacity", type=int, default=None)
    ap.add_argument("--records", default="near", choices=["near", "all"],
                    help="section2: record only the steps near the section plane (default) or every accepted step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary timings")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.steps_capacity is None:
        args.steps_capacity = STEPS_CAPACITY if args.records == "near" else 192

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import hiten_b200 as hb
    from hiten_b200 import propagate as P
    from hiten_b200 import sharded, synodic
    from hiten_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.n_per_gpu
    integ = hb.make_integ(arith=args.arith)
    kind = {"section2": dict(steps_capacity=args.steps_capacity, records=args.records), "section3": dict(pool_records=8), "fused": {}}[args.pipeline]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count

    def ctas_for(n_tube_local):
        if args.max_ctas != "auto":
            return int(args.max_ctas)
        if not args.concurrent_tubes or args.pipeline == "section3":
            return 0
        return sm_count // 2 if n_tube_local < 8 * sm_count * 256 else 0

    def make_job(n_total_global):
        """Runners + resident inputs of this rank for a job of n_total_global trajectories (both tubes, all ranks)."""
        ics, mu = W.c5_batch(n_total_global, rank, world)
        job = {"ics": ics, "mu": mu, "tubes": {}}
        job["max_ctas"] = ctas_for(n_total_global // 2 // world)
        scratch = None
        for key in TUBES:
            x = ics[key]
            integ = hb.make_integ(arith=args.arith, max_ctas=job["max_ctas"])
            d = sharded.DistributedTubeSection(
                n_total_global // 2, mu, W.c5_grid(key), W.c5_section(key, mu), forward=W.C5_TUBES[key]["forward"],
                flip=(0, 6), integ=integ,
                runner_factory=lambda nl, key=key, integ=integ: synodic.TubeSectionRunner(
                    nl, mu, W.c5_grid(key), W.c5_section(key, mu), forward=W.C5_TUBES[key]["forward"], flip=(0, 6),
                    integ=integ, device=dev, scratch=None if args.concurrent_tubes else scratch, **kind))
            assert len(d.index) == len(x)
            if scratch is None:
                scratch = d.runner.scratch
            job["tubes"][key] = {"dist": d, "run": d.runner,
                                 "y0": torch.from_numpy(np.ascontiguousarray(x.T)).to(dev),      # resident input [6, n]
                                 "host": torch.from_numpy(x).pin_memory()}                      # e2e input [n, 6]
        job["scratch"] = scratch
        return job

    job = make_job(n * world)
    mu = job["mu"]
    main_max_ctas = job["max_ctas"]
    peer_exchange = world > 1 and all(job["tubes"][key]["dist"].px is not None for key in TUBES)

    side = [torch.cuda.Stream(dev) for _ in TUBES] if args.concurrent_tubes else None

    trace = [] if os.environ.get("HITEN_B200_BENCH_TRACE") else None     # host time stamps inside a step (debug aid)

    def step_of(j):
        def step():
            # The two tubes' pipelines on two streams (each with its own step scratch): the propagation kernels are
            # persistent with one CTA per SM, so tube 2's CTAs take over an SM the moment tube 1's CTA there runs out of
            # trajectories, and tube 1's scan / emit kernels fill the SMs tube 2's propagation vacates -- the idle tail of
            # each persistent launch is covered by the other tube's work.
            main = torch.cuda.current_stream()
            ts = [time.perf_counter()] if trace is not None else None
            for i, key in enumerate(TUBES):
                t = j["tubes"][key]
                if side is not None:
                    side[i].wait_stream(main)
                    t["run"].launch(t["y0"], side[i])
                else:
                    t["run"].launch(t["y0"])
            if ts is not None:
                ts.append(time.perf_counter())
            if world > 1:                                   # the one exchange: hit records, counts, end states -> rank 0
                peer = []
                for i, key in enumerate(TUBES):             # tube 1's copies run under tube 2's propagation
                    peer.append(j["tubes"][key]["dist"].start_gather(None if side is None else side[i]))
                    if ts is not None:
                        ts.append(time.perf_counter())
                for i, (key, started) in enumerate(zip(TUBES, peer)):
                    d = j["tubes"][key]["dist"]
                    if started:
                        d.finish_gather()
                    else:
                        with torch.cuda.stream(main if side is None else side[i]):
                            d.gather_device()
                    if ts is not None:
                        ts.append(time.perf_counter())
            if side is not None:
                for st_ in side:
                    main.wait_stream(st_)
            if ts is not None:
                trace.append([1e3 * (b - ts[0]) for b in ts[1:]])
        return step

    step_resident = step_of(job)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    wall0 = time.perf_counter()
    t_dev = time_steps(step_resident, args.steps, flush, barrier, torch)
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()

    def tallies(j):
        acc = rej = hits = over = 0
        ok = True
        for key in TUBES:
            run = j["tubes"][key]["run"]
            if run.scratch is not None:
                nt = ctypes.c_int64(0)
                run.lib.hb_read_record_overflow(run.ws.data_ptr(), ctypes.byref(nt), None)
                over += int(nt.value)
            hits += run.hit_count()
            acc += int(run.nacc.sum().item())
            rej += int(run.nrej.sum().item())
            ok = ok and bool((run.status == 0).all().item())
        return acc, rej, hits, over, ok

    steps_acc, steps_rej, n_hits, n_overflow, ok = tallies(job)
    steps_per_pass = steps_acc + steps_rej

    def pilot_orders(j):
        """Launch orders that need NO earlier pass over the batch: the attempted steps of a PILOT -- every 40th displacement
        row of the tube (2000 orbit nodes each, ~3 % of the trajectories), one propagate-only launch -- interpolated linearly
        in the row index per node (the cost is smooth in (node, log displacement): R2 0.96-0.98, tools/sim_launch_order.py).
        Returns the wall time of pilots + models + sorts for both tubes in ms."""
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for key in TUBES:
            t_ = j["tubes"][key]
            run_, y0_ = t_["run"], t_["y0"]
            n_rows = run_.n // 2000
            rows = np.unique(np.concatenate((np.arange(0, n_rows, 40), [n_rows - 1])))
            idx = torch.from_numpy((rows[:, None] * 2000 + np.arange(2000)[None, :]).ravel()).to(dev)
            r_ = hb.cr3bp_propagate(y0_[:, idx].contiguous(), mu, float(W.c5_grid(key)[-1]), forward=W.C5_TUBES[key]["forward"],
                                    flip=(0, 6), integ=hb.make_integ(arith=args.arith))
            cp = (r_.n_acc + r_.n_rej).to(torch.float64).view(len(rows), 2000)
            full = torch.arange(n_rows, device=dev, dtype=torch.float64)
            rt = torch.from_numpy(rows.astype(np.float64)).to(dev)
            hi = torch.searchsorted(rt, full).clamp(1, len(rows) - 1)
            lo = hi - 1
            wgt = ((full - rt[lo]) / (rt[hi] - rt[lo])).clamp(0.0, 1.0)[:, None]
            pred = ((1.0 - wgt) * cp[lo] + wgt * cp[hi]).reshape(-1)
            run_.order_by_cost(pred)
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - w0)

    def ordered_leg(j, stepfn, ref_steps, ref_hits, pilot=False):
        """Secondary line: the same step with each tube's persistent propagation launch handing out its trajectories
        LONGEST FIRST (hb_integ.order -- a scheduling hint, outputs stay in the caller's indexing; SURVEY 8e "sorting by
        expected cost").  The cost is the step count of the previous pass over the same batch, i.e. the best case of
        what a caller with similar successive batches gets from TubeSectionRunner.order_by_cost(); it is NOT the headline,
        which hands the trajectories out in the order the workload generator produced them."""
        if args.pipeline != "section2":
            return None
        pilot_ms = None
        try:
            if pilot:
                pilot_ms = pilot_orders(j)
            else:
                for key in TUBES:
                    j["tubes"][key]["run"].order_by_cost()
            for _ in range(2):
                stepfn()
            barrier()
            t_o = time_steps(stepfn, args.steps, flush, barrier, torch)
            a_, r_, h_, _, ok_ = tallies(j)
        finally:
            for key in TUBES:
                j["tubes"][key]["run"].set_order(None)
        v = torch.tensor([t_o, float(a_ + r_), float(h_)], dtype=torch.float64, device=dev)
        if world > 1:
            vmax = v.clone()
            dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            v[0] = vmax[0]
        t_o, st_o, hi_o = v[0].item(), v[1].item(), v[2].item()
        return {"ms_per_step": 1e3 * t_o / args.steps, "rk_steps_per_s": st_o * args.steps / t_o,
                "crossings_per_s": hi_o * args.steps / t_o,
                "same_steps_and_crossings_as_natural_order": bool(a_ + r_ == ref_steps and h_ == ref_hits and ok_),
                "order": ("per tube, argsort(-predicted cost): a pilot launch of every 40th displacement row (16 000 of 500 000 "
                          "trajectories per tube), interpolated per node; no earlier pass over the batch needed; the pilot is "
                          "run once, OUTSIDE the timed steps (pilot_and_model_ms: its wall time for both tubes)") if pilot else
                         "per tube, argsort(-(n_acc + n_rej)) of the previous pass over the same batch (hb_integ.order)",
                "pilot_and_model_ms": pilot_ms}

    cost_ordered = ordered_leg(job, step_resident, steps_per_pass, n_hits)
    cost_ordered_pilot = None
    if world == 1 and cost_ordered is not None and n % 4000 == 0 and n >= 160000:
        try:
            cost_ordered_pilot = ordered_leg(job, step_resident, steps_per_pass, n_hits, pilot=True)
        except (ValueError, RuntimeError, IndexError) as exc:          # a secondary line must not take the headline down
            cost_ordered_pilot = {"error": repr(exc)}
    records_written = sum(int(job["tubes"][key]["run"].records_written().sum().item()) for key in TUBES) \
        if args.pipeline == "section2" else 0

    # per-kernel split of the pipeline: caller-owned CUDA events recorded by the library between its kernels on the
    # launching stream, in a separate pass over the same inputs
    stage_ms = None
    if args.pipeline != "fused":
        n_ev = 5 if args.pipeline == "section2" else 4
        acc_ms = np.zeros(n_ev - 1)
        for key in TUBES:
            t = job["tubes"][key]
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
            t["run"].set_stage_events(ev)
            for i in range(args.steps):
                flush.fill_(float(i))
                t["run"].launch(t["y0"])
                torch.cuda.synchronize()
                acc_ms += np.array([ev[a].elapsed_time(ev[a + 1]) for a in range(n_ev - 1)])
            t["run"].set_stage_events(None)
        stage_ms = (acc_ms / args.steps).tolist()

    # e2e: host buffers in, host results out, every step's H2D and D2H inside the timed region (overlapped with the
    # neighbouring steps' compute by TubeSectionStream's double buffering); hits are put in the reference's order
    streams = {}
    for key in TUBES:
        skw = dict(steps_capacity=args.steps_capacity, records=args.records) if args.pipeline == "section2" else \
            (dict(steps_capacity=0, pool_records=8) if args.pipeline == "section3" else dict(steps_capacity=0, pool_records=0))
        streams[key] = synodic.TubeSectionStream(len(job["ics"][key]), mu, W.c5_grid(key), W.c5_section(key, mu),
                                                 forward=W.C5_TUBES[key]["forward"], flip=(0, 6), integ=integ, device=dev,
                                                 scratch=job["scratch"], **skw)

    def e2e_pass(k):
        total, last_ordered = 0, None
        for key in TUBES:
            for r in streams[key].run([job["tubes"][key]["host"]] * k):
                last_ordered = r.hits                       # already in the reference's order (sorted on the device)
                last = r.n_hits
            total += last
        return total, last_ordered

    e2e_pass(2)
    barrier()
    e2e_t0 = time.perf_counter()
    k_e2e, _ = e2e_pass(args.steps)
    barrier()
    e2e_t = time.perf_counter() - e2e_t0
    e2e_ok = bool(k_e2e == n_hits)
    del streams

    # strong scaling (configs[4] = 1e6 trajectories in total over the N GPUs): same step on 1/N of the batch per rank
    strong = None
    if world > 1:
        del job["tubes"]
        sjob = make_job(n)
        sstep = step_of(sjob)
        for _ in range(3):
            sstep()
        barrier()
        if trace is not None:
            del trace[:]
        ts = time_steps(sstep, args.steps, flush, barrier, torch)
        if trace is not None:
            print(f"[trace rank {rank}] strong leg, ms since step start (launched, start_gather x2, finish x2): "
                  f"{np.mean(np.array(trace), axis=0).round(3).tolist()}", file=sys.stderr)
        s_acc, s_rej, s_hits, _, s_ok = tallies(sjob)
        strong = [ts, float(s_acc + s_rej), float(s_hits)]
        strong_ordered = ordered_leg(sjob, sstep, s_acc + s_rej, s_hits)
        strong_max_ctas = sjob["max_ctas"]
        d0 = sjob["tubes"][TUBES[0]]["dist"]
        strong_exchange = ("nccl gather" if d0.px is None else
                           "hb_peer_put: each rank's kernel reads its hit count on the device and writes records + end states "
                           "into rank 0's buffer over NVLink peer memory, no host wait before the closing barrier"
                           if d0._device_put is True else "peer memory, copy engines (as the weak-scaling step)")
    else:
        strong = [t_dev, float(steps_per_pass), float(n_hits)]
        strong_ordered = cost_ordered
        strong_max_ctas = main_max_ctas
        strong_exchange = None

    tt = torch.tensor([t_dev, e2e_t, float(steps_per_pass), float(n_hits), float(steps_acc), strong[0], strong[1], strong[2],
                       float(n_overflow)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_dev, e2e_t, strong_t = tmax[0].item(), tmax[1].item(), tmax[5].item()
        total_steps_pass, total_hits, total_acc = tsum[2].item(), tsum[3].item(), tsum[4].item()
        strong_steps, strong_hits, total_overflow = tsum[6].item(), tsum[7].item(), tsum[8].item()
    else:
        total_steps_pass, total_hits, total_acc = float(steps_per_pass), float(n_hits), float(steps_acc)
        strong_t, strong_steps, strong_hits, total_overflow = strong[0], strong[1], strong[2], float(n_overflow)

    extra = {"strong_scaling": {"total_trajectories": n, "n_gpus": world, "ms_per_step": 1e3 * strong_t / args.steps,
                                "rk_steps_per_s": strong_steps * args.steps / strong_t,
                                "crossings_per_s": strong_hits * args.steps / strong_t,
                                "max_ctas_per_launch": strong_max_ctas or None, "exchange": strong_exchange,
                                "note": "configs[4]'s 1e6 trajectories in TOTAL, 1/N per GPU, same step incl. the gather; "
                                        "efficiency = rk_steps_per_s / (N x the N=1 value)",
                                "cost_ordered": strong_ordered},
             "cost_ordered": cost_ordered, "cost_ordered_pilot_model": cost_ordered_pilot}
    if not args.no_extra and world == 1:
        if args.concurrent_tubes and args.pipeline != "fused":       # the secondary lines run one tube at a time: one scratch
            job["tubes"][TUBES[1]]["run"].scratch = job["scratch"]
            torch.cuda.empty_cache()
        ws = P.workspace(dev)
        y0_l1, tf_l1 = job["tubes"]["l1"]["y0"], float(W.c5_grid("l1")[-1])
        for name in ("parity", "fast"):
            ig = hb.make_integ(arith=name)
            hold = {}

            def prop():
                hold["r"] = hb.cr3bp_propagate(y0_l1, mu, tf_l1, forward=-1, flip=(0, 6), integ=ig, ws=ws)

            for _ in range(3):
                prop()
            tp = time_steps(prop, args.steps, flush, barrier, torch)
            sp = int((hold["r"].n_acc.sum() + hold["r"].n_rej.sum()).item())
            extra[f"propagate_only_{name}_l1_tube"] = {"rk_steps_per_s": sp * args.steps / tp, "ms": 1e3 * tp / args.steps,
                                                       "tflops": sp * args.steps * FLOP_PER_STEP / tp / 1e12}
        # the same step in the other arithmetic variant and through the other two forms of the pipeline
        other = "fast" if args.arith == "parity" else "parity"
        variants = [(f"step_{other}", other, kind), ("step_section3_parity", "parity", dict(pool_records=8)),
                    ("step_fused_kernel_parity", "parity", {})]
        if args.pipeline == "section2" and args.records == "near":
            variants.insert(1, ("step_section2_all_records_parity", "parity", dict(steps_capacity=192, records="all")))
        for label, ar, kw in variants:
            if label == "step_section3_parity" and args.pipeline == "section3":
                continue
            rs = {key: synodic.TubeSectionRunner(len(job["ics"][key]), mu, W.c5_grid(key), W.c5_section(key, mu),
                                                 forward=W.C5_TUBES[key]["forward"], flip=(0, 6),
                                                 integ=hb.make_integ(arith=ar), device=dev,
                                                 scratch=job["scratch"] if kw is kind and "steps_capacity" in kw else None, **kw)
                  for key in (TUBES[:1] if kw is not kind and "steps_capacity" in kw else TUBES)}
            if len(rs) == 1:                                 # all-records scratch (49 GB): the second tube shares it
                rs[TUBES[1]] = synodic.TubeSectionRunner(len(job["ics"][TUBES[1]]), mu, W.c5_grid(TUBES[1]),
                                                         W.c5_section(TUBES[1], mu), forward=W.C5_TUBES[TUBES[1]]["forward"],
                                                         flip=(0, 6), integ=hb.make_integ(arith=ar), device=dev,
                                                         scratch=rs[TUBES[0]].scratch, **kw)

            def step2():
                for key in TUBES:
                    rs[key].launch(job["tubes"][key]["y0"])

            for _ in range(2):
                step2()
            t2 = time_steps(step2, max(args.steps // 2, 2), flush, barrier, torch)
            s2 = sum(int((rs[key].nacc.sum() + rs[key].nrej.sum()).item()) for key in TUBES)
            k2 = sum(rs[key].hit_count() for key in TUBES)
            extra[label] = {"rk_steps_per_s": s2 * max(args.steps // 2, 2) / t2, "ms_per_step": 1e3 * t2 / max(args.steps // 2, 2),
                            "crossings_per_pass": k2, "tflops_whole_step": s2 * max(args.steps // 2, 2) * FLOP_PER_STEP / t2 / 1e12}
            del rs
            torch.cuda.empty_cache()
        if "step_fast" in extra:
            extra["step_fast"]["parity_verdict"] = (
                "vs the reference's golden vectors (tests/test_gpu_fast_verdict.py): crossing counts identical on all five "
                "golden tubes (1209 crossings), crossing points <= 1e-9 on 1208 of 1209 (all 386 of C5; max 2.4e-9 on one C1 "
                "hit), end states > 1e-9 relative on 23 of 670 trajectories (max 7.5e-8, inside the reference's own 1-ulp "
                "sensitivity): NOT the parity headline")
        if args.pipeline == "section2":
            # the step with Manifold.compute()'s trajectory filters judged from the step records
            r1, r2 = W.c5_safe_radii()
            rf, fscratch = {}, None
            for key in TUBES:                                # every accepted step is recorded here: own (larger) scratch
                rf[key] = synodic.TubeSectionRunner(len(job["ics"][key]), mu, W.c5_grid(key), W.c5_section(key, mu),
                                                    forward=W.C5_TUBES[key]["forward"], flip=(0, 6), integ=integ, device=dev,
                                                    steps_capacity=192, scratch=fscratch, filters=(r1, r2, W.ENERGY_TOL))
                fscratch = rf[key].scratch

            def step3():
                for key in TUBES:
                    rf[key].launch(job["tubes"][key]["y0"])

            step3()
            t3 = time_steps(step3, 2, flush, barrier, torch)
            kept = {key: int((rf[key].filter_result()[1] == 1).sum().item()) for key in TUBES}
            extra["step_with_trajectory_filters"] = {
                "ms_per_step": 1e3 * t3 / 2, "kept_trajectories": kept,
                "note": "hb_section2_filter: 6-component dense evaluation + r1, r2, Jacobi constant at every grid sample "
                        "(safe radii 3.318e-05 / 9.04e-06, energy_tol 1e-6 as Manifold.compute())"}
            del rf, fscratch
            torch.cuda.empty_cache()
        # the connection search between the two hit sets of the last step
        from hiten_b200 import connections as cn
        hu, hs = job["tubes"]["l1"]["run"].sorted_hits(), job["tubes"]["l2"]["run"].sorted_hits()
        t0 = time.perf_counter()
        rc = cn.find_connections(hu.points, hs.points, hu.states, hs.states, 1e-3, 1.0, 1e-8,
                                 traj_indices_u=hu.trajectory_indices, traj_indices_s=hs.trajectory_indices)
        extra["connections_between_the_two_sections"] = {
            "ms": 1e3 * (time.perf_counter() - t0), "hits_u": len(hu.times), "hits_s": len(hs.times),
            "pairs_considered": rc.pairs_considered, "accepted": int(len(rc.delta_v)),
            "note": "ConnectionPipeline.solve's backend step (eps2d=1e-3, delta_v_tol=1, ballistic_tol=1e-8) on this step's "
                    "hits, wall time incl. H2D of the hit sets and D2H of the result"}
        extra.update(secondary_configs(hb, torch, max(args.steps // 2, 2), flush, barrier))

    if rank == 0:
        value = total_steps_pass * args.steps / t_dev
        e2e_value = total_steps_pass * args.steps / e2e_t
        peak = hb.dfma_peak(200.0)
        hbm = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] \
            if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6500.0
        rec_bytes = records_written * 512.0
        if stage_ms is not None:
            # Roofline of the DOMINANT kernel (the propagation kernel of both tubes: k_dop853_6 in record mode + its
            # first-step pre-pass, or the producer/consumer kernel of section3): attempted steps x 1350 flop over that
            # kernel's own duration.
            achieved = steps_per_pass * FLOP_PER_STEP / (stage_ms[0] * 1e-3)
            names = ("propagate_record", "step_scan", "emit_candidates", "order_dedup") if args.pipeline == "section2" \
                else ("propagate_and_scan", "emit_candidates", "order_dedup")
            extra["pipeline"] = {"stage_ms": dict(zip(names, stage_ms)), "share_of_step_dominant": stage_ms[0] / sum(stage_ms),
                                 "hbm_peak_gbs": hbm}
            if args.pipeline == "section2":
                extra["pipeline"].update({
                    "propagate_record_hbm_write_gbs": rec_bytes / (stage_ms[0] * 1e-3) / 1e9,
                    "step_scan_hbm_read_gbs": rec_bytes / (stage_ms[1] * 1e-3) / 1e9,
                    "step_scan_frac_of_hbm_peak": rec_bytes / (stage_ms[1] * 1e-3) / 1e9 / hbm})
            traffic = (rec_bytes if args.pipeline == "section2" else 0.0) + len(job["ics"]["l1"]) * 2 * (48 + 48 + 12)
        else:
            achieved = (steps_per_pass * FLOP_PER_STEP + steps_acc * FLOP_PER_SEGMENT) * args.steps / t_dev
            traffic = None
        for v in extra.values():
            if isinstance(v, dict) and "tflops" in v:
                v["frac_of_fp64_peak"] = v["tflops"] * 1e12 / peak
            if isinstance(v, dict) and "tflops_whole_step" in v:
                v["frac_of_fp64_peak_whole_step"] = v["tflops_whole_step"] * 1e12 / peak
        launches_per_tube = {"section2": 6 if args.arith == "parity" else 5, "section3": 5 if args.arith == "parity" else 4,
                             "fused": 2 if args.arith == "parity" else 1}[args.pipeline]
        line = {
            "metric": "fp64 CR3BP RK steps/s", "value": value, "unit": "RK steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "crossings_per_s": total_hits * args.steps / t_dev,
            "whole_step_tflops": total_steps_pass * args.steps * FLOP_PER_STEP / t_dev / 1e12,
            "whole_step_frac_of_fp64_peak": total_steps_pass * args.steps * FLOP_PER_STEP / t_dev / (peak * world),
            "config": {
                "workload": WORKLOAD,
                "trajectories_per_gpu": n, "trajectories_per_tube_per_gpu": n // 2,
                "grid_samples": {key: int(len(W.c5_grid(key))) for key in TUBES}, "arith": args.arith,
                "path": {"section2": "hb_cr3bp_section2 per tube (propagate + screen + record -> step scan -> emit -> order+dedup)",
                         "section3": "hb_cr3bp_section3 per tube (propagating + scanning warps in one kernel -> emit -> order+dedup)",
                         "fused": "hb_cr3bp_section per tube (fused kernel)"}[args.pipeline],
                "tube_scheduling": ("the two tubes' pipelines on two streams, a step scratch each: a persistent propagation "
                                    "launch's idle tail is covered by the other tube's kernels" if args.concurrent_tubes
                                    else "one tube after the other on one stream"),
                "max_ctas_per_launch": main_max_ctas or None,
                "steps_capacity": args.steps_capacity if args.pipeline == "section2" else None,
                "step_records": None if args.pipeline != "section2" else
                                {"mode": args.records, "written_per_pass_this_gpu": records_written,
                                 "share_of_accepted_steps": records_written / max(steps_acc, 1)},
                "record_overflow_trajectories_rerun": total_overflow,
                "l2": "flushed between timed iterations (256 MB fill); inputs 48 B and step records ~50 KB per trajectory (>> L2)",
                "rk_steps_per_pass": total_steps_pass, "accepted_steps_per_pass": total_acc,
                "crossings_per_pass": total_hits, "all_status_ok": ok,
                "exchange": None if world == 1 else (
                    "per tube, inside the timed step: every rank writes its hit records and end states into rank 0's "
                    "receive buffer over NVLink peer memory (symmetric memory, copy engines; tube 1's transfer runs under "
                    "tube 2's propagation), one signal-pad barrier at the end"
                    if peer_exchange else
                    "per tube: all-gather of hit counts, NCCL gather of hit records (padded to the largest shard) and end "
                    "states to rank 0, inside the timed step"),
            },
            "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": {"section2": "k_dop853_6<MODE_RECORD> (+ k_first_steps), both tubes",
                                    "section3": "k_tube_section_pc (+ k_first_steps), both tubes",
                                    "fused": "k_dop853_6_section, both tubes"}[args.pipeline],
                         "note": "FP64 FMA pipe roofline of the dominant kernel: achieved = attempted steps x 1350 algorithmic "
                                 "flop (SURVEY 8d) / the kernel's own duration (CUDA events recorded between the pipeline's "
                                 "kernels on the launching stream); peak = hb_dfma_peak measured in this process "
                                 "(MEASURED_PEAKS.json has no FP64 entry). traffic = bytes this run moved through HBM for that "
                                 "kernel, counted from the run: 512 B per step record written + 108 B per trajectory "
                                 "(ncu dram bytes of the same kernel: profiles/r02_*)"},
            "e2e": {"value": e2e_value, "unit": "RK steps/s", "h2d_bytes_per_step": int(n * 48),
                    "d2h_bytes_per_step": int(n * (48 + 16) + 72 * k_e2e), "same_hits_as_resident_run": e2e_ok,
                    "api": "synodic.TubeSectionStream.run per tube (pinned host batches in, pinned host results out, hits "
                           "put in the reference's order (trajectory, seq) inside the timed region)",
                    "crossings_per_s": total_hits * args.steps / e2e_t},
            "gpu_launches": world * args.steps * 2 * launches_per_tube,
            "clocks": clocks, "wall_s_timed_region": wall, "extra": extra,
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle_lib as O
            nthr = O.lib().ho_max_threads()
            v, ns, dt = cpu_port_throughput(job["ics"], mu, nthr)
            port = {"value": v, "unit": "RK steps/s", "cores": nthr, "kind": "port",
                    "sample": f"{ns} trajectories of the same batch (tube propagation + dense grid + section detection per "
                              f"trajectory), {dt:.1f} s"}
            line["cpu_baseline"] = port
            if reference_available():
                try:
                    arm = ReferenceArm()
                    pick = lambda k: {key: job["ics"][key][:: max(1, len(job["ics"][key]) // k)][:k] for key in TUBES}
                    arm.step(pick(1), mu)                       # Numba compiles here
                    sample = pick(120)                            # ~12 s of the reference's own code on one core
                    h, dt = arm.step(sample, mu)
                    line["cpu_baseline"] = {
                        "value": oracle_step_count(sample, mu) / dt, "unit": "RK steps/s", "cores": 1, "kind": "reference",
                        "crossings_per_s": h / dt,
                        "sample": f"{sum(len(v) for v in sample.values())} trajectories of the same batch (120 per tube, "
                                  f"strided): the reference's own "
                                  f"_propagate_dynsys per initial condition + _SynodicDetectionBackend.run from oracle/_ref, "
                                  f"{dt:.1f} s on 1 core (it has no parallel driver for this loop)"}
                    line["cpu_baseline_port"] = port
                except Exception as exc:                        # the reference failed to import / run: keep the port
                    line["cpu_baseline_note"] = f"reference arm failed ({type(exc).__name__}: {exc}); port reported"
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
