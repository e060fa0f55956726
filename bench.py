#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native HITEN propagation hot path.

Workload (BASELINE.json configs[4] geometry at per-GPU size; "C5-tube"): stable-manifold tube of the
Earth-Moon L1 halo (Az=0.2 S) of configs[0]: 2000 orbit nodes x D displacements log-spaced in
[1e-7, 1e-5], every trajectory propagated backward over tf = 0.75*2*pi with
DOP853 at rtol = atol = 1e-12 -- N = 131072 trajectories per GPU (8 GPUs ~ 1e6 = configs[4]),
weak scaling, no data-path collective, one gather of end states at the end of a step (N > 1).

A "step" = one pass of the hot path over the per-GPU batch: the fused tube + synodic-section kernel
(hb_cr3bp_section: DOP853 propagation, dense samples on the reference's dt = 1e-3 grid streamed through the
reference's section detector, hits appended to a buffer, end states written).  metric = fp64 CR3BP RK steps/s
(attempted DOP853 steps, accepted + rejected, whole job); crossings/s is reported beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--arith parity|fast] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

FLOP_PER_STEP = 1350.0          # algorithmic flop per attempted 6-state DOP853 step
FLOP_PER_SEGMENT = 1050.0       # dense-output cache of one accepted step (3 RHS + D/A_ext rows)
FLOP_PER_SAMPLE = 84.0          # one dense sample (7-term Horner x 6 components)
GRID_DT = 1.0e-3                # Manifold.compute default dt -> 4713 samples over tf
N_PER_GPU = 1_000_000            # BASELINE configs[4]: 1e6 manifold trajectories (fits one B200: 82 GB of step scratch)
TF = 0.75 * 2.0 * np.pi


def build_ics(n, rank=0, world=1):
    """Deterministic synthetic batch: 2000 tube nodes x displacements, interleaved over ranks."""
    from hiten_b200.manifold import manifold_initial_conditions
    t = np.load(os.path.join(REPO, "tests", "golden", "tube_nodes_c1.npz"))
    total = n * world
    n_disp = (total + 1999) // 2000
    disp = np.logspace(-7.0, -5.0, n_disp)
    ics = manifold_initial_conditions(t["x_node"], t["man"], disp)[:total]
    # displacement-major order: neighbouring lanes carry neighbouring orbit phases
    return np.ascontiguousarray(ics[rank::world][:n]), float(t["mu"])


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


def cpu_tube_section(ics, mu, n_threads):
    """The same step on the CPU oracle: dense tube on the dt=1e-3 grid + section detection. Returns RK steps, hits."""
    import oracle_lib as O
    s = O.system(O.SYS_CR3BP6, mu, fwd=-1, flip=(0, 6))
    m = max(int(abs(TF) / GRID_DT) + 1, 100)
    t_eval = np.linspace(0.0, TF, m)
    times = -t_eval
    steps, hits = 0, 0
    chunk = 512                                            # 512 x 4713 x 48 B = 116 MB of dense output at a time
    for a in range(0, len(ics), chunk):
        dense, c = O.batch_dense(s, O.DOP853, O.default_tol(), ics[a:a + chunk], t_eval, n_threads)
        steps += int(c.sum())
        hits += O.batch_synodic_count(times, dense, 1, 0.0, -1, (0, 2), 50, 1e-6, 1e-9, 1e-6, n_threads)
    return steps, hits


def cpu_port_throughput(ics, mu, n_threads, seconds_target=12.0):
    """The oracle (C port of the reference algorithm) on the host cores, bounded sample of the same step."""
    t0 = time.perf_counter()
    cpu_tube_section(ics[:256], mu, n_threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    n = int(min(len(ics), max(512, 256 / dt * seconds_target)))
    t0 = time.perf_counter()
    steps, _ = cpu_tube_section(ics[:n], mu, n_threads)
    dt = time.perf_counter() - t0
    return steps / dt, n, dt


def run_reference(args):
    """--impl reference: the reference algorithm's CPU restatement (oracle port, all host threads) on the same
    step -- tube propagation, dense samples on the dt=1e-3 grid, section detection -- on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O
    n_threads = O.lib().ho_max_threads()
    sample = 8192
    ics, mu = build_ics(sample)
    for _ in range(args.warmup):
        cpu_tube_section(ics[:512], mu, n_threads)
    steps_total, hits_total, t_total = 0.0, 0.0, 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        st, hi = cpu_tube_section(ics, mu, n_threads)
        t_total += time.perf_counter() - t0
        steps_total += st
        hits_total += hi
    val = steps_total / t_total
    line = {
        "impl": "reference", "metric": "fp64 CR3BP RK steps/s", "value": val, "unit": "RK steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "crossings_per_s": hits_total / t_total,
        "config": {"workload": "C5-tube (BASELINE configs[4] trajectories, configs[1] section): EM L1 halo (Az=0.2 S) "
                               "stable-manifold tube, DOP853 rtol=atol=1e-12, backward tf=0.75*2pi, dense samples on "
                               "the dt=1e-3 grid (4713) + synodic detector y=0 / (x,z) / direction=-1; "
                               f"bounded sample of {sample} trajectories per step (CPU oracle port)"},
        "cpu_baseline": {"value": val, "unit": "RK steps/s", "cores": n_threads, "kind": "port",
                         "sample": f"{sample} trajectories per step, {args.steps} steps"},
        "e2e": {"value": val, "unit": "RK steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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
    for n_seeds, label in ((100_000, "cm_map_tao4_1e5_seeds"), (1_000_000, "cm_map_tao4_1e6_seeds")):
        seeds = torch.from_numpy(g["seeds_p3"][rng.integers(0, 512, n_seeds)]).cuda()

        def run_cm():
            hold["r"] = cm.poincare_map(tab, seeds, opts)

        for _ in range(3):
            run_cm()                                 # first call compiles the specialised kernel (NVRTC)
        t = time_steps(run_cm, steps, flush, barrier, torch)
        f, _, tt = hold["r"]
        cm_steps = float((tt / 0.01).ceil().sum().item())
        out[label] = {"steps_per_s": cm_steps * steps / t, "crossings_per_s": int(f.sum().item()) * steps / t,
                      "ms_per_return": 1e3 * t / steps, "tflops": cm_steps * steps * 8100.0 / t / 1e12}
    out["cm_map_tao4_1e5_seeds"]["note"] = ("critical-path bound: the slowest seed needs ~1300 sequential steps of "
                                            "~9 us (11.9 ms on its own); 1e6 seeds fill the machine")
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
    ap.add_argument("--arith", default="parity", choices=["parity", "fast"])
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--steps-capacity", type=int, default=160,
                    help="accepted steps per trajectory the hb_cr3bp_section2 scratch holds (0: fused hb_cr3bp_section)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary (propagate-only / fast) timings")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import hiten_b200 as hb
    from hiten_b200 import propagate as P
    from hiten_b200 import synodic

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
    ics, mu = build_ics(n, rank, world)
    integ = hb.make_integ(arith=args.arith)
    m = max(int(abs(TF) / GRID_DT) + 1, 100)                              # manifold.py:396-397 -> 4713
    t_eval = np.linspace(0.0, TF, m)
    sec = synodic.make_section("y", 0.0, ("x", "z"), -1)                  # configs[1]'s SynodicMap call
    runner = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), integ=integ, device=dev,
                                       steps_capacity=args.steps_capacity)
    y0_soa = torch.from_numpy(np.ascontiguousarray(ics.T)).to(dev)       # resident input [6, N]
    host_in = torch.from_numpy(ics).pin_memory()                          # e2e input  [N, 6]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2
    gather_yf = gather_cnt = None
    if world > 1 and rank == 0:
        gather_yf = [torch.empty((6, n), dtype=torch.float64, device=dev) for _ in range(world)]
        gather_cnt = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(world)]

    def step_resident():
        runner.launch(y0_soa)
        if world > 1:                                   # the one exchange: end states + hit counts to rank 0
            dist.gather(runner.yf, gather_yf, dst=0)
            dist.gather(runner.per, gather_cnt, dst=0)

    def make_e2e():
        # the public streaming API: host batches in (pinned), host results out (pinned), double-buffered copies
        return synodic.TubeSectionStream(n, mu, t_eval, sec, forward=-1, flip=(0, 6), integ=integ, device=dev,
                                         steps_capacity=args.steps_capacity, scratch=runner.scratch)

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
    n_hits = runner.hit_count()
    steps_acc = int(runner.nacc.sum().item())
    steps_per_pass = steps_acc + int(runner.nrej.sum().item())
    ok = bool((runner.status == 0).all().item())
    t_kernel_local = t_dev

    # per-kernel split of the pipeline (CUDA events recorded inside hb_cr3bp_section2, same stream), separate pass
    stage_ms = None
    if args.steps_capacity > 0:
        import ctypes
        lib = runner.lib
        lib.hb_section2_profile(1)
        acc = np.zeros(4)
        buf = (ctypes.c_float * 4)()
        for i in range(args.steps):
            flush.fill_(float(i))
            runner.launch(y0_soa)
            lib.hb_section2_read_profile(buf)
            acc += np.array(list(buf))
        lib.hb_section2_profile(0)
        stage_ms = (acc / args.steps).tolist()

    # e2e: host buffers in, host results out, every step's H2D and D2H inside the timed region (overlapped with the
    # neighbouring steps' compute by TubeSectionStream's double buffering)
    stream_api = make_e2e()
    for r in stream_api.run([host_in] * 3):
        pass
    barrier()
    e2e_t0 = time.perf_counter()
    k_e2e = 0
    for r in stream_api.run([host_in] * args.steps):
        k_e2e = r.n_hits
    barrier()
    e2e_t = time.perf_counter() - e2e_t0
    e2e_ok = bool(k_e2e == n_hits and (r.status == 0).all())
    del stream_api

    tt = torch.tensor([t_dev, e2e_t, float(steps_per_pass), float(n_hits), float(steps_acc)], dtype=torch.float64,
                      device=dev)
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_dev, e2e_t = tmax[0].item(), tmax[1].item()
        total_steps_pass, total_hits, total_acc = tsum[2].item(), tsum[3].item(), tsum[4].item()
    else:
        total_steps_pass, total_hits, total_acc = float(steps_per_pass), float(n_hits), float(steps_acc)

    extra = {}
    if not args.no_extra and world == 1:
        # secondary numbers that explain the headline: propagation only (end states), both arithmetic variants,
        # and the fused step in the other variant
        ws = P.workspace(dev)
        for name in ("parity", "fast"):
            ig = hb.make_integ(arith=name)
            hold = {}

            def prop():
                hold["r"] = hb.cr3bp_propagate(y0_soa, mu, TF, forward=-1, flip=(0, 6), integ=ig, ws=ws)

            for _ in range(3):
                prop()
            tp = time_steps(prop, args.steps, flush, barrier, torch)
            sp = int((hold["r"].n_acc.sum() + hold["r"].n_rej.sum()).item())
            extra[f"propagate_only_{name}"] = {"rk_steps_per_s": sp * args.steps / tp,
                                               "tflops": sp * args.steps * FLOP_PER_STEP / tp / 1e12}
        other = "fast" if args.arith == "parity" else "parity"
        for label, ar, cap in ((f"section_{other}", other, args.steps_capacity),
                               ("section_fused_kernel_parity", "parity", 0), ("section_fused_kernel_fast", "fast", 0)):
            r2 = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), integ=hb.make_integ(arith=ar),
                                           device=dev, steps_capacity=cap, scratch=runner.scratch if cap > 0 else None)
            for _ in range(3):
                r2.launch(y0_soa)
            t2 = time_steps(lambda: r2.launch(y0_soa), args.steps, flush, barrier, torch)
            s2 = int((r2.nacc.sum() + r2.nrej.sum()).item())
            extra[label] = {"rk_steps_per_s": s2 * args.steps / t2, "ms_per_step": 1e3 * t2 / args.steps,
                            "crossings_per_s": r2.hit_count() * args.steps / t2}
            del r2
        if args.steps_capacity > 0:
            # the same step with Manifold.compute()'s trajectory filters judged from the step records:
            # all 4713 samples of every trajectory are rebuilt and tested (safe radii, Jacobi drift)
            r3 = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), integ=integ, device=dev,
                                           steps_capacity=args.steps_capacity, scratch=runner.scratch,
                                           filters=(3.318e-05, 9.04e-06, 1e-6))
            for _ in range(3):
                r3.launch(y0_soa)
            t3 = time_steps(lambda: r3.launch(y0_soa), args.steps, flush, barrier, torch)
            s3 = int((r3.nacc.sum() + r3.nrej.sum()).item())
            kept = int((r3.filter_result()[1] == 1).sum().item())
            filt_ms = 1e3 * t3 / args.steps - 1e3 * t_dev / args.steps
            extra[f"section_{args.arith}_with_trajectory_filters"] = {
                "rk_steps_per_s": s3 * args.steps / t3, "ms_per_step": 1e3 * t3 / args.steps,
                "filter_kernel_ms": filt_ms, "samples_per_s": n * float(m) / (filt_ms * 1e-3) if filt_ms > 0 else None,
                "kept_trajectories": kept,
                "note": "hb_section2_filter: 6-component dense evaluation + r1, r2, Jacobi constant at every grid sample"}
            del r3
        extra.update(secondary_configs(hb, torch, args.steps, flush, barrier))

    if rank == 0:
        value = total_steps_pass * args.steps / t_dev
        e2e_value = total_steps_pass * args.steps / e2e_t
        peak = hb.dfma_peak(200.0)
        # Roofline of the DOMINANT kernel (k_dop853_6 in record mode + its first-step pre-pass): attempted steps x
        # 1350 algorithmic flop over that kernel's own duration.  With the fused kernel (--steps-capacity 0) the
        # whole step is one kernel and the accepted steps' dense caches (1050 flop) are added.
        if stage_ms is not None:
            achieved = steps_per_pass * FLOP_PER_STEP / (stage_ms[0] * 1e-3)
            hbm = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"] \
                if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else 6500.0
            rec_bytes = steps_acc * 512.0
            extra["pipeline"] = {
                "stage_ms": dict(zip(("propagate_record", "step_scan", "emit_candidates", "order_dedup"), stage_ms)),
                "share_of_step_dominant": stage_ms[0] / sum(stage_ms),
                "propagate_record_hbm_write_gbs": rec_bytes / (stage_ms[0] * 1e-3) / 1e9,
                "step_scan_hbm_read_gbs": rec_bytes / (stage_ms[1] * 1e-3) / 1e9,
                "step_scan_frac_of_hbm_peak": rec_bytes / (stage_ms[1] * 1e-3) / 1e9 / hbm,
                "hbm_peak_gbs": hbm,
            }
        else:
            achieved = (steps_per_pass * FLOP_PER_STEP + steps_acc * FLOP_PER_SEGMENT) * args.steps / t_kernel_local
        for v in extra.values():
            if "tflops" in v:
                v["frac_of_fp64_peak"] = v["tflops"] * 1e12 / peak
        line = {
            "metric": "fp64 CR3BP RK steps/s", "value": value, "unit": "RK steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "crossings_per_s": total_hits * args.steps / t_dev,
            "config": {
                "workload": "C5-tube (BASELINE configs[4] = 1e6 trajectories per GPU, configs[1] section): EM L1 halo (Az=0.2 S) "
                            "stable-manifold tube, 2000 nodes x log-spaced displacements, DOP853 rtol=atol=1e-12, "
                            "backward tf=0.75*2pi, dense samples on the dt=1e-3 grid (4713) streamed through the "
                            "synodic detector y=0 / (x,z) / direction=-1 (segment_refine=50), hits + end states out",
                "trajectories_per_gpu": n, "grid_samples": m, "arith": args.arith,
                "path": "hb_cr3bp_section2 (propagate+record -> step scan -> emit -> order+dedup)"
                        if args.steps_capacity > 0 else "hb_cr3bp_section (fused kernel)",
                "steps_capacity": args.steps_capacity,
                "l2": "flushed between timed iterations (256 MB fill); inputs 48 B and step records ~44 KB per trajectory (>> L2)",
                "rk_steps_per_pass": total_steps_pass, "accepted_steps_per_pass": total_acc,
                "crossings_per_pass": total_hits, "all_status_ok": ok,
            },
            "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         "traffic": (4.437e10 if (n == N_PER_GPU and args.steps_capacity > 0 and args.arith == "parity") else None),
                         "kernel": "k_dop853_6<MODE_RECORD> (+ k_first_steps)" if stage_ms is not None
                                   else "k_dop853_6_section",
                         "note": "FP64 FMA pipe roofline of the dominant kernel (its HBM side: 512 B written per "
                                 "accepted step = 44.3 GB per launch at ~2.2 TB/s, a third of the HBM roof; traffic = "
                                 "dram read+write of one ncu --set full capture, profiles/r01_pipeline_v5_summary.md). "
                                 "achieved = attempted steps x 1350 algorithmic flop (SURVEY 8d) / the kernel's own "
                                 "duration (CUDA events recorded between the pipeline's kernels on the launching "
                                 "stream); peak = hb_dfma_peak measured in this process"},
            "e2e": {"value": e2e_value, "unit": "RK steps/s", "h2d_bytes_per_step": int(n * 48),
                    "d2h_bytes_per_step": int(n * (48 + 16) + 72 * k_e2e), "same_hits_as_resident_run": e2e_ok,
                    "api": "synodic.TubeSectionStream.run (pinned host batches in, pinned host results out)",
                    "crossings_per_s": total_hits * args.steps / e2e_t},
            "gpu_launches": world * args.steps * ((6 if args.arith == "parity" else 5) if args.steps_capacity > 0
                                          else (2 if args.arith == "parity" else 1)),
            "clocks": clocks, "wall_s_timed_region": wall, "extra": extra,
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle_lib as O
            nthr = O.lib().ho_max_threads()
            v, ns, dt = cpu_port_throughput(ics, mu, nthr)
            line["cpu_baseline"] = {"value": v, "unit": "RK steps/s", "cores": nthr, "kind": "port",
                                    "sample": f"{ns} trajectories of the same batch (tube propagation + dense grid + "
                                              f"section detection per trajectory), {dt:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
