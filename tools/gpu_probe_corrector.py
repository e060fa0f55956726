"""Throughput probe for hb_correct_orbits: N perturbed halo guesses in one lock-step batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hiten_b200 import corrector

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "correction.npz"))
mu = float(g["mu"])
rng = np.random.default_rng(1)
for fam in ("halo", "lyapunov", "vertical"):
    base = g[f"{fam}_x0"][g[f"{fam}_iters"] >= 0]
    for n in (100, 10000, 100000):
        x0 = base[rng.integers(0, len(base), n)].copy()
        x0[:, corrector.FAMILIES[fam][0]] += 1e-4 * rng.standard_normal((n, 2))
        xd = torch.from_numpy(np.ascontiguousarray(x0.T)).cuda()
        opts = corrector.make_opts(fam)
        corrector.correct_orbits(xd, mu, opts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = corrector.correct_orbits(xd, mu, opts)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = res.status.cpu().numpy()
        it = res.iterations.cpu().numpy()
        print(f"{fam:9s} n={n:6d}: {dt * 1e3:8.1f} ms  {n / dt:10.3e} orbits/s  converged {np.mean(st == 0):.3f} "
              f"iters mean {it.mean():.1f} max {it.max()}  steps6 {res.rk_steps6:.3e} steps42 {res.rk_steps42:.3e} "
              f"-> {(res.rk_steps6 + res.rk_steps42) / dt:.3e} RK steps/s", flush=True)
