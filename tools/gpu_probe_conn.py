"""Timing of hb_connections on config-5-sized hit sets (2e6 x 2e6 section points)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from hiten_b200 import connections as cn
rng = np.random.default_rng(7)
for n, eps in ((200_000, 5e-4), (2_000_000, 1.5e-4)):
    pu = rng.uniform(-0.4, 0.4, (n, 2))
    ps = np.vstack((pu[rng.choice(n, n // 2, replace=False)] + rng.uniform(-1, 1, (n // 2, 2)) * eps * 0.8,
                    rng.uniform(-0.4, 0.4, (n - n // 2, 2))))
    Xu, Xs = rng.normal(0, 0.2, (n, 6)), rng.normal(0, 0.2, (n, 6))
    d = [torch.from_numpy(a).cuda() for a in (pu, ps, Xu, Xs)]
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = cn.find_connections(*d, eps, 0.5, 1e-3)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"n_u = n_s = {n}: {1e3 * dt:.1f} ms, pairs considered {r.pairs_considered}, accepted {len(r.delta_v)}")
