"""Timing probe for hb_manifold_ics / hb_tube_filter (CUDA events, inputs > L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hiten_b200 import manifold

def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

rng = np.random.default_rng(0)
S = 2000
phi = torch.from_numpy(rng.standard_normal((S, 42))).cuda()
tt = torch.from_numpy(np.linspace(0, 2.75, S)).cuda()
fr = torch.from_numpy(np.arange(0.0, 1.0, 0.0005)).cuda()
dd = torch.from_numpy(np.logspace(-7, -5, 500)).cuda()
ev = rng.standard_normal(6)
ms = timeit(lambda: manifold.tube_initial_conditions(phi, tt, 2.75, ev, 1, fr, dd))
print(f"hb_manifold_ics 2000 x 500 = 1e6 ICs: {ms:.3f} ms ({1e6 / ms * 1e3:.3e} ICs/s)")
for n, m in ((16384, 4713), (131072, 1024)):
    s = torch.randn((n, m, 6), dtype=torch.float64, device="cuda")
    ms = timeit(lambda: manifold.tube_filter(s, 0.01215, safe_r1=1e-5, safe_r2=1e-5, energy_tol=1e-6), 5)
    gb = n * m * 48 / 1e9
    print(f"hb_tube_filter {n} x {m} samples ({gb:.2f} GB): {ms:.3f} ms = {gb / ms * 1e3:.0f} GB/s, {n * m / ms * 1e3:.3e} samples/s")
    del s
