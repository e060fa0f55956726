"""Timing of the specialised centre-manifold map (BASELINE config 3: Tao-4, dt = 0.01, one return) at 1e4 / 1e5 / 1e6 seeds."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from hiten_b200 import centermanifold as cm  # noqa: E402

g = np.load(os.path.join(REPO, "tests", "golden", "cm_map.npz"))
tab = cm.PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
rng = np.random.default_rng(1)
for arith in ("parity", "fast"):
    opts = cm.make_opts(0.01, 2000, "symplectic", 4, "p3", 20.0, arith)
    for n in (10_000, 100_000, 1_000_000):
        seeds = torch.from_numpy(g["seeds_p3"][rng.integers(0, 512, n)]).cuda()
        for _ in range(2):
            f, o, tt = cm.poincare_map(tab, seeds, opts)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f, o, tt = cm.poincare_map(tab, seeds, opts)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        steps = float((tt / 0.01).ceil().sum().item())
        print(f"{arith} n={n}: {best:.2f} ms, {steps / best * 1e3:.3e} Tao-4 steps/s, hits {int(f.sum().item())}, "
              f"checksum {float(o.double().sum().item()):.15e}", flush=True)
