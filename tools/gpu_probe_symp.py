"""Throughput probe of hb_ham_symplectic_dense (table-driven Tao integrator over a grid) on one GPU."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from hiten_b200 import symplectic as S  # noqa: E402
from hiten_b200.centermanifold import PolyTable  # noqa: E402

g = np.load(os.path.join(REPO, "tests", "golden", "cm_map.npz"))
tab = PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
rng = np.random.default_rng(0)
for n, m, order in ((20000, 101, 4), (100000, 101, 4), (100000, 51, 6), (100000, 201, 2)):
    seeds = g["seeds_p3"][rng.integers(0, len(g["seeds_p3"]), n)]
    y0 = np.zeros((n, 6))
    y0[:, 1], y0[:, 4], y0[:, 2], y0[:, 5] = seeds[:, 0], seeds[:, 1], seeds[:, 2], seeds[:, 3]
    yd = torch.from_numpy(y0).cuda()
    t = np.linspace(0.0, 0.01 * (m - 1), m)
    for arith in ("parity", "fast"):
        S.integrate_symplectic(tab, yd, t, order, arith=arith)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        S.integrate_symplectic(tab, yd, t, order, arith=arith)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"n={n} m={m} order={order} {arith}: {ms:.2f} ms, {n * (m - 1) / ms * 1e3:.3e} Tao steps/s", flush=True)
