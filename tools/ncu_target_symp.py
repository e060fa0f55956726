"""ncu target: the run-time specialised Tao grid kernel (symp_grid) on 1e5 trajectories x 100 intervals, parity."""
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
seeds = g["seeds_p3"][rng.integers(0, len(g["seeds_p3"]), 100000)]
y0 = np.zeros((100000, 6))
y0[:, 1], y0[:, 4], y0[:, 2], y0[:, 5] = seeds[:, 0], seeds[:, 1], seeds[:, 2], seeds[:, 3]
yd = torch.from_numpy(y0).cuda()
t = np.linspace(0.0, 1.0, 101)
for _ in range(3):
    S.integrate_symplectic(tab, yd, t, 4, arith=sys.argv[1] if len(sys.argv) > 1 else "parity")
torch.cuda.synchronize()
