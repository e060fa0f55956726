"""ncu driver: one return of the specialised CM-map kernel v3 (Tao-4, parity, 1e6 seeds; BASELINE config 3 scaled up)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from hiten_b200 import centermanifold as cm
g = np.load("tests/golden/cm_map.npz")
tab = cm.PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
rng = np.random.default_rng(1)
seeds = torch.from_numpy(g["seeds_p3"][rng.integers(0, 512, 1_000_000)]).cuda()
opts = cm.make_opts(0.01, 2000, "symplectic", 4, "p3", 20.0, "parity")
for _ in range(2):
    cm.poincare_map(tab, seeds, opts)
torch.cuda.synchronize()
