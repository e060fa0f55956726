"""ncu driver: one launch of the specialised CM-map kernel (Tao-4, 1e5 seeds) and of the STM kernel (12800 traj)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hiten_b200 as hb
from hiten_b200 import centermanifold as cm
g = np.load("tests/golden/cm_map.npz")
tab = cm.PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
rng = np.random.default_rng(1)
seeds = torch.from_numpy(g["seeds_p3"][rng.integers(0, 512, 100_000)]).cuda()
opts = cm.make_opts(0.01, 2000, "symplectic", 4, "p3", 20.0, "parity")
for _ in range(2):
    cm.poincare_map(tab, seeds, opts)
s = np.load("tests/golden/stm_family.npz")
x0 = torch.from_numpy(np.ascontiguousarray(np.tile(s["x0"], (128, 1)).T)).cuda()
T = torch.from_numpy(np.tile(s["period"], 128)).cuda()
for _ in range(2):
    hb.cr3bp_stm(x0, float(s["mu"]), 0.0, tf_per_traj=T, integ=hb.make_integ(arith="parity"))
torch.cuda.synchronize()
