import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from hiten_b200 import centermanifold as cm
g = np.load("tests/golden/cm_map.npz")
tab = cm.PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
rng = np.random.default_rng(1)
opts = cm.make_opts(0.01, 2000, "symplectic", 4, "p3", 20.0, "parity")
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
for n in (100_000, 1_000_000):
    seeds = torch.from_numpy(g["seeds_p3"][rng.integers(0, 512, n)]).cuda()
    ts = []
    for rep in range(6):
        flush.fill_(float(rep))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f, o, t = cm.poincare_map(tab, seeds, opts); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 2))
    print(n, ts, float((t / 0.01).ceil().sum().item()) / (min(ts) * 1e-3))
