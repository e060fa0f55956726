import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hiten_b200 as hb
from hiten_b200 import synodic
g = np.load("tests/golden/synodic_c1.npz")
mu, tf, steps, fwd = float(g["mu"]), float(g["tf"]), int(g["steps"]), int(g["forward"])
t_eval = np.linspace(0.0, tf, steps)
sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
y0 = torch.from_numpy(np.ascontiguousarray(g["x0W"].T)).cuda()
for arith in ("parity", "fast"):
    integ = hb.make_integ(arith=arith)
    a = synodic.TubeSectionRunner(50, mu, t_eval, sec, forward=fwd, flip=(0, 6), integ=integ)
    b = synodic.TubeSectionRunner(50, mu, t_eval, sec, forward=fwd, flip=(0, 6), integ=integ, steps_capacity=256)
    a.launch(y0); b.launch(y0)
    ha, hb_ = a.sorted_hits(), b.sorted_hits()
    print(arith, len(ha.times), len(hb_.times), "status", b.status.cpu().numpy().max(), "per-traj equal", np.array_equal(ha.hits_per_traj, hb_.hits_per_traj))
    d = np.nonzero(ha.hits_per_traj != hb_.hits_per_traj)[0]
    print(" differing trajectories", d, ha.hits_per_traj[d], hb_.hits_per_traj[d])
    for i in d[:2]:
        print("  fused", ha.times[ha.trajectory_indices == i]); print("  two  ", hb_.times[hb_.trajectory_indices == i])
    print(" nacc equal", torch.equal(a.nacc, b.nacc), "cand", None)
