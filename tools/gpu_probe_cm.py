"""Timing probe: centre-manifold map (config 3, 1e5 seeds) and 42-state STM (config 4 x128)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hiten_b200 as hb
from hiten_b200 import centermanifold as cm

g = np.load("tests/golden/cm_map.npz")
tab = cm.PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
rng = np.random.default_rng(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
seeds = torch.from_numpy(g["seeds_p3"][rng.integers(0, 512, n)]).cuda()
for JIT in (True, False):
  for method, order in (("symplectic", 4), ("fixed", 4), ("symplectic", 6), ("fixed", 8)):
      for arith in ("parity", "fast"):
          opts = cm.make_opts(0.01, 2000, method, order, "p3", 20.0, arith)
          best = 1e9
          for rep in range(4):
              e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
              e0.record(); f, o, t = cm.poincare_map(tab, seeds, opts, jit=JIT); e1.record(); torch.cuda.synchronize()
              if rep: best = min(best, e0.elapsed_time(e1))
          steps = float((t / 0.01).ceil().sum().item())
          print(json.dumps({"jit": JIT, "cm": method, "order": order, "arith": arith, "n": n, "ms": round(best, 2),
                            "steps_per_s": steps / best * 1e3, "crossings_per_s": int(f.sum().item()) / best * 1e3}))
s = np.load("tests/golden/stm_family.npz")
reps = 128
x0 = torch.from_numpy(np.ascontiguousarray(np.tile(s["x0"], (reps, 1)).T)).cuda()
T = torch.from_numpy(np.tile(s["period"], reps)).cuda()
for arith in ("parity", "fast"):
    integ = hb.make_integ(arith=arith)
    best = 1e9
    for rep in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = hb.cr3bp_stm(x0, float(s["mu"]), 0.0, tf_per_traj=T, integ=integ); e1.record(); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    st = int((r.n_acc.sum() + r.n_rej.sum()).item())
    print(json.dumps({"stm": arith, "n": x0.shape[1], "ms": round(best, 2), "steps_per_s": st / best * 1e3,
                      "tflops_alg": st * 9500 / best * 1e3 / 1e12}))
