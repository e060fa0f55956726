"""Kernel-variant timing on the bench workload (131072 trajectories)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
import hiten_b200 as hb
from hiten_b200 import propagate as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
ics, mu = bench.build_ics(n)
y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
ws = P.workspace(y0.device)
for arith in ("parity", "fast"):
    integ = hb.make_integ(arith=arith)
    best = 1e9
    for rep in range(6):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        r = hb.cr3bp_propagate(y0, mu, bench.TF, forward=-1, flip=(0, 6), integ=integ, ws=ws)
        e1.record(); torch.cuda.synchronize()
        if rep >= 2: best = min(best, e0.elapsed_time(e1))
    steps = int(r.n_acc.sum().item() + r.n_rej.sum().item())
    print(json.dumps({"lib": os.environ.get("HITEN_B200_LIB", "default"), "n": n, "arith": arith, "ms": best,
                      "steps_per_s": steps / (best * 1e-3), "tflops_alg": steps * 1350 / (best * 1e-3) / 1e12}))
