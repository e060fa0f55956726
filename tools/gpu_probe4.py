"""How expensive is 'propagate + dense cache on (almost) every step'?  Dense mode on coarse grids."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
import hiten_b200 as hb
from hiten_b200 import propagate as P

n = 131072
ics, mu = bench.build_ics(n)
y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
ws = P.workspace(y0.device)
for arith in ("parity", "fast"):
    integ = hb.make_integ(arith=arith)
    for m in (2, 50, 200, 400):
        te = torch.linspace(0.0, bench.TF, m, dtype=torch.float64).cuda()
        best = 1e9
        for rep in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); r = hb.cr3bp_dense(y0, mu, te, forward=-1, flip=(0, 6), integ=integ, ws=ws); e1.record()
            torch.cuda.synchronize()
            if rep: best = min(best, e0.elapsed_time(e1))
        del r
        print(json.dumps({"arith": arith, "grid": m, "ms": round(best, 3)}))
