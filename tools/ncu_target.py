"""Small driver for ncu captures: a few launches of the propagate kernel on the bench workload."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
import hiten_b200 as hb
from hiten_b200 import propagate as P

arith = sys.argv[1] if len(sys.argv) > 1 else "parity"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ics, mu = bench.build_ics(n)
y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
ws = P.workspace(y0.device)
integ = hb.make_integ(arith=arith)
for _ in range(reps):
    r = hb.cr3bp_propagate(y0, mu, bench.TF, forward=-1, flip=(0, 6), integ=integ, ws=ws)
torch.cuda.synchronize()
print(arith, n, int(r.n_acc.sum() + r.n_rej.sum()))
