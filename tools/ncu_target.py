"""Small driver for ncu captures: a few launches of one kernel on the bench workload.
usage: ncu_target.py {propagate|section} {parity|fast} [n] [reps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
import hiten_b200 as hb
from hiten_b200 import propagate as P, synodic

kind = sys.argv[1] if len(sys.argv) > 1 else "section"
arith = sys.argv[2] if len(sys.argv) > 2 else "parity"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 131072
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ics, mu = bench.build_ics(n)
y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
integ = hb.make_integ(arith=arith)
if kind == "propagate":
    ws = P.workspace(y0.device)
    for _ in range(reps):
        r = hb.cr3bp_propagate(y0, mu, bench.TF, forward=-1, flip=(0, 6), integ=integ, ws=ws)
    torch.cuda.synchronize()
    print(kind, arith, n, int(r.n_acc.sum() + r.n_rej.sum()))
else:
    m = max(int(abs(bench.TF) / bench.GRID_DT) + 1, 100)
    run = synodic.TubeSectionRunner(n, mu, np.linspace(0.0, bench.TF, m), synodic.make_section("y", 0.0, ("x", "z"), -1),
                                    forward=-1, flip=(0, 6), integ=integ)
    for _ in range(reps):
        run.launch(y0)
    torch.cuda.synchronize()
    print(kind, arith, n, int(run.nacc.sum() + run.nrej.sum()), run.hit_count())
