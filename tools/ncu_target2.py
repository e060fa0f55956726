import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
import hiten_b200 as hb
from hiten_b200 import synodic
arith = sys.argv[1] if len(sys.argv) > 1 else "parity"
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_PER_GPU
ics, mu = bench.build_ics(n)
y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
m = max(int(abs(bench.TF) / bench.GRID_DT) + 1, 100)
run = synodic.TubeSectionRunner(n, mu, np.linspace(0.0, bench.TF, m), synodic.make_section("y", 0.0, ("x", "z"), -1),
                                forward=-1, flip=(0, 6), integ=hb.make_integ(arith=arith), steps_capacity=160)
for _ in range(1):
    run.launch(y0)
torch.cuda.synchronize()
print(run.hit_count(), int((run.status != 0).sum()))
