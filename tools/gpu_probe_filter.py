"""Timing probe: section pipeline with and without the record filter (hb_section2_filter)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import hiten_b200 as hb
from hiten_b200 import synodic, manifold

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
ics, mu = bench.build_ics(n)
m = max(int(abs(bench.TF) / bench.GRID_DT) + 1, 100)
t_eval = np.linspace(0.0, bench.TF, m)
sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()

def timeit(fn, k=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

for arith in ("parity", "fast"):
    integ = hb.make_integ(arith=arith)
    a = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=160, integ=integ)
    ta = timeit(lambda: a.launch(y0))
    b = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=160, integ=integ,
                                  scratch=a.scratch, filters=(3.318e-05, 9.04e-06, 1e-6))
    tb = timeit(lambda: b.launch(y0))
    print(f"{arith}: n={n} pipeline {ta:.2f} ms, with filter {tb:.2f} ms -> filter {tb - ta:.2f} ms "
          f"({n * m / (tb - ta) * 1e3:.3e} samples/s)", flush=True)
tube = torch.randn((16384, 4713, 6), dtype=torch.float64, device="cuda")
tf = timeit(lambda: manifold.tube_filter(tube, mu, safe_r1=1e-5, safe_r2=1e-5, energy_tol=1e-6))
print(f"hb_tube_filter 16384 x 4713: {tf:.3f} ms = {tube.numel() * 8 / tf / 1e6:.0f} GB/s")
