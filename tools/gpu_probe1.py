"""First GPU probe: timing of the propagate kernel (parity / fast) on replicated C1 ICs + DFMA peak."""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hiten_b200 as hb
from hiten_b200 import propagate as P

g = np.load("tests/golden/c1_manifold.npz")
mu, tf = float(g["mu"]), float(g["tf"])
print("dfma peak TF/s:", hb.dfma_peak(100.0) / 1e12, hb.dfma_peak(300.0) / 1e12)
for n in (50, 4096, 65536, 262144, 1048576):
    y0 = np.tile(g["x0W"], (n // 50 + 1, 1))[:n]
    y0d = torch.from_numpy(np.ascontiguousarray(y0.T)).cuda()
    for arith in ("parity", "fast"):
        integ = hb.make_integ(arith=arith)
        ws = P.workspace(y0d.device)
        for rep in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            r = hb.cr3bp_propagate(y0d, mu, tf, forward=-1, flip=(0, 6), integ=integ, ws=ws)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        steps = int(r.n_acc.sum().item() + r.n_rej.sum().item())
        print(json.dumps({"n": n, "arith": arith, "ms": ms, "steps": steps, "steps_per_s": steps / (ms * 1e-3),
                          "tflops_alg": steps * 1350 / (ms * 1e-3) / 1e12}))
