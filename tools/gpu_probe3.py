"""Timing of propagate + fused-section kernels (131072 trajectories) for kernel-variant comparisons."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
import hiten_b200 as hb
from hiten_b200 import propagate as P, synodic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
ics, mu = bench.build_ics(n)
y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
ws = P.workspace(y0.device)
m = max(int(abs(bench.TF) / bench.GRID_DT) + 1, 100)
t_eval = np.linspace(0.0, bench.TF, m)
sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
for arith in ("parity", "fast"):
    integ = hb.make_integ(arith=arith)
    run = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), integ=integ, steps_capacity=int(os.environ.get("HB_STEPS_CAP", "0")))
    for kind in ("propagate", "section"):
        best = 1e9
        for rep in range(6):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            if kind == "propagate":
                r = hb.cr3bp_propagate(y0, mu, bench.TF, forward=-1, flip=(0, 6), integ=integ, ws=ws)
            else:
                run.launch(y0)
            e1.record(); torch.cuda.synchronize()
            if rep >= 2: best = min(best, e0.elapsed_time(e1))
        steps = int((r.n_acc.sum() + r.n_rej.sum()).item()) if kind == "propagate" else int((run.nacc.sum() + run.nrej.sum()).item())
        print(json.dumps({"lib": os.environ.get("HITEN_B200_LIB", "default"), "kind": kind, "arith": arith, "ms": round(best, 3),
                          "steps_per_s": steps / (best * 1e-3), "hits": run.hit_count() if kind == "section" else None}))
print(json.dumps({"nacc_max": int(run.nacc.max().item()), "nacc_mean": float(run.nacc.float().mean().item()),
                  "att_std": float((run.nacc + run.nrej).float().std().item()), "status_nonzero": int((run.status != 0).sum().item())}))
