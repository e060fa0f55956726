import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hiten_b200 as hb
from hiten_b200 import synodic
g = np.load("tests/golden/c5_connection.npz")
mu, tf, steps = float(g["mu"]), float(g["l2_tf"]), int(g["l2_steps"])
t_eval = np.linspace(0.0, tf, steps)
sec = synodic.make_section("y", 0.0, ("x", "z"), 0)
x0 = np.tile(g["l2_x0W"], (int(sys.argv[1]) if len(sys.argv) > 1 else 1, 1))
y0 = torch.from_numpy(np.ascontiguousarray(x0.T)).cuda()
res = {}
for name, kw in (("all", dict(steps_capacity=192, records="all")), ("near", dict(steps_capacity=192, records="near")),
                 ("s3", dict(pool_records=16)), ("fused", dict())):
    r = synodic.TubeSectionRunner(len(x0), mu, t_eval, sec, forward=1, **kw)
    r.launch(y0)
    res[name] = r.sorted_hits()
    print(name, len(res[name].times), int((r.status == 4).sum()))
dense = hb.cr3bp_dense(x0, mu, t_eval, forward=1, keep_on_device=True)
c = synodic.detect(dense.states, t_eval, sec)
print("dense chain", len(c.times))
for name, h in res.items():
    same = len(h.times) == len(c.times) and np.array_equal(h.times, c.times) and np.array_equal(h.trajectory_indices, c.trajectory_indices)
    print(name, "== dense chain:", same)
    if not same:
        a = set(zip(h.trajectory_indices.tolist(), h.times.tolist())); b = set(zip(c.trajectory_indices.tolist(), c.times.tolist()))
        extra, missing = sorted(a - b)[:5], sorted(b - a)[:5]
        print("  extra", extra, "missing", missing, "n extra", len(a - b), "n missing", len(b - a))
        tr = sorted({t for t, _ in (a - b)})
        print("  trajectories with extras:", len(tr), tr[:20], "mod 200:", sorted({t % 200 for t in tr})[:30])
