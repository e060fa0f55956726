"""ncu target for the SURVEY 8f#3 / 8f#4 kernels: one filtered section pass (65536 trajectories), one stored-tube
filter, one batched halo correction (10000 orbits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from hiten_b200 import synodic, manifold, corrector

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "filter"):
    n = 65536
    ics, mu = bench.build_ics(n)
    m = max(int(abs(bench.TF) / bench.GRID_DT) + 1, 100)
    t_eval = np.linspace(0.0, bench.TF, m)
    sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
    y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
    r = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=160,
                                  filters=(3.318e-05, 9.04e-06, 1e-6))
    for _ in range(2):
        r.launch(y0)
    torch.cuda.synchronize()
    tube = torch.randn((4096, 4713, 6), dtype=torch.float64, device="cuda")
    for _ in range(2):
        manifold.tube_filter(tube, mu, safe_r1=1e-5, safe_r2=1e-5, energy_tol=1e-6)
    torch.cuda.synchronize()
if what in ("all", "correct"):
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "correction.npz"))
    rng = np.random.default_rng(1)
    base = g["halo_x0"]
    x0 = base[rng.integers(0, len(base), 10000)].copy()
    x0[:, [0, 4]] += 1e-4 * rng.standard_normal((10000, 2))
    res = corrector.correct_orbits(torch.from_numpy(np.ascontiguousarray(x0.T)).cuda(), float(g["mu"]),
                                   corrector.make_opts("halo"))
    torch.cuda.synchronize()
    print("converged", float((res.status == 0).float().mean()), "steps", res.rk_steps6, res.rk_steps42)
