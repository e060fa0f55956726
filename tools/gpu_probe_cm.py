"""Probe: one return of the centre-manifold map (Tao order 4, dt = 0.01, section p3) for several seed counts (seeds of the
golden set resampled with small distinct offsets are NOT used: lifted seeds like bench.py's).  usage: gpu_probe_cm.py n1 n2 ..."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import hiten_b200 as hb
from hiten_b200 import centermanifold as cm
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cm_map.npz"))
tab = cm.PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])
Htab = cm.PolyTable.from_blocks(g["H_deg"], g["H_coef"], g["H_exp"]) if hasattr(cm.PolyTable, "from_blocks") else None
opts = cm.make_opts(0.01, 2000, "symplectic", 4, "p3", 20.0, "parity")
rng = np.random.default_rng(0)
base = g["seeds_p3"]
for n in [int(a) for a in sys.argv[1:]] or [2000, 10000, 100000]:
    # distinct seeds: the 512 golden seeds scaled towards the origin by distinct factors (they stay inside the Hill region)
    f = 1.0 - 0.5 * rng.random((n, 1))
    seeds = torch.from_numpy(base[rng.integers(0, len(base), n)] * f).cuda()
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        flags, out, tt = cm.poincare_map(tab, seeds, opts)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    steps = float((tt[flags.bool()] / 0.01).sum().item()) if hasattr(tt, "sum") else 0.0
    print({"n": n, "ms": 1e3 * best, "max_steps": float(tt.max().item() / 0.01), "ok": float(flags.float().mean().item())})
