"""CPU study of the launch order of the persistent DOP853 kernel (hb_integ.order), no GPU needed.

The C oracle's step counts equal the kernel's (parity arithmetic is bit-exact), so the attempted steps of every trajectory of
a configs[4] tube are computed on the host, and the persistent queue (148 SMs x 256 lanes, a lane takes the next trajectory when
it finishes one, every attempted step costs one loop iteration) is simulated for several hand-out orders:
  natural       the workload generator's order (displacement-major, node-minor);
  exact         longest first with the true costs (what TubeSectionRunner.order_by_cost() has after a pass over the same batch);
  node mean     longest first with a cost known per orbit node only;
  pilot         longest first with a cost MODEL that needs no earlier pass: the true costs of every K-th displacement row
                (a pilot of 2000 x ceil(250 / K) trajectories), interpolated linearly in the displacement index.
Prints the makespan of each order in loop iterations and relative to the perfect balance (sum / lanes), and the number of
warp-level refill events (with exact costs the lanes of a warp hold equally long trajectories and refill together).
usage: python tools/sim_launch_order.py [n_per_tube] [threads]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle_lib as O
from hiten_b200 import workloads as W

n_tube = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
LANES = 148 * 256


def simulate(cost, order, lanes=LANES):
    """-> (loop iterations until the last lane is done, WARP-LEVEL refill events: iterations in which at least one lane of a
    warp takes a new trajectory -- each is a divergent ~1.5 us refill the warp's other lanes wait for)."""
    c = cost[order]
    n = len(c)
    rem = np.zeros(lanes, dtype=np.int64)
    active = np.zeros(lanes, dtype=bool)
    nxt = it = events = 0
    while True:
        need = np.flatnonzero(~active)
        if nxt < n and len(need):
            got = need[: min(len(need), n - nxt)]
            rem[got] = c[nxt:nxt + len(got)]
            active[got] = True
            nxt += len(got)
            events += len(np.unique(got // 32))
        if not active.any():
            return float(it), events
        m = rem[active].min()                     # jump to the next finishing time
        rem[active] -= m
        it += m
        active &= rem != 0


def makespan(cost, order):
    return simulate(cost, order)


ics, mu = W.c5_batch(2 * n_tube)
for key in ("l1", "l2"):
    x = ics[key]
    fwd = W.C5_TUBES[key]["forward"]
    sys_ = O.system(O.SYS_CR3BP6, mu=mu, fwd=fwd, flip=(0, 6) if fwd < 0 else None)
    _, counts = O.batch_final(sys_, O.DOP853, O.default_tol(), x, 0.0, float(W.c5_grid(key)[-1]), n_threads=threads)
    cost = counts.sum(axis=1).astype(np.int64)
    n = len(cost)
    node = np.arange(n) % 2000
    row = np.arange(n) // 2000
    n_rows = int(row.max()) + 1
    ideal = cost.sum() / LANES
    res = {"natural": makespan(cost, np.arange(n)),
           "exact": makespan(cost, np.argsort(-cost, kind="stable"))}
    mean = np.bincount(node, weights=cost) / np.bincount(node)
    res["node mean"] = makespan(cost, np.argsort(-mean[node], kind="stable"))
    grid = np.full((n_rows, 2000), np.nan)
    grid[row, node] = cost
    for K in (40, 20, 10):
        pil = np.unique(np.concatenate((np.arange(0, n_rows, K), [n_rows - 1])))
        model = np.empty((n_rows, 2000))
        for j in range(2000):
            model[:, j] = np.interp(np.arange(n_rows), pil, np.nan_to_num(grid[pil, j], nan=np.nanmean(grid[pil])))
        pred = model[row, node]
        expl = 1.0 - ((cost - pred) ** 2).mean() / cost.var()
        res[f"pilot 1/{K} rows ({len(pil) * 2000} trajectories, R2 {expl:.3f})"] = makespan(cost, np.argsort(-pred, kind="stable"))
    print(f"{key}: {n} trajectories, attempted steps min / mean / max {cost.min()} / {cost.mean():.1f} / {cost.max()}, "
          f"perfect balance {ideal:.0f} iterations per lane")
    for k, (v, ev) in res.items():
        print(f"   {k:58s} makespan {v:8.0f} = {v / ideal:.3f} x perfect   ({100 * (v / res['natural'][0] - 1):+.1f} % vs natural)"
              f"   warp refill events {ev}")
