"""GPU: the `fast` arithmetic variant (FMA-contracted, rsqrt-based field, fp32 step-size factor) against the REFERENCE's
golden vectors -- not against `parity`.  BASELINE.json's criterion: identical crossing counts, crossing points within
1e-9 in synodic coordinates, end states within 1e-9 relative.  The distributions are printed (pytest -s) and recorded in
DESIGN.md; what is asserted is what holds: counts and per-trajectory hit structure are identical on every golden tube,
the bulk of the points sits at ~1e-11, and the tail -- trajectories that pass close to a primary, where one ulp grows by
1e4-1e6 along the arc (the reference's own 1-ulp sensitivity: median 3e-11, max 1.25e-8) -- exceeds 1e-9.
So `fast` does NOT meet the 1e-9 criterion on every crossing and is reported as a secondary number only."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [("synodic_c1.npz", ""), ("synodic_c2.npz", ""), ("synodic_se.npz", ""), ("c5_connection.npz", "l1_"),
         ("c5_connection.npz", "l2_")]


def run_fast(g, pre):
    import torch
    import hiten_b200 as hb
    from hiten_b200 import synodic
    mu, tf, steps, fwd = float(g["mu"]), float(g[pre + "tf"]), int(g[pre + "steps"]), int(g[pre + "forward"])
    t_eval = np.linspace(0.0, tf, steps)
    axis = int(np.nonzero(g[pre + "req_normal"])[0][0])
    sec = synodic.make_section(axis, float(g[pre + "req_offset"]), tuple(str(c) for c in g[pre + "req_plane_coords"]),
                               int(g[pre + "req_direction"]), int(g[pre + "req_segment_refine"]),
                               float(g[pre + "req_tol_on_surface"]), float(g[pre + "req_dedup_time_tol"]),
                               float(g[pre + "req_dedup_point_tol"]))
    x0 = g[pre + "x0W"]
    run = synodic.TubeSectionRunner(len(x0), mu, t_eval, sec, forward=fwd, flip=(0, 6), steps_capacity=256,
                                    integ=hb.make_integ(arith="fast"))
    run.launch(torch.from_numpy(np.ascontiguousarray(x0.T)).cuda())
    h = run.sorted_hits()
    return h, run.yf.t().cpu().numpy()


@pytest.mark.parametrize("name,pre", CASES)
def test_fast_vs_reference_goldens(name, pre):
    g = np.load(os.path.join(HERE, "golden", name))
    h, yf = run_fast(g, pre)
    kept = g[pre + "kept"] if (pre + "kept") in g.files else np.ones(len(yf), dtype=bool)
    sel = kept[h.trajectory_indices]
    traj = (np.cumsum(kept) - 1)[h.trajectory_indices[sel]]
    pts, times = h.points[sel], h.times[sel]
    ref_traj, ref_pts, ref_t = g[pre + "hit_traj"], g[pre + "hit_point"], g[pre + "hit_time"]
    rel = np.linalg.norm(yf - g[pre + "yf"], axis=1) / np.linalg.norm(g[pre + "yf"], axis=1)
    same_count = len(times) == len(ref_t) and np.array_equal(traj, ref_traj)
    line = f"[fast] {name}{' ' + pre if pre else ''}: crossings {len(times)} vs {len(ref_t)} (per-trajectory structure " \
           f"{'identical' if same_count else 'DIFFERENT'}); end states rel median {np.median(rel):.2e} p90 " \
           f"{np.percentile(rel, 90):.2e} max {rel.max():.2e} (> 1e-9: {(rel > 1e-9).sum()}/{len(rel)})"
    if same_count:
        dp = np.abs(pts - ref_pts).max(axis=1)
        line += f"; |d point| median {np.median(dp):.2e} p90 {np.percentile(dp, 90):.2e} max {dp.max():.2e} " \
                f"(> 1e-9: {(dp > 1e-9).sum()}/{len(dp)})"
    print(line)
    assert same_count                                   # identical crossing counts, trajectory by trajectory
    assert np.median(dp) <= 1e-9 and np.median(rel) <= 1e-9
    assert dp.max() <= 1e-5 and rel[kept].max() <= 1e-4  # the tail is bounded by the arc's own sensitivity, not garbage
