"""GPU parity of hb_cr3bp_section3 (step records handed from propagating to scanning warps through shared memory
inside one persistent kernel): the same hits, counts and end states as the reference (golden vectors) and as
hb_cr3bp_section2, bit for bit."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _golden_section(g, pre=""):
    from hiten_b200 import synodic
    axis = int(np.nonzero(g[pre + "req_normal"])[0][0])
    return synodic.make_section(axis, float(g[pre + "req_offset"]), tuple(str(c) for c in g[pre + "req_plane_coords"]),
                                int(g[pre + "req_direction"]), int(g[pre + "req_segment_refine"]),
                                float(g[pre + "req_tol_on_surface"]), float(g[pre + "req_dedup_time_tol"]),
                                float(g[pre + "req_dedup_point_tol"]))


@pytest.mark.parametrize("name,pre", [("synodic_c1.npz", ""), ("synodic_c2.npz", ""), ("synodic_se.npz", ""),
                                      ("c5_connection.npz", "l1_"), ("c5_connection.npz", "l2_")])
def test_section3_bit_exact_vs_reference(name, pre):
    import torch
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", name))
    mu, tf, steps, fwd = float(g["mu"]), float(g[pre + "tf"]), int(g[pre + "steps"]), int(g[pre + "forward"])
    x0 = g[pre + "x0W"]
    run = synodic.TubeSectionRunner(len(x0), mu, np.linspace(0.0, tf, steps), _golden_section(g, pre), forward=fwd,
                                    flip=(0, 6), pool_records=8)
    run.launch(torch.from_numpy(np.ascontiguousarray(x0.T)).cuda())
    h = run.sorted_hits()
    assert (run.status == 0).all().item()
    assert np.array_equal(run.yf.t().cpu().numpy(), g[pre + "yf"])
    kept = g[pre + "kept"] if (pre + "kept") in g.files else np.ones(len(x0), dtype=bool)
    sel = kept[h.trajectory_indices]                     # the reference's SynodicMap only sees the kept tubes
    assert np.array_equal((np.cumsum(kept) - 1)[h.trajectory_indices[sel]], g[pre + "hit_traj"])
    assert np.array_equal(h.times[sel], g[pre + "hit_time"]) and np.array_equal(h.states[sel], g[pre + "hit_state"])


@pytest.mark.parametrize("axis,offset,plane,direction", [
    ("x", 0.95, ("y", "vy"), 0), ("y", 0.0, ("x", "z"), 1), ("z", 0.0, ("x", "y"), 0),
    ("vx", 0.0, ("x", "y"), 0), ("vy", 0.0, ("x", "z"), -1), ("vz", 0.0, ("x", "y"), 0)])
@pytest.mark.parametrize("arith", ["parity", "fast"])
def test_section3_every_component_equals_section2(axis, offset, plane, direction, arith):
    import torch
    import hiten_b200 as hb
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", "synodic_c2.npz"))
    mu, tf, steps, fwd = float(g["mu"]), float(g["tf"]), int(g["steps"]), int(g["forward"])
    t_eval = np.linspace(0.0, tf, steps)
    sec = synodic.make_section(axis, offset, plane, direction)
    y0 = torch.from_numpy(np.ascontiguousarray(g["x0W"].T)).cuda()
    integ = hb.make_integ(arith=arith)
    a = synodic.TubeSectionRunner(200, mu, t_eval, sec, forward=fwd, flip=(0, 6), integ=integ, pool_records=8)
    b = synodic.TubeSectionRunner(200, mu, t_eval, sec, forward=fwd, flip=(0, 6), integ=integ, steps_capacity=192)
    a.launch(y0); b.launch(y0)
    ha, hb_ = a.sorted_hits(), b.sorted_hits()
    assert len(ha.times) > 0
    assert np.array_equal(ha.trajectory_indices, hb_.trajectory_indices)
    assert np.array_equal(ha.times, hb_.times) and np.array_equal(ha.states, hb_.states)
    assert torch.equal(a.yf, b.yf) and torch.equal(a.nacc, b.nacc) and torch.equal(a.nrej, b.nrej)


def test_section3_forward_time_and_small_pool_overflow_rerun():
    """Forward propagation (no sign flip), and a pool too small for the batch: trajectories that found it exhausted are
    flagged and rerun by the runner; the final hits are the same."""
    import torch
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", "c5_connection.npz"))
    mu, tf, steps = float(g["mu"]), float(g["l2_tf"]), int(g["l2_steps"])
    t_eval = np.linspace(0.0, tf, steps)
    sec = synodic.make_section("y", 0.0, ("x", "z"), 0)           # many crossings per trajectory
    x0 = np.tile(g["l2_x0W"], (64, 1))
    y0 = torch.from_numpy(np.ascontiguousarray(x0.T)).cuda()
    ref = synodic.TubeSectionRunner(len(x0), mu, t_eval, sec, forward=1, steps_capacity=192)
    ref.launch(y0)
    n_ref_over = int((ref.status == 4).sum().item())
    want = ref.sorted_hits()
    print(f"[section3] reference run: {len(want.times)} hits, {n_ref_over} record overflows, "
          f"max records {int(ref.records_written().max().item())}")
    big = synodic.TubeSectionRunner(len(x0), mu, t_eval, sec, forward=1, pool_records=16)
    big.launch(y0)
    got = big.sorted_hits()
    assert np.array_equal(got.trajectory_indices, want.trajectory_indices)
    assert np.array_equal(got.times, want.times) and np.array_equal(got.states, want.states)
    small = synodic.TubeSectionRunner(len(x0), mu, t_eval, sec, forward=1, pool_records=1)
    small.launch(y0)
    n_over = int((small.status == 4).sum().item())
    got = small.sorted_hits()
    assert n_over > 0 and (small.status == 0).all().item()
    assert np.array_equal(got.trajectory_indices, want.trajectory_indices)
    assert np.array_equal(got.times, want.times) and np.array_equal(got.states, want.states)
