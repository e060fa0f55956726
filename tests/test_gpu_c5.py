"""GPU parity on BASELINE configs[4] as the reference runs it (examples/heteroclinic_connection.py:29-63):
L1 halo Az = 0.5 S stable tube (integration_fraction 0.9) and L2 halo Az = 0.3663368 N unstable tube (1.0), step 0.005,
section x = 1 - mu / (y, z) / direction -1 with the stable-manifold flip of connections/interfaces.py:350, trajectory
filters of Manifold.compute(), then the connection search -- against tests/golden/c5_connection.npz, the reference's own
numbers for this flow (tests/golden/make_c5.py).  Parity arithmetic: bit for bit."""
import os

import numpy as np
import pytest

from test_oracle_c5 import ENERGY_TOL, IDX, c5, reference_safe_radii

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def section_of(g, key):
    from hiten_b200 import synodic
    axis = int(np.nonzero(g[f"{key}_req_normal"])[0][0])
    return synodic.make_section(axis, float(g[f"{key}_req_offset"]), tuple(str(c) for c in g[f"{key}_req_plane_coords"]),
                                int(g[f"{key}_req_direction"]), int(g[f"{key}_req_segment_refine"]),
                                float(g[f"{key}_req_tol_on_surface"]), float(g[f"{key}_req_dedup_time_tol"]),
                                float(g[f"{key}_req_dedup_point_tol"]))


def check_hits(g, key, h, keep):
    """h: SectionHits over ALL trajectories of the tube (discarded ones contribute nothing); the reference numbers its
    hits by position among the KEPT trajectories (SynodicMap sees manifold.result's states_list)."""
    rank = np.cumsum(keep) - 1
    assert keep[h.trajectory_indices].all()
    assert np.array_equal(rank[h.trajectory_indices], g[f"{key}_hit_traj"])
    assert np.array_equal(h.times, g[f"{key}_hit_time"])
    assert np.array_equal(h.states, g[f"{key}_hit_state"])
    assert np.array_equal(h.points, g[f"{key}_hit_point"])


@pytest.mark.parametrize("key", ["l1", "l2"])
@pytest.mark.parametrize("steps_capacity", [192, 96])
def test_c5_tube_section_runner_with_filters_bit_exact(key, steps_capacity):
    """TubeSectionRunner(filters=...): the section pipeline (hb_cr3bp_section2 + hb_section2_filter) on the two C5
    tubes.  Capacity 96 forces the overflow path (fused-kernel rerun + stored-tube filter) on the long arcs."""
    import torch
    from hiten_b200 import synodic
    g = c5()
    mu, tf, steps, fwd = float(g["mu"]), float(g[f"{key}_tf"]), int(g[f"{key}_steps"]), int(g[f"{key}_forward"])
    t_eval = np.linspace(0.0, tf, steps)
    r1, r2 = reference_safe_radii()
    run = synodic.TubeSectionRunner(200, mu, t_eval, section_of(g, key), forward=fwd, flip=(0, 6),
                                    steps_capacity=steps_capacity, filters=(r1, r2, ENERGY_TOL))
    run.launch(torch.from_numpy(np.ascontiguousarray(g[f"{key}_x0W"].T)).cuda())
    h = run.sorted_hits()
    q, keep = run.filter_result()
    keep = keep.cpu().numpy() == 1
    assert (run.status == 0).all().item()
    assert np.array_equal(run.yf.t().cpu().numpy(), g[f"{key}_yf"])
    assert np.array_equal(keep, g[f"{key}_kept"])
    check_hits(g, key, h, keep)


@pytest.mark.parametrize("key", ["l1", "l2"])
def test_c5_stored_tube_chain_and_fused_kernel_bit_exact(key):
    """The two other forms of the same step: dense tube + tube filter + detector on the stored tube (what the drop-in's
    _run_compute / _SynodicDetectionBackend.run do), and the fused kernel hb_cr3bp_section."""
    import hiten_b200 as hb
    from hiten_b200 import manifold, synodic
    g = c5()
    mu, tf, steps, fwd = float(g["mu"]), float(g[f"{key}_tf"]), int(g[f"{key}_steps"]), int(g[f"{key}_forward"])
    t_eval = np.linspace(0.0, tf, steps)
    r1, r2 = reference_safe_radii()
    sec = section_of(g, key)
    tube = hb.cr3bp_dense(g[f"{key}_x0W"], mu, t_eval, forward=fwd, flip=(0, 6), keep_on_device=True)
    assert np.array_equal(tube.states[:, -1, :].cpu().numpy(), g[f"{key}_yf"])
    _, keep = manifold.tube_filter(tube.states, mu, safe_r1=r1, safe_r2=r2, energy_tol=ENERGY_TOL)
    keep = keep.cpu().numpy() == 1
    assert np.array_equal(keep, g[f"{key}_kept"])
    sel = np.nonzero(keep)[0]
    got = synodic.detect(tube.states[sel], fwd * t_eval, sec)
    assert np.array_equal(got.trajectory_indices, g[f"{key}_hit_traj"])
    assert np.array_equal(got.times, g[f"{key}_hit_time"]) and np.array_equal(got.states, g[f"{key}_hit_state"])
    fused, res = synodic.tube_section(g[f"{key}_x0W"][sel], mu, t_eval, sec, forward=fwd, flip=(0, 6), steps_capacity=0)
    assert np.array_equal(fused.trajectory_indices, g[f"{key}_hit_traj"])
    assert np.array_equal(fused.times, g[f"{key}_hit_time"]) and np.array_equal(fused.states, g[f"{key}_hit_state"])
    assert np.array_equal(res.yf, g[f"{key}_yf"][sel])


@pytest.mark.parametrize("key", ["l1", "l2"])
@pytest.mark.parametrize("records", ["near", "all"])
def test_c5_section2_without_filters_equals_stored_tube_chain(key, records):
    """The bench step's form (no trajectory filters; sparse or full step records) on the 200 golden trajectories of each
    tube against the stored-tube chain: every trajectory's hits, discarded ones included."""
    import torch
    import hiten_b200 as hb
    from hiten_b200 import synodic
    g = c5()
    mu, tf, steps, fwd = float(g["mu"]), float(g[f"{key}_tf"]), int(g[f"{key}_steps"]), int(g[f"{key}_forward"])
    t_eval = np.linspace(0.0, tf, steps)
    sec = section_of(g, key)
    run = synodic.TubeSectionRunner(200, mu, t_eval, sec, forward=fwd, flip=(0, 6), steps_capacity=192, records=records)
    run.launch(torch.from_numpy(np.ascontiguousarray(g[f"{key}_x0W"].T)).cuda())
    h = run.sorted_hits()
    tube = hb.cr3bp_dense(g[f"{key}_x0W"], mu, t_eval, forward=fwd, flip=(0, 6), keep_on_device=True)
    want = synodic.detect(tube.states, fwd * t_eval, sec)
    assert (run.status == 0).all().item() and len(want.times) >= len(g[f"{key}_hit_time"])
    assert np.array_equal(h.trajectory_indices, want.trajectory_indices)
    assert np.array_equal(h.times, want.times) and np.array_equal(h.states, want.states)
    assert np.array_equal(run.yf.t().cpu().numpy(), g[f"{key}_yf"])


def test_c5_connections_bit_exact_vs_reference():
    from hiten_b200 import connections as cn
    g = c5()
    r = cn.find_connections(g["conn_pu"], g["conn_ps"], g["conn_Xu"], g["conn_Xs"], float(g["conn_eps"]),
                            float(g["conn_dv_tol"]), float(g["conn_bal_tol"]), traj_indices_u=g["conn_tu"],
                            traj_indices_s=g["conn_ts"])
    assert r.pairs_considered == int(g["conn_pairs_considered"]) and len(r.delta_v) == 6
    assert np.array_equal(r.index_u, g["conn_iu"]) and np.array_equal(r.index_s, g["conn_is"])
    assert np.array_equal(r.kind, g["conn_kind"]) and np.array_equal(r.delta_v, g["conn_dv"])
    assert np.array_equal(r.point2d, g["conn_pt"])
    assert np.array_equal(r.state_u, g["conn_su"]) and np.array_equal(r.state_s, g["conn_ss"])
    assert np.array_equal(r.trajectory_index_u, g["conn_tiu"]) and np.array_equal(r.trajectory_index_s, g["conn_tis"])


def test_c5_whole_flow_on_the_device_gives_the_reference_connections():
    """Tubes -> filters -> section hits -> connection search without leaving the device API: the 6 connections of the
    reference's example, bit for bit (Delta-v, meeting point, both states, trajectory indices)."""
    import torch
    from hiten_b200 import connections as cn
    from hiten_b200 import synodic
    g = c5()
    mu = float(g["mu"])
    r1, r2 = reference_safe_radii()
    hits = {}
    for key in ("l1", "l2"):
        tf, steps, fwd = float(g[f"{key}_tf"]), int(g[f"{key}_steps"]), int(g[f"{key}_forward"])
        run = synodic.TubeSectionRunner(200, mu, np.linspace(0.0, tf, steps), section_of(g, key), forward=fwd,
                                        flip=(0, 6), steps_capacity=192, filters=(r1, r2, ENERGY_TOL))
        run.launch(torch.from_numpy(np.ascontiguousarray(g[f"{key}_x0W"].T)).cuda())
        h = run.sorted_hits()
        keep = run.filter_result()[1].cpu().numpy() == 1
        hits[key] = (h, (np.cumsum(keep) - 1)[h.trajectory_indices])
    (hu, tu), (hs, ts) = hits["l1"], hits["l2"]
    r = cn.find_connections(hu.points, hs.points, hu.states, hs.states, float(g["conn_eps"]), float(g["conn_dv_tol"]),
                            float(g["conn_bal_tol"]), traj_indices_u=tu, traj_indices_s=ts)
    # the reference's request lists the same hits in worker-completion order: compare through the hits themselves
    assert len(r.delta_v) == 6
    assert np.array_equal(r.delta_v, g["conn_dv"]) and np.array_equal(r.kind, g["conn_kind"])
    assert np.array_equal(r.point2d, g["conn_pt"])
    assert np.array_equal(r.state_u, g["conn_su"]) and np.array_equal(r.state_s, g["conn_ss"])
    assert np.array_equal(r.trajectory_index_u, g["conn_tiu"]) and np.array_equal(r.trajectory_index_s, g["conn_tis"])


def test_c5_two_tubes_side_by_side_with_capped_grids_same_bits():
    """hb_integ.max_ctas: the two tubes' pipelines on two streams, each persistent launch capped at half the SMs (how
    bench.py runs the strong-scaling shards), against one uncapped tube after the other: hits, end states and step
    counts bit for bit on 20000 trajectories per tube."""
    import torch
    import hiten_b200 as hb
    from hiten_b200 import synodic
    from hiten_b200 import workloads as W
    ics, mu = W.c5_batch(40000)
    dev = torch.device("cuda", 0)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    res = {}
    for cap in (0, max(sms // 2, 1)):
        integ = hb.make_integ(max_ctas=cap)
        runs, streams = {}, {}
        for key in ("l1", "l2"):
            runs[key] = synodic.TubeSectionRunner(len(ics[key]), mu, W.c5_grid(key), W.c5_section(key, mu),
                                                  forward=W.C5_TUBES[key]["forward"], flip=(0, 6), integ=integ, device=dev,
                                                  steps_capacity=128, records="near")
            streams[key] = torch.cuda.Stream(dev) if cap else torch.cuda.current_stream()
        y0 = {key: torch.from_numpy(np.ascontiguousarray(ics[key].T)).to(dev) for key in runs}
        torch.cuda.synchronize()
        for key in runs:
            runs[key].launch(y0[key], streams[key])
        torch.cuda.synchronize()
        res[cap] = {key: (runs[key].sorted_hits(), runs[key].yf.cpu().numpy(), runs[key].nacc.cpu().numpy(),
                          runs[key].nrej.cpu().numpy()) for key in runs}
    a, b = res[0], res[max(sms // 2, 1)]
    for key in a:
        assert len(a[key][0].times) > 1000
        assert np.array_equal(a[key][0].trajectory_indices, b[key][0].trajectory_indices)
        assert np.array_equal(a[key][0].times, b[key][0].times) and np.array_equal(a[key][0].states, b[key][0].states)
        for i in (1, 2, 3):
            assert np.array_equal(a[key][i], b[key][i])


def test_c5_launch_order_is_a_scheduling_hint_only():
    """hb_integ.order: the persistent propagation launch hands out the trajectories longest-first (the step counts of a
    first launch as the cost) or in a random order -- hits, end states, step counts and record counts are those of the
    natural order bit for bit (outputs stay in the caller's indexing), for the pipeline and for hb_cr3bp_propagate."""
    import torch
    import hiten_b200 as hb
    from hiten_b200 import synodic
    from hiten_b200 import workloads as W
    ics, mu = W.c5_batch(40000)
    dev = torch.device("cuda", 0)
    for key in ("l1", "l2"):
        run = synodic.TubeSectionRunner(len(ics[key]), mu, W.c5_grid(key), W.c5_section(key, mu),
                                        forward=W.C5_TUBES[key]["forward"], flip=(0, 6), device=dev, steps_capacity=128,
                                        records="near")
        y0 = torch.from_numpy(np.ascontiguousarray(ics[key].T)).to(dev)

        def snapshot():
            run.launch(y0)
            h = run.sorted_hits()
            return (h.trajectory_indices, h.times, h.states, run.yf.cpu().numpy(), run.nacc.cpu().numpy(),
                    run.nrej.cpu().numpy(), run.status.cpu().numpy(), run.records_written().cpu().numpy())
        a = snapshot()
        assert len(a[1]) > 1000 and a[4].max() > 1.5 * a[4].min()
        run.order_by_cost()                                   # longest first, from the launch above
        order = run._order.cpu().numpy()
        assert np.array_equal(np.sort(order), np.arange(run.n))
        cost = a[4] + a[5]
        assert np.all(np.diff(cost[order]) <= 0)
        b = snapshot()
        g = torch.Generator().manual_seed(5)
        run.set_order(torch.randperm(run.n, generator=g).to(torch.int32).to(dev))
        c = snapshot()
        run.set_order(None)
        d = snapshot()
        for other in (b, c, d):
            for u, v in zip(a, other):
                assert np.array_equal(u, v)
        # the plain propagation entry point honours the same hint
        tf = float(W.c5_grid(key)[-1])
        kw = dict(forward=W.C5_TUBES[key]["forward"], flip=(0, 6))
        r0 = hb.cr3bp_propagate(y0, mu, tf, **kw)
        ordered = hb.with_order(hb.make_integ(), hb.cost_order(r0.n_acc + r0.n_rej))
        r1 = hb.cr3bp_propagate(y0, mu, tf, integ=ordered, **kw)
        assert torch.equal(r0.yf, r1.yf) and torch.equal(r0.n_acc, r1.n_acc) and torch.equal(r0.n_rej, r1.n_rej)
        with pytest.raises(ValueError):
            run.set_order(torch.zeros(3, dtype=torch.int32, device=dev))


def test_tube_section_in_chunks_equals_one_launch():
    """tube_section(steps_capacity="auto") on a batch whose step scratch does not fit the free memory (forced here with
    the free-memory override): three chunks through one scratch, the last one shorter -- same hits, counts and end states
    as the single launch, trajectory indices of the whole batch."""
    import hiten_b200 as hb
    from hiten_b200 import synodic
    from hiten_b200 import workloads as W
    ics, mu = W.c5_batch(6000)
    x = ics["l2"]
    assert len(x) == 3000
    kw = dict(forward=W.C5_TUBES["l2"]["forward"], flip=(0, 6))
    chunk, cap = synodic._auto_plan(3000, "cuda:0", "near", free_bytes=160e6)
    assert cap == 128 and 1024 <= chunk < 1600
    h1, r1 = synodic.tube_section(x, mu, W.c5_grid("l2"), W.c5_section("l2", mu), **kw)
    h3, r3 = synodic.tube_section(x, mu, W.c5_grid("l2"), W.c5_section("l2", mu), _free_bytes=160e6, **kw)
    assert len(h1.times) > 100
    assert np.array_equal(h1.trajectory_indices, h3.trajectory_indices)
    assert np.array_equal(h1.times, h3.times) and np.array_equal(h1.states, h3.states)
    assert np.array_equal(h1.points, h3.points) and np.array_equal(h1.hits_per_traj, h3.hits_per_traj)
    assert np.array_equal(r1.yf, r3.yf) and np.array_equal(r1.n_acc, r3.n_acc) and np.array_equal(r1.status, r3.status)
