"""GPU parity: dense tube on device -> section detection on device vs the reference's hits (C1, C2)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from test_oracle_synodic import oracle_hits

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _section(g):
    from hiten_b200 import synodic
    return synodic.make_section("y", float(g["req_offset"]), ("x", "z"), int(g["req_direction"]),
                                int(g["req_segment_refine"]), float(g["req_tol_on_surface"]),
                                float(g["req_dedup_time_tol"]), float(g["req_dedup_point_tol"]))


@pytest.mark.parametrize("name", ["c1", "c2", "se"])
def test_gpu_detector_bit_exact_on_identical_samples(name):
    """Same dense samples in -> identical hits out (integer/index work and IEEE arithmetic: bit-exact)."""
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", f"synodic_{name}.npz"))
    dense, times, hi, ht, hs = oracle_hits(g)
    got = synodic.detect(dense, times, _section(g))
    assert np.array_equal(got.trajectory_indices, g["hit_traj"])
    assert np.array_equal(got.times, g["hit_time"])
    assert np.array_equal(got.states, g["hit_state"])
    assert np.array_equal(got.points, g["hit_point"])


@pytest.mark.parametrize("name", ["c1", "c2", "se"])
def test_gpu_tube_and_section_vs_reference(name):
    """Full GPU chain (Manifold.compute -> SynodicMap.compute): identical crossing counts, crossing points
    within 1e-9 in synodic coordinates (BASELINE.json north_star)."""
    import hiten_b200 as hb
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", f"synodic_{name}.npz"))
    mu, tf, steps, fwd = float(g["mu"]), float(g["tf"]), int(g["steps"]), int(g["forward"])
    t_eval = np.linspace(0.0, tf, steps)
    res = hb.cr3bp_dense(g["x0W"], mu, t_eval, forward=fwd, flip=(0, 6), keep_on_device=True)
    got = synodic.detect(res.states, fwd * t_eval, _section(g))
    assert len(got.times) == len(g["hit_time"])
    assert np.array_equal(got.trajectory_indices, g["hit_traj"])
    dp = np.abs(got.points - g["hit_point"]).max(axis=1)
    dt = np.abs(got.times - g["hit_time"])
    ds = np.abs(got.states - g["hit_state"]).max(axis=1)
    print(f"[parity] {name}: {len(dp)} hits; |d point| median {np.median(dp):.2e} max {dp.max():.2e}; "
          f"|d t| max {dt.max():.2e}; |d state| max {ds.max():.2e}; >1e-9: {(dp > 1e-9).sum()}")
    assert dp.max() <= 1e-9
    assert np.array_equal(got.times, g["hit_time"]) and np.array_equal(got.states, g["hit_state"])   # bit-exact


def test_detector_ragged_and_empty():
    from hiten_b200 import synodic
    sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
    t = np.linspace(0.0, 1.0, 11)
    x = np.zeros((11, 6)); x[:, 1] = 0.5 - t
    lens = [11, 1, 7, 2]
    states = np.concatenate([x[:m] for m in lens])
    times = np.concatenate([t[:m] for m in lens])
    off = np.concatenate([[0], np.cumsum(lens)])
    got = synodic.detect(states, times, sec, offsets=off)
    assert got.hits_per_traj.tolist() == [1, 0, 1, 0]
    ref_t, ref_x = O.synodic_detect(t[:7], x[:7], 1, 0.0, -1)
    assert np.array_equal(got.times[got.trajectory_indices == 2], ref_t)
    empty = synodic.detect(np.empty((0, 5, 6)), t[:5], sec)
    assert len(empty.times) == 0


@pytest.mark.parametrize("name", ["c1", "c2", "se"])
def test_fused_section_equals_two_kernel_chain(name):
    """hb_cr3bp_section (no dense tube) == hb_cr3bp_dense + hb_synodic_detect, bit for bit, and both match the
    reference's crossing counts / points."""
    import hiten_b200 as hb
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", f"synodic_{name}.npz"))
    mu, tf, steps, fwd = float(g["mu"]), float(g["tf"]), int(g["steps"]), int(g["forward"])
    t_eval = np.linspace(0.0, tf, steps)
    sec = _section(g)
    dense = hb.cr3bp_dense(g["x0W"], mu, t_eval, forward=fwd, flip=(0, 6), keep_on_device=True)
    two = synodic.detect(dense.states, fwd * t_eval, sec)
    fused, res = synodic.tube_section(g["x0W"], mu, t_eval, sec, forward=fwd, flip=(0, 6))
    assert np.array_equal(fused.trajectory_indices, two.trajectory_indices)
    assert np.array_equal(fused.times, two.times) and np.array_equal(fused.states, two.states)
    assert np.array_equal(fused.hits_per_traj, two.hits_per_traj)
    assert np.array_equal(res.yf, dense.states[:, -1, :].cpu().numpy())
    assert len(fused.times) == len(g["hit_time"])
    assert np.abs(fused.points - g["hit_point"]).max() <= 1e-9


@pytest.mark.parametrize("name", ["c1", "c2", "se"])
@pytest.mark.parametrize("arith", ["parity", "fast"])
@pytest.mark.parametrize("records", ["all", "near"])
def test_two_kernel_section_equals_fused(name, arith, records):
    """hb_cr3bp_section2 (record + scan; every step recorded, or only the steps near the section plane) ==
    hb_cr3bp_section, bit for bit; parity also == the reference's hits."""
    import torch
    import hiten_b200 as hb
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", f"synodic_{name}.npz"))
    mu, tf, steps, fwd = float(g["mu"]), float(g["tf"]), int(g["steps"]), int(g["forward"])
    t_eval = np.linspace(0.0, tf, steps)
    sec = _section(g)
    y0 = torch.from_numpy(np.ascontiguousarray(g["x0W"].T)).cuda()
    integ = hb.make_integ(arith=arith)
    a = synodic.TubeSectionRunner(len(g["x0W"]), mu, t_eval, sec, forward=fwd, flip=(0, 6), integ=integ)
    b = synodic.TubeSectionRunner(len(g["x0W"]), mu, t_eval, sec, forward=fwd, flip=(0, 6), integ=integ, steps_capacity=256,
                                  records=records)
    a.launch(y0); b.launch(y0)
    ha, hb_ = a.sorted_hits(), b.sorted_hits()
    assert (b.status == 0).all().item()
    assert np.array_equal(ha.trajectory_indices, hb_.trajectory_indices)
    assert np.array_equal(ha.hits_per_traj, hb_.hits_per_traj)
    assert torch.equal(a.nacc, b.nacc)
    if arith == "parity":      # separately rounded arithmetic: identical in every kernel
        assert np.array_equal(ha.times, hb_.times) and np.array_equal(ha.states, hb_.states)
        assert torch.equal(a.yf, b.yf)
    else:                      # FMA contraction differs between kernels in the fast variant
        assert np.abs(ha.points - hb_.points).max() <= 1e-8 and np.abs(ha.times - hb_.times).max() <= 1e-8
    if arith == "parity":
        assert np.array_equal(hb_.times, g["hit_time"]) and np.array_equal(hb_.states, g["hit_state"])


def test_two_kernel_overflow_is_flagged():
    import torch
    import hiten_b200 as hb
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", "synodic_c1.npz"))
    t_eval = np.linspace(0.0, float(g["tf"]), int(g["steps"]))
    y0 = torch.from_numpy(np.ascontiguousarray(g["x0W"].T)).cuda()
    r = synodic.TubeSectionRunner(50, float(g["mu"]), t_eval, _section(g), forward=-1, flip=(0, 6), steps_capacity=80,
                                  records="all")
    r.launch(y0)
    st = r.status.cpu().numpy()
    na = r.nacc.cpu().numpy()
    assert ((st == 4) == (na > 96)).all() and (st == 4).any() and (st == 0).any()      # capacity rounds up to 96
    # ... and transparently rerun with the fused kernel: the final hits are the reference's, bit for bit
    h = r.sorted_hits()
    assert (r.status == 0).all().item()
    assert np.array_equal(h.times, g["hit_time"]) and np.array_equal(h.states, g["hit_state"])
    assert np.array_equal(h.trajectory_indices, g["hit_traj"])
    assert int(h.hits_per_traj.sum()) == len(g["hit_time"]) == r.hit_count()


@pytest.mark.parametrize("axis,offset,plane,direction", [
    ("x", 0.95, ("y", "vy"), 0), ("y", 0.0, ("x", "z"), 1), ("z", 0.0, ("x", "y"), 0),
    ("vx", 0.0, ("x", "y"), 0), ("vy", 0.0, ("x", "z"), -1), ("vz", 0.0, ("x", "y"), 0)])
@pytest.mark.parametrize("records", ["all", "near"])
def test_pipeline_every_section_component(axis, offset, plane, direction, records):
    """The scan kernel is instantiated per section component: each one against the fused kernel AND against the
    stored tube + detector chain (parity arithmetic: bit for bit), on the 200-trajectory tube of config 2."""
    import torch
    import hiten_b200 as hb
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", "synodic_c2.npz"))
    mu, tf, steps, fwd = float(g["mu"]), float(g["tf"]), int(g["steps"]), int(g["forward"])
    t_eval = np.linspace(0.0, tf, steps)
    sec = synodic.make_section(axis, offset, plane, direction)
    a, ra = synodic.tube_section(g["x0W"], mu, t_eval, sec, forward=fwd, flip=(0, 6), steps_capacity=0)
    b, rb = synodic.tube_section(g["x0W"], mu, t_eval, sec, forward=fwd, flip=(0, 6), steps_capacity=192, records=records)
    dense = hb.cr3bp_dense(g["x0W"], mu, t_eval, forward=fwd, flip=(0, 6), keep_on_device=True)
    c = synodic.detect(dense.states, fwd * t_eval, sec)
    assert len(c.times) > 0, "test section has no crossings"
    for h in (a, b):
        assert np.array_equal(h.trajectory_indices, c.trajectory_indices)
        assert np.array_equal(h.times, c.times) and np.array_equal(h.states, c.states)
        assert np.array_equal(h.hits_per_traj, c.hits_per_traj)
    assert np.array_equal(ra.yf, rb.yf) and (rb.status == 0).all()


def test_tube_section_auto_picks_pipeline_and_matches_reference():
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", "synodic_c2.npz"))
    t_eval = np.linspace(0.0, float(g["tf"]), int(g["steps"]))
    x0 = np.tile(g["x0W"], (2, 1))                      # 400 trajectories: above the auto threshold
    assert synodic._auto_steps_capacity(len(x0), "cuda") > 0
    h, res = synodic.tube_section(x0, float(g["mu"]), t_eval, _section(g), forward=int(g["forward"]), flip=(0, 6))
    k = len(g["hit_time"])
    assert len(h.times) == 2 * k and (res.status == 0).all()
    assert np.array_equal(h.times[:k], g["hit_time"]) and np.array_equal(h.states[:k], g["hit_state"])
    assert np.array_equal(h.times[k:], g["hit_time"]) and np.array_equal(h.trajectory_indices[k:], g["hit_traj"] + 200)


def test_stream_api_double_buffered_batches_match_reference():
    """TubeSectionStream: three host batches (two distinct) through the double-buffered pipeline; every result equals
    the reference's hits for that batch, bit for bit."""
    from hiten_b200 import synodic
    g = np.load(os.path.join(HERE, "golden", "synodic_c2.npz"))
    t_eval = np.linspace(0.0, float(g["tf"]), int(g["steps"]))
    x0 = np.ascontiguousarray(g["x0W"])
    rev = np.ascontiguousarray(x0[::-1])
    st = synodic.TubeSectionStream(len(x0), float(g["mu"]), t_eval, _section(g), forward=int(g["forward"]), flip=(0, 6),
                                   steps_capacity=192)
    outs = []
    for r in st.run([x0, rev, x0]):
        rec = r.hits[np.lexsort((r.hits["seq"], r.hits["traj"]))]
        outs.append((r.n_hits, rec["traj"].copy(), rec["t"].copy(), rec["state"].copy(), r.end_states.copy(), r.status.copy()))
    k = len(g["hit_time"])
    for i in (0, 2):
        n_hits, traj, t, state, yf, status = outs[i]
        assert n_hits == k and (status == 0).all()
        assert np.array_equal(traj, g["hit_traj"]) and np.array_equal(t, g["hit_time"]) and np.array_equal(state, g["hit_state"])
        assert np.array_equal(yf, g["yf"])
    n_hits, traj, t, state, yf, status = outs[1]                 # reversed batch: same hits under the index map
    assert n_hits == k and np.array_equal(yf, g["yf"][::-1])
    back = len(x0) - 1 - traj
    o = np.lexsort((t * 0, back))                                # stable within a trajectory
    assert np.array_equal(np.sort(back), np.sort(g["hit_traj"]))
    assert np.array_equal(np.sort(t), np.sort(g["hit_time"]))


def test_pipeline_equals_fused_on_a_bench_sized_slice_with_overflow_rerun():
    """20000 trajectories of the bench batch: the pipeline (with a scratch too small for the longest trajectories, so
    that the fused-kernel rerun path is exercised) returns exactly the fused kernel's hits, counts and end states."""
    import torch
    from hiten_b200 import workloads as W
    from hiten_b200 import synodic
    n = 20000
    ics, mu = W.c1_tube_batch(n)
    m = max(int(abs(W.C1_TF) / W.GRID_DT) + 1, 100)
    t_eval = np.linspace(0.0, W.C1_TF, m)
    sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
    y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
    a = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6))
    b = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=96, records="all")
    a.launch(y0); b.launch(y0)
    overflowed = int((b.status == 4).sum().item())
    ha, hb_ = a.sorted_hits(), b.sorted_hits()
    assert 0 < overflowed < n // 2                      # some, not most, trajectories need more than 96 steps
    assert (b.status == 0).all().item() and a.hit_count() == b.hit_count() > n
    assert np.array_equal(ha.trajectory_indices, hb_.trajectory_indices)
    assert np.array_equal(ha.times, hb_.times) and np.array_equal(ha.states, hb_.states)
    assert np.array_equal(ha.hits_per_traj, hb_.hits_per_traj)
    assert torch.equal(a.yf, b.yf) and torch.equal(a.nacc, b.nacc) and torch.equal(a.nrej, b.nrej)
    # the runner has sized its scratch for the longest trajectory: the next launch needs no rerun
    assert b.steps_capacity >= int(a.nacc.max().item()) > 96
    b.launch(y0)
    assert b.hit_count() == a.hit_count() and (b.status == 0).all().item() and b._extra == (None, None)


def test_full_size_1e6_trajectories_oracle_sample_and_order_invariance():
    """BASELINE configs[4] at full size (the bench batch: 1e6 trajectories through hb_cr3bp_section2). Size-independent
    checks: (a) a strided sample of 512 trajectories equals the oracle bit for bit (hit times, hit states, end states);
    (b) the reversed batch gives the same per-trajectory results under the index map — the persistent work queue and
    the per-trajectory candidate lists make the result independent of scheduling; (c) every hit lies on the section."""
    import torch
    from hiten_b200 import workloads as W
    from hiten_b200 import synodic
    n = 1_000_000
    ics, mu = W.c1_tube_batch(n)
    m = max(int(abs(W.C1_TF) / W.GRID_DT) + 1, 100)
    t_eval = np.linspace(0.0, W.C1_TF, m)
    sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
    run = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=160)   # the bench step
    y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
    run.launch(y0)
    h = run.sorted_hits()
    assert (run.status == 0).all().item()
    yf = run.yf.t().cpu().numpy().copy()
    nacc = run.nacc.cpu().numpy().copy()
    assert len(h.times) == run.hit_count() > n
    # (c) on the plane: a linear root between dense samples (exact up to rounding) or a left node within the
    # reference's on-surface tolerance 1e-6 (synodic/backend.py:555-572)
    assert np.abs(h.states[:, 1]).max() <= 1e-6 and np.median(np.abs(h.states[:, 1])) <= 1e-15
    # (a) oracle on a strided sample
    pick = np.arange(0, n, n // 512)[:512]
    s = O.system(O.SYS_CR3BP6, mu, fwd=-1, flip=(0, 6))
    dense, cnt = O.batch_dense(s, O.DOP853, O.default_tol(), ics[pick], t_eval, 8)
    assert np.array_equal(dense[:, -1, :], yf[pick])
    starts = np.concatenate(([0], np.cumsum(h.hits_per_traj)))
    for j, i in enumerate(pick):
        t, x = O.synodic_detect(-t_eval, dense[j], 1, 0.0, -1, (0, 2), 50, 1e-6, 1e-9, 1e-6)
        a, b = starts[i], starts[i + 1]
        assert b - a == len(t)
        assert np.array_equal(h.times[a:b], t) and np.array_equal(h.states[a:b], x)
        assert np.all(h.trajectory_indices[a:b] == i)
    # (b) reversed batch
    run.launch(torch.flip(y0, dims=[1]).contiguous())
    hr = run.sorted_hits()
    assert (run.status == 0).all().item()
    assert np.array_equal(run.yf.t().cpu().numpy(), yf[::-1])
    assert np.array_equal(run.nacc.cpu().numpy(), nacc[::-1])
    assert np.array_equal(hr.hits_per_traj, h.hits_per_traj[::-1])
    order = np.lexsort((np.arange(len(hr.times)), n - 1 - hr.trajectory_indices))     # stable within a trajectory
    assert np.array_equal(hr.times[order], h.times) and np.array_equal(hr.states[order], h.states)


def test_sparse_records_small_capacity_overflow_rerun_and_record_counts():
    """records="near": a tube needs a few dozen recorded steps per trajectory; a capacity of 32 overflows some
    trajectories (flagged, rerun: same final hits), and the recorded share of the accepted steps is small."""
    import torch
    from hiten_b200 import synodic
    from hiten_b200 import workloads as W
    n = 20000
    ics, mu = W.c1_tube_batch(n)
    m = max(int(abs(W.C1_TF) / W.GRID_DT) + 1, 100)
    t_eval = np.linspace(0.0, W.C1_TF, m)
    sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
    y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
    a = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=192, records="all")
    b = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=96, records="near")
    c = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=32, records="near")
    for r in (a, b, c):
        r.launch(y0)
    assert int((b.status == 4).sum().item()) == 0
    n_over = int((c.status == 4).sum().item())
    ha, hb_, hc = a.sorted_hits(), b.sorted_hits(), c.sorted_hits()
    print(f"[sparse] 20000 trajectories: capacity 32 overflows {n_over}")
    for h in (hb_, hc):
        assert np.array_equal(h.trajectory_indices, ha.trajectory_indices)
        assert np.array_equal(h.times, ha.times) and np.array_equal(h.states, ha.states)
        assert np.array_equal(h.hits_per_traj, ha.hits_per_traj)
    assert torch.equal(a.yf, b.yf) and torch.equal(a.yf, c.yf) and (c.status == 0).all().item()


def test_hit_buffer_grows_instead_of_raising_and_segment_refine_zero():
    """(a) a propagation with more than 8 crossings per trajectory overflows the default hit buffer: the runner sizes it
    for all hits and reruns (the reference has no cap); an explicit hit_capacity stays a hard limit.  (b) the
    segment_refine = 0 branch of the detector (one linear root per sample segment, backend.py:782-821) through the
    pipeline, the fused kernel and the stored-tube detector: identical hits, equal to the oracle's."""
    import torch
    import hiten_b200 as hb
    from hiten_b200 import synodic
    from hiten_b200._lib import HitenB200Error
    g = np.load(os.path.join(HERE, "golden", "synodic_c2.npz"))
    mu, fwd = float(g["mu"]), int(g["forward"])
    x0 = np.tile(g["x0W"], (2, 1))                                   # 400 trajectories, ~3.4 crossings each
    t_eval = np.linspace(0.0, float(g["tf"]), int(g["steps"]))
    y0 = torch.from_numpy(np.ascontiguousarray(x0.T)).cuda()
    run = synodic.TubeSectionRunner(len(x0), mu, t_eval, _section(g), forward=fwd, flip=(0, 6), steps_capacity=192)
    run.cap = 256                                                    # (the default is max(1024, 8 n): force the overflow)
    run.hits = torch.empty(run.cap * 9, dtype=torch.float64, device="cuda")
    run.launch(y0)
    h = run.sorted_hits()
    k = len(g["hit_time"])
    assert len(h.times) == 2 * k > 256 and run.cap >= 2 * k
    assert np.array_equal(h.times[:k], g["hit_time"]) and np.array_equal(h.times[k:], g["hit_time"])
    tight = synodic.TubeSectionRunner(len(x0), mu, t_eval, _section(g), forward=fwd, flip=(0, 6), steps_capacity=192,
                                      hit_capacity=64)
    tight.launch(y0)
    with pytest.raises(HitenB200Error):
        tight.hit_count()
    # (b)
    sec0 = synodic.make_section("y", 0.0, ("x", "z"), -1, segment_refine=0)
    a, _ = synodic.tube_section(g["x0W"], mu, t_eval, sec0, forward=fwd, flip=(0, 6), steps_capacity=0)
    b, _ = synodic.tube_section(g["x0W"], mu, t_eval, sec0, forward=fwd, flip=(0, 6), steps_capacity=192)
    dense = hb.cr3bp_dense(g["x0W"], mu, t_eval, forward=fwd, flip=(0, 6), keep_on_device=True)
    c = synodic.detect(dense.states, fwd * t_eval, sec0)
    assert len(c.times) > 100
    for hh in (a, b):
        assert np.array_equal(hh.trajectory_indices, c.trajectory_indices)
        assert np.array_equal(hh.times, c.times) and np.array_equal(hh.states, c.states)
    tube = dense.states.cpu().numpy()
    for i in (0, 57, 199):
        t, x = O.synodic_detect(fwd * t_eval, tube[i], 1, 0.0, -1, (0, 2), 0, 1e-6, 1e-9, 1e-6)
        sel = c.trajectory_indices == i
        assert np.array_equal(c.times[sel], t) and np.array_equal(c.states[sel], x)
