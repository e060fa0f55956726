"""GPU parity of the manifold-tube host steps: hb_manifold_ics / hb_tube_filter vs the reference's
own outputs (tests/golden/manifold_ics.npz) and vs the oracle at bench sizes."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden", "manifold_ics.npz")


def full_phi(g, tag):
    phi = np.zeros((g[f"{tag}_tt"].size, 42))
    phi[g[f"{tag}_rows"]] = g[f"{tag}_phi_rows"]
    return phi


@pytest.mark.parametrize("tag", ["sp", "un"])
def test_initial_conditions_bit_exact_vs_reference(tag):
    from hiten_b200 import manifold
    g = np.load(G)
    x0, idx = manifold.tube_initial_conditions(full_phi(g, tag), g[f"{tag}_tt"], float(g["period"]),
                                               g[f"{tag}_eigvec"].astype(complex), int(g[f"{tag}_direction"]),
                                               g["fractions"], g["displacements"])
    assert np.array_equal(idx.cpu().numpy(), g[f"{tag}_idx"])
    assert np.array_equal(x0.t().cpu().numpy(), g[f"{tag}_x0W"])


def test_initial_conditions_large_tube_vs_oracle():
    """2000 nodes x 500 displacements (the config-5 tube) from a dense random STM: every IC equals the oracle's."""
    import oracle_lib as O
    from hiten_b200 import manifold
    rng = np.random.default_rng(3)
    S = 2000
    tt = -np.linspace(0.0, 2.75, S)
    phi = rng.standard_normal((S, 42))
    phi[5, :18] = 0.0                                       # |MAN[0:3]| = 0 -> magnitude 1.0 branch
    phi[:, 38] = rng.choice([0.0, 3e-16, 0.1], S)           # tiny z gets zeroed
    ev = rng.standard_normal(6)
    fr = np.arange(0.0, 1.0, 0.0005)
    dd = np.logspace(-7, -5, 500)
    x0, idx = manifold.tube_initial_conditions(phi, tt, 2.75, ev, -1, fr, dd)
    ref, ridx = O.manifold_ics(phi, tt, 2.75, ev, -1, fr, dd)
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(x0.t().cpu().numpy(), ref)


def test_filter_synthetic_bit_exact_vs_reference():
    from hiten_b200 import manifold
    g = np.load(G)
    out, keep = manifold.tube_filter(g["syn_states"], float(g["mu"]), safe_r1=3.318e-05, safe_r2=9.04e-06,
                                     energy_tol=1e-6)
    ref = g["syn_filter"]
    assert np.array_equal(out.cpu().numpy(), ref, equal_nan=True)
    want = ~((ref[:, 0] < 3.318e-05) | (ref[:, 1] < 9.04e-06)) & ~(ref[:, 2] > 1e-6)
    assert np.array_equal(keep.cpu().numpy().astype(bool), want)


def test_filter_on_device_tubes_bit_exact_vs_reference():
    """The 50 default tubes: propagated on the GPU, filtered where they lie, equal to the reference's quantities."""
    import hiten_b200 as hb
    from hiten_b200 import manifold
    g = np.load(G)
    mu = float(g["mu"])
    t_eval = np.linspace(0.0, float(g["c1_tf"]), int(g["c1_steps"]))
    res = hb.cr3bp_dense(g["c1_x0W"], mu, t_eval, forward=-1, flip=(0, 6), keep_on_device=True)
    out, keep = manifold.tube_filter(res.states, mu, safe_r1=3.318e-05, safe_r2=9.04e-06, energy_tol=1e-7)
    assert np.array_equal(out.cpu().numpy(), g["c1_filter"])
    assert np.array_equal(keep.cpu().numpy().astype(bool), ~(g["c1_filter"][:, 2] > 1e-7))
    assert 0 < int(keep.sum()) < 50                          # the threshold splits the batch


def test_filter_many_tubes_vs_oracle():
    import torch
    import oracle_lib as O
    from hiten_b200 import manifold
    rng = np.random.default_rng(11)
    s = rng.uniform(-1.5, 1.5, size=(3000, 257, 6))
    want = O.tube_filter(s, 0.0121505856)
    out, _ = manifold.tube_filter(torch.from_numpy(s).cuda(), 0.0121505856)
    assert np.array_equal(out.cpu().numpy(), want)
    # an array that is only 8-byte aligned takes the scalar-load path: same numbers
    flat = torch.empty(s.size + 1, dtype=torch.float64, device="cuda")
    view = flat[1:].view(s.shape)
    view.copy_(torch.from_numpy(s))
    assert view.data_ptr() % 16 == 8
    out, _ = manifold.tube_filter(view, 0.0121505856)
    assert np.array_equal(out.cpu().numpy(), want)


def _c1_section_runner(g, steps_capacity, energy_tol):
    import torch
    from hiten_b200 import synodic
    mu = float(g["mu"])
    t_eval = np.linspace(0.0, float(g["c1_tf"]), int(g["c1_steps"]))
    sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
    r = synodic.TubeSectionRunner(len(g["c1_x0W"]), mu, t_eval, sec, forward=-1, flip=(0, 6),
                                  steps_capacity=steps_capacity, filters=(3.318e-05, 9.04e-06, energy_tol))
    r.launch(torch.from_numpy(np.ascontiguousarray(g["c1_x0W"].T)).cuda())
    return r


@pytest.mark.parametrize("steps_capacity", [160, 96])
def test_record_filter_bit_exact_vs_reference(steps_capacity):
    """hb_section2_filter: the filter quantities of the 50 default tubes rebuilt from the step records (capacity 96:
    part of the batch overflows the records and is judged on stored tubes instead) equal the reference's own."""
    g = np.load(G)
    r = _c1_section_runner(g, steps_capacity, 1e-7)
    out, keep = r.filter_result()
    assert np.array_equal(out.cpu().numpy(), g["c1_filter"])
    want = ~(g["c1_filter"][:, 2] > 1e-7)
    assert np.array_equal(keep.cpu().numpy() == 1, want) and 0 < want.sum() < 50
    # hits of discarded trajectories are gone, the others are the unfiltered run's
    from hiten_b200 import synodic
    h = r.sorted_hits()
    plain = np.load(os.path.join(os.path.dirname(__file__), "golden", "synodic_c1.npz"))
    assert np.array_equal(plain["x0W"], g["c1_x0W"])
    sel = want[plain["hit_traj"]]
    assert np.array_equal(h.trajectory_indices, plain["hit_traj"][sel])
    assert np.array_equal(h.times, plain["hit_time"][sel]) and np.array_equal(h.states, plain["hit_state"][sel])
    assert (h.hits_per_traj[~want] == 0).all()


def test_record_filter_equals_stored_tube_filter_on_a_bench_slice():
    """4000 trajectories of the bench batch, both arithmetic variants: records -> filter == dense tube -> filter."""
    import torch
    from hiten_b200 import workloads as W
    import hiten_b200 as hb
    from hiten_b200 import manifold, synodic
    n = 4000
    ics, mu = W.c1_tube_batch(n)
    m = max(int(abs(W.C1_TF) / W.GRID_DT) + 1, 100)
    t_eval = np.linspace(0.0, W.C1_TF, m)
    sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
    y0 = torch.from_numpy(np.ascontiguousarray(ics.T)).cuda()
    for arith in ("parity", "fast"):
        integ = hb.make_integ(arith=arith)
        r = synodic.TubeSectionRunner(n, mu, t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=192, integ=integ,
                                      filters=(1e-3, 1e-3, 1e-9))
        r.launch(y0)
        out, keep = r.filter_result()
        tube = hb.cr3bp_dense(y0, mu, t_eval, forward=-1, flip=(0, 6), integ=integ, keep_on_device=True)
        ref, rkeep = manifold.tube_filter(tube.states, mu, safe_r1=1e-3, safe_r2=1e-3, energy_tol=1e-9)
        if arith == "parity":
            assert torch.equal(out, ref) and torch.equal(keep, rkeep)
        else:
            assert torch.allclose(out, ref, rtol=1e-6, atol=1e-12)
        assert 0 < int(keep.sum()) < n
        del tube
