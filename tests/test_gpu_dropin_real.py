"""GPU: hiten_b200.install() under the REAL reference with the REAL CUDA library (no stand-ins).

The reference package is imported from oracle/_ref (oracle/build_ref.sh: unmodified copy + import stubs, shipped with the
snapshot) or, in the build container, from the read-only checkout.  Its own user-level calls -- orbit.manifold().compute(),
SynodicMap.compute(), cm.poincare_map().compute(), ConnectionPipeline.solve() -- then run with the funnels rebound to
libhiten_b200.so, and the results are compared with the golden vectors the unpatched reference produced
(tests/golden/make_*.py).

Two kinds of check per flow:
  * seam, bit for bit: the generating orbit is corrected and its STM / eigen-data computed by the unpatched reference
    first (cached in its services), then install(): the tube, the trajectory filters, the section hits and the
    connections that come out of the rebound calls equal the golden vectors exactly;
  * everything installed from the first call (the orbit's own correction and STM run on the GPU too, which agree with
    the reference to ~1e-12 / 2e-11, amplified along the manifold): identical crossing counts, points within 1e-6.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import _refenv  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(_refenv.available() is None, reason="no reference package (run oracle/build_ref.sh)")]


def G(name):
    return np.load(os.path.join(HERE, "golden", name))


@pytest.fixture(scope="module")
def ref():
    _refenv.enable()
    import hiten_b200
    assert not hiten_b200.dropin.is_installed()
    from hiten import System
    system = System.from_bodies("earth", "moon")
    yield system
    hiten_b200.uninstall()


def _prime(manifold):
    """Let the unpatched reference compute (and cache) everything _run_compute reads besides the tube itself."""
    svc = manifold.dynamics
    svc.compute_stm(steps=2000)
    svc.eigenvalues, svc.eigenvectors
    return svc


def _library_loaded():
    with open("/proc/self/maps") as f:
        return any("libhiten_b200.so" in line for line in f)


def test_c1_manifold_and_synodic_map_seam_bit_exact(ref):
    """configs[0] + its section: unpatched orbit/STM, then the rebound manifold.compute() and
    SynodicMap.compute() on the real kernels == the reference's 50 tubes' ends and 121 hits, bit for bit."""
    import hiten_b200
    from hiten import SynodicMap
    g = G("synodic_c1.npz")
    l1 = ref.get_libration_point(1)
    halo = l1.create_orbit("halo", amplitude_z=0.2, zenith="southern")
    halo.correct()
    assert np.array_equal(np.asarray(halo.initial_state, float), g["orbit_x0"])
    manifold = halo.manifold(stable=True, direction="positive")
    _prime(manifold)
    hiten_b200.install()
    try:
        ysos, dysos, states_list, times_list, successes, attempts = manifold.compute(show_progress=False)
        assert (successes, attempts) == (50, 50) and states_list[0].shape == (int(g["steps"]), 6)
        assert np.array_equal(np.stack([s[0] for s in states_list]), g["x0W"])
        assert np.array_equal(np.stack([s[-1] for s in states_list]), g["yf"])
        assert times_list[0][-1] == -float(g["tf"])
        smap = SynodicMap(manifold)
        smap.compute(section_axis="y", section_offset=0.0, plane_coords=("x", "z"), direction=-1)
        sec = smap.dynamics.get_section()
        pts, sts = np.asarray(sec.points, float), np.asarray(sec.states, float)
        assert pts.shape == (121, 2)
        # worker chunks complete in any order (synodic/engine.py:137): compare as sets of rows
        a, b = np.lexsort(sts.T[::-1]), np.lexsort(g["hit_state"].T[::-1])
        assert np.array_equal(sts[a], g["hit_state"][b]) and np.array_equal(pts[a], g["hit_point"][b])
    finally:
        hiten_b200.uninstall()
    assert _library_loaded()


def test_c1_everything_installed_counts_and_points(ref):
    """The same flow with the drop-in active from the first call: halo.correct() (the reference's Newton loop over GPU
    event / STM propagations), the orbit's STM, the tube and the section all run on the device."""
    import hiten_b200
    from hiten import SynodicMap
    g = G("synodic_c1.npz")
    hiten_b200.install()
    try:
        l1 = ref.get_libration_point(1)
        halo = l1.create_orbit("halo", amplitude_z=0.2, zenith="southern")
        halo.correct()
        assert np.abs(np.asarray(halo.initial_state, float) - g["orbit_x0"]).max() <= 1e-10
        assert abs(halo.period - float(g["orbit_period"])) <= 1e-10
        manifold = halo.manifold(stable=True, direction="positive")
        _, _, states_list, times_list, successes, attempts = manifold.compute(show_progress=False)
        assert (successes, attempts) == (50, 50)
        # initial conditions: ~1e-11 from the reference's, except where fraction * T falls exactly midway between two
        # of the 2000 STM samples (fraction 0.5): `_totime`'s argmin then picks either neighbour depending on the last
        # bits of the period -- a one-sample shift along the orbit (2.7e-4) that the reference itself shows under a
        # 1-ulp change of T
        dev = np.abs(np.stack([s[0] for s in states_list]) - g["x0W"]).max(axis=1)
        shifted = dev > 1e-9
        print(f"[dropin] C1 fully installed: IC deviation median {np.median(dev):.1e}, one-sample ties {int(shifted.sum())}")
        assert shifted.sum() <= 2 and dev[~shifted].max() <= 1e-10 and dev.max() <= 5e-4
        smap = SynodicMap(manifold)
        smap.compute(section_axis="y", section_offset=0.0, plane_coords=("x", "z"), direction=-1)
        pts = np.asarray(smap.get_points())
        assert pts.shape == (121, 2)                                  # identical crossing count
        d = np.abs(np.sort(pts[:, 0]) - np.sort(g["hit_point"][:, 0]))
        print(f"[dropin] C1 fully installed: 121 hits, |d x| of sorted crossing points median {np.median(d):.2e} max {d.max():.2e}")
        assert np.median(d) <= 1e-6 and (d > 1e-6).sum() <= 3 * max(int(shifted.sum()), 1)
    finally:
        hiten_b200.uninstall()


@pytest.fixture(scope="module")
def cm(ref):
    l1 = ref.get_libration_point(1)
    c = l1.get_center_manifold(degree=6)
    c.compute()
    return c


def test_c2_vertical_tube_and_section_everything_installed(ref, cm):
    """configs[1]: vertical orbit from the centre manifold, manifold.compute(step=0.005) -> 200 tubes, SynodicMap y = 0,
    (x, z), direction -1 -> the reference's 679 hits (count identical; points at the tolerance the orbit's own
    correction leaves)."""
    import hiten_b200
    from hiten import SynodicMap, VerticalOrbit
    g = G("synodic_c2.npz")
    l1 = ref.get_libration_point(1)
    hiten_b200.install()
    try:
        ic_seed = cm.to_synodic([0.0, 0.0], 0.6, "q3")
        orbit = VerticalOrbit(l1, initial_state=ic_seed)
        orbit.correct()
        assert np.abs(np.asarray(orbit.initial_state, float) - g["orbit_x0"]).max() <= 1e-9
        manifold = orbit.manifold(stable=True, direction="positive")
        _, _, states_list, _, successes, attempts = manifold.compute(step=0.005, show_progress=False)
        assert (successes, attempts) == (200, 200)
        smap = SynodicMap(manifold)
        smap.compute(section_axis="y", section_offset=0.0, plane_coords=("x", "z"), direction=-1)
        pts = np.asarray(smap.get_points())
        print(f"[dropin] C2 fully installed: {len(pts)} hits (reference 679)")
        assert pts.shape == (679, 2)
        d = np.abs(np.sort(pts[:, 0]) - np.sort(g["hit_point"][:, 0])).max()
        assert d <= 1e-5
    finally:
        hiten_b200.uninstall()


def test_c3_centre_manifold_map_with_and_without_the_drop_in(ref, cm):
    """configs[2] through the public API: cm.poincare_map(0.7).compute("p3") with the drop-in == without it, bit for
    bit (the CM path is bit-exact), on the real kernels."""
    import hiten_b200
    from hiten.algorithms.poincare.centermanifold.options import CenterManifoldMapOptions
    from hiten.algorithms.poincare.core.options import IterationOptions, SeedingOptions
    from hiten.algorithms.types.options import IntegrationOptions, WorkerOptions

    def run():
        pm = cm.poincare_map(energy=0.7)
        opts = CenterManifoldMapOptions(
            integration=IntegrationOptions(dt=0.01, order=4, c_omega_heuristic=20, max_steps=2000),
            iteration=IterationOptions(n_iter=2), seeding=SeedingOptions(n_seeds=20), workers=WorkerOptions(n_workers=1))
        pm.compute(section_coord="p3", options=opts)
        return np.asarray(pm.get_points(section_coord="p3"))

    want = run()
    hiten_b200.install()
    try:
        got = run()
    finally:
        hiten_b200.uninstall()
    assert got.shape == want.shape and got.shape[0] > 0
    assert np.array_equal(got, want)


def test_c5_heteroclinic_example_seam_bit_exact(ref):
    """configs[4] as examples/heteroclinic_connection.py runs it: both halos corrected and primed by the unpatched
    reference, then install(): the two manifold.compute() calls (incl. the energy filter that drops part of the L2
    tube) and ConnectionPipeline.solve() (two SynodicMap sections + the connection search) run on the GPU and return the
    reference's own 6 connections, bit for bit."""
    import hiten_b200
    from hiten.algorithms.connections import ConnectionPipeline
    from hiten.algorithms.connections.config import ConnectionConfig
    from hiten.algorithms.connections.options import ConnectionOptions
    from hiten.algorithms.poincare import SynodicMapConfig
    g = G("c5_connection.npz")
    mu = ref.mu
    l1, l2 = ref.get_libration_point(1), ref.get_libration_point(2)
    halo_l1 = l1.create_orbit("halo", amplitude_z=0.5, zenith="southern")
    halo_l1.correct()
    halo_l2 = l2.create_orbit("halo", amplitude_z=0.3663368, zenith="northern")
    halo_l2.correct()
    assert np.array_equal(np.asarray(halo_l1.initial_state, float), g["l1_orbit_x0"])
    assert np.array_equal(np.asarray(halo_l2.initial_state, float), g["l2_orbit_x0"])
    manifold_l1 = halo_l1.manifold(stable=True, direction="positive")
    manifold_l2 = halo_l2.manifold(stable=False, direction="negative")
    _prime(manifold_l1)
    _prime(manifold_l2)
    hiten_b200.install()
    try:
        r1 = manifold_l1.compute(integration_fraction=0.9, step=0.005, show_progress=False)
        r2 = manifold_l2.compute(integration_fraction=1.0, step=0.005, show_progress=False)
        for key, res in (("l1", r1), ("l2", r2)):
            states_list, successes, attempts = res[2], res[4], res[5]
            kept = g[f"{key}_kept"]
            assert (successes, attempts) == (int(kept.sum()), 200)
            assert np.array_equal(np.stack([s[0] for s in states_list]), g[f"{key}_x0W"][kept])
            assert np.array_equal(np.stack([s[-1] for s in states_list]), g[f"{key}_yf"][kept])
        section_cfg = SynodicMapConfig(section_axis="x", section_offset=1 - mu, plane_coords=("y", "z"))
        conn = ConnectionPipeline.with_default_engine(config=ConnectionConfig(section=section_cfg, direction=-1))
        result = conn.solve(manifold_l1, manifold_l2,
                            options=ConnectionOptions(delta_v_tol=1, ballistic_tol=1e-8, eps2d=1e-3))
        res = list(result.connections) if hasattr(result, "connections") else list(result)
    finally:
        hiten_b200.uninstall()
    assert len(res) == 6
    assert np.array_equal(np.array([r.delta_v for r in res]), g["conn_dv"])
    assert np.array_equal(np.array([0 if r.kind == "ballistic" else 1 for r in res]), g["conn_kind"])
    assert np.array_equal(np.array([r.point2d for r in res]).reshape(-1, 2), g["conn_pt"])
    assert np.array_equal(np.array([r.state_u for r in res]).reshape(-1, 6), g["conn_su"])
    assert np.array_equal(np.array([r.state_s for r in res]).reshape(-1, 6), g["conn_ss"])
    assert np.array_equal(np.array([r.trajectory_index_u for r in res]), g["conn_tiu"])
    assert np.array_equal(np.array([r.trajectory_index_s for r in res]), g["conn_tis"])


def test_propagate_dynsys_and_compute_stm_on_the_device(ref):
    """The two funnels everything else calls, rebound and run on the real kernels against the unpatched reference in the
    same process: 6-state grid bit for bit, 42-state STM within 1e-8 (BASELINE.json north_star; measured ~2e-11)."""
    import hiten_b200
    import hiten.algorithms.dynamics.base as dbase
    from hiten.algorithms.dynamics.rtbp import _compute_stm
    g = G("synodic_c1.npz")
    x0 = g["orbit_x0"]
    want = dbase._propagate_dynsys(ref.dynsys, x0, 0.0, 1.3, forward=-1, steps=50, flip_indices=slice(0, 6))
    x_ref, t_ref, phi_ref, PHI_ref = _compute_stm(ref.var_dynsys, x0, 1.1, steps=40, forward=-1)
    hiten_b200.install()
    try:
        got = dbase._propagate_dynsys(ref.dynsys, x0, 0.0, 1.3, forward=-1, steps=50, flip_indices=slice(0, 6))
        assert type(got) is type(want)
        assert np.array_equal(got.times, want.times) and np.array_equal(got.states, want.states)
        x, t, phi, PHI = _compute_stm(ref.var_dynsys, x0, 1.1, steps=40, forward=-1)
        assert np.array_equal(t, t_ref) and PHI.shape == PHI_ref.shape
        assert np.abs(PHI - PHI_ref).max() <= 1e-8 * np.abs(PHI_ref).max()
        print(f"[dropin] _compute_stm on the device vs reference: {np.abs(PHI - PHI_ref).max() / np.abs(PHI_ref).max():.2e}")
        with pytest.raises(ValueError):
            dbase._propagate_dynsys(ref.dynsys, x0[:5], 0.0, 1.0)
    finally:
        hiten_b200.uninstall()


def test_c3_1e5_seeds_through_the_public_api(ref, cm):
    """BASELINE configs[2] at its stated size from the user's call: install(cm_seeds_from_options=True), then
    cm.poincare_map(0.7).compute("p3") with SeedingOptions(n_seeds=100000): the seeding strategy's candidates are lifted
    in one hb_cm_lift batch, the engine's lifting loop reads that batch, the map runs on the GPU.  A 256-seed sample of
    the first backend request is recomputed by the unpatched reference backend: flags, states and times identical."""
    import hiten_b200
    from hiten.algorithms.poincare.centermanifold.backend import _CenterManifoldBackend
    from hiten.algorithms.poincare.centermanifold.options import CenterManifoldMapOptions
    from hiten.algorithms.poincare.centermanifold.types import CenterManifoldBackendRequest
    from hiten.algorithms.poincare.core.options import IterationOptions, SeedingOptions
    from hiten.algorithms.types.options import IntegrationOptions, WorkerOptions
    import time
    hiten_b200.install(cm_seeds_from_options=True)
    seen = []
    try:
        patched_run = _CenterManifoldBackend.run

        def spy(self, request):
            resp = patched_run(self, request)
            seen.append((request, resp))
            return resp

        _CenterManifoldBackend.run = spy
        pm = cm.poincare_map(energy=0.7)
        # the reference keys its section cache on the option NAMES only (services/base.py:155-172 turns the options dict
        # into a tuple of its keys), so an earlier compute("p3") at this energy would be returned as is: start clean
        pm.dynamics.clear()
        pm.dynamics.reset()
        opts = CenterManifoldMapOptions(
            integration=IntegrationOptions(dt=0.01, order=4, c_omega_heuristic=20, max_steps=2000),
            iteration=IterationOptions(n_iter=1), seeding=SeedingOptions(n_seeds=100_000), workers=WorkerOptions(n_workers=1))
        t0 = time.perf_counter()
        pm.compute(section_coord="p3", options=opts)
        dt = time.perf_counter() - t0
        pts = np.asarray(pm.get_points(section_coord="p3"))
        _CenterManifoldBackend.run = patched_run
    finally:
        hiten_b200.uninstall()
    req, resp = seen[0]
    n_seeds = len(req.seeds)
    print(f"[dropin] C3: {n_seeds} lifted seeds -> {len(pts)} section points in {dt:.1f} s wall (strategy + lifting + map)")
    assert n_seeds > 90_000 and len(pts) > 0.9 * n_seeds and len(np.unique(req.seeds, axis=0)) == n_seeds
    # the unpatched reference backend on a strided 256-seed sample of the same request
    pick = np.arange(0, n_seeds, n_seeds // 256)[:256]
    sub = CenterManifoldBackendRequest(seeds=req.seeds[pick], dt=req.dt, jac_H=req.jac_H, clmo_table=req.clmo_table,
                                       section_coord=req.section_coord, forward=req.forward, max_steps=req.max_steps,
                                       method=req.method, order=req.order, c_omega_heuristic=req.c_omega_heuristic)
    want = _CenterManifoldBackend().run(sub)                       # the reference's own code (uninstalled above)
    flags = np.asarray(resp.flags)
    assert np.array_equal(flags[pick], np.asarray(want.flags))
    rows = np.cumsum(flags.astype(bool)) - 1                        # response rows of the successful seeds
    ok = flags[pick].astype(bool)
    assert np.array_equal(np.asarray(resp.states)[rows[pick][ok]], np.asarray(want.states))
    assert np.array_equal(np.asarray(resp.times)[rows[pick][ok]], np.asarray(want.times))


def test_cubic_synodic_request_through_the_rebound_backend(ref):
    """interp_kind="cubic" (backend.py:762) is served by hb_synodic_detect_cubic under install(): the rebound
    _SynodicDetectionBackend.run returns the hits of the reference's own cubic branch bit for bit (golden
    synodic_cubic.npz and the unpatched backend run in this process)."""
    import dataclasses
    import hiten_b200
    from hiten.algorithms.poincare.synodic.backend import _SynodicDetectionBackend
    from hiten.algorithms.poincare.synodic.types import SynodicBackendRequest
    g = G("synodic_cubic.npz")
    tf, steps, fwd = float(g["l2_tf"]), int(g["l2_steps"]), int(g["l2_forward"])
    t_eval = np.linspace(0.0, tf, steps)
    dense = hiten_b200.cr3bp_dense(g["l2_x0W"], float(ref.mu), t_eval, forward=fwd, flip=(0, 6)).states
    normal = np.zeros(6); normal[1] = 1.0
    fields = {f.name for f in dataclasses.fields(SynodicBackendRequest)}
    kw = dict(trajectories=[(fwd * t_eval, d) for d in dense], trajectory_indices=list(range(len(dense))), normal=normal,
              offset=0.0, plane_coords=("x", "z"), interp_kind="cubic", segment_refine=50, tol_on_surface=1e-6,
              dedup_time_tol=1e-9, dedup_point_tol=1e-6, max_hits_per_traj=None, newton_max_iter=10, direction=-1)
    req = SynodicBackendRequest(**{k: v for k, v in kw.items() if k in fields})
    assert not hiten_b200.dropin.is_installed()
    want = _SynodicDetectionBackend().run(req)
    hiten_b200.install()
    try:
        got = _SynodicDetectionBackend().run(req)
    finally:
        hiten_b200.uninstall()
    assert _library_loaded() and len(want.times) == 72
    assert np.array_equal(got.times, want.times) and np.array_equal(got.states, want.states)
    assert np.array_equal(got.points, want.points) and np.array_equal(got.trajectory_indices, want.trajectory_indices)
    assert np.array_equal(got.times, g["l2_r50_dm_time"]) and np.array_equal(got.states, g["l2_r50_dm_state"])


def test_invariant_torus_and_compute_stm_variants_on_the_real_kernels(ref):
    """InvariantTori.compute() (its STM pass is _compute_stm with steps = n_theta1, types/services/torus.py:215) and
    _compute_stm(method="fixed" / order=5) under install() with the real library: equal to the unpatched reference within
    the 42-state tolerance."""
    import hiten_b200
    from hiten import InvariantTori
    from hiten.algorithms.dynamics import rtbp
    l1 = ref.get_libration_point(1)
    halo = l1.create_orbit("halo", amplitude_z=0.2, zenith="southern")
    halo.correct()
    halo.propagate()
    assert not hiten_b200.dropin.is_installed()
    want = np.asarray(InvariantTori(halo).compute(epsilon=1e-3, n_theta1=64, n_theta2=16))
    x0, T = np.asarray(halo.initial_state, float), float(halo.period)
    ref_rows = {}
    for key, kw in (("rk4", dict(method="fixed", order=4, steps=801)), ("rk45", dict(method="adaptive", order=5, steps=50))):
        ref_rows[key] = np.asarray(rtbp._compute_stm(ref.var_dynsys, x0, T, forward=1, **kw)[3])
    hiten_b200.install()
    try:
        got = np.asarray(InvariantTori(halo).compute(epsilon=1e-3, n_theta1=64, n_theta2=16))
        for key, kw in (("rk4", dict(method="fixed", order=4, steps=801)), ("rk45", dict(method="adaptive", order=5, steps=50))):
            rows = np.asarray(rtbp._compute_stm(ref.var_dynsys, x0, T, forward=1, **kw)[3])
            scale = np.abs(ref_rows[key][:, :36]).max(axis=1, keepdims=True)
            assert (np.abs(rows[:, :36] - ref_rows[key][:, :36]) / scale).max() <= 1e-8
            assert np.abs(rows[:, 36:] - ref_rows[key][:, 36:]).max() <= 1e-9
    finally:
        hiten_b200.uninstall()
    assert _library_loaded()
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-8
