"""CPU: the drop-in installer against the REAL reference objects (skipped where /root/reference is absent).

The GPU entry points are replaced by oracle-backed stand-ins (tests/fake_gpu.py), so what is tested here is the
host plumbing of hiten_b200.install(): name rebinding, system / event recognition, argument translation, result
types, filters -- by running the reference's own user-level calls with and without the drop-in.
"""
import os
import sys

import numpy as np
import pytest

REF_SRC = os.environ.get("HITEN_REFERENCE_SRC", "/root/reference/src")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference sources not present on this box")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import _refenv
    _refenv.enable()
    from hiten import System
    system = System.from_bodies("earth", "moon")
    l1 = system.get_libration_point(1)
    halo = l1.create_orbit("halo", amplitude_z=0.2, zenith="southern")
    halo.correct()
    return system, l1, halo


def test_install_rebinds_every_by_value_import(ref):
    import hiten_b200
    import hiten.algorithms.dynamics.base as dbase
    orig = dbase._propagate_dynsys
    hiten_b200.install()
    try:
        mods = [m for n, m in sys.modules.items() if n.startswith("hiten") and hasattr(m, "_propagate_dynsys")]
        assert len(mods) >= 6                                    # SURVEY 8b lists nine binding sites
        assert all(m._propagate_dynsys is not orig for m in mods)
        assert hiten_b200.dropin.is_installed()
    finally:
        hiten_b200.uninstall()
    assert dbase._propagate_dynsys is orig


def test_propagate_dynsys_and_stm_match_reference(ref, monkeypatch):
    import fake_gpu
    import hiten_b200
    from hiten.algorithms.dynamics.rtbp import _compute_stm
    import hiten.algorithms.dynamics.base as dbase
    system, l1, halo = ref
    x0 = np.asarray(halo.initial_state, float)
    want = dbase._propagate_dynsys(system.dynsys, x0, 0.0, 1.3, forward=-1, steps=50, flip_indices=slice(0, 6))
    x_ref, t_ref, phi_ref, PHI_ref = _compute_stm(system.var_dynsys, x0, 1.1, steps=40, forward=-1)
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    try:
        got = dbase._propagate_dynsys(system.dynsys, x0, 0.0, 1.3, forward=-1, steps=50, flip_indices=slice(0, 6))
        assert type(got) is type(want)
        assert np.array_equal(got.times, want.times) and np.array_equal(got.states, want.states)
        x, t, phi, PHI = _compute_stm(system.var_dynsys, x0, 1.1, steps=40, forward=-1)
        assert np.array_equal(t, t_ref) and PHI.shape == PHI_ref.shape
        assert np.abs(PHI - PHI_ref).max() <= 1e-10 * np.abs(PHI_ref).max()
        with pytest.raises(ValueError):                              # same validation error as the reference
            dbase._propagate_dynsys(system.dynsys, x0[:5], 0.0, 1.0)
    finally:
        hiten_b200.uninstall()


def test_manifold_and_synodic_map_through_the_public_api(ref, monkeypatch):
    import fake_gpu
    import hiten_b200
    from hiten import SynodicMap
    system, l1, halo = ref
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synodic_c1.npz"))
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    try:
        manifold = halo.manifold(stable=True, direction="positive")
        ysos, dysos, states_list, times_list, successes, attempts = manifold.compute(show_progress=False)
        assert (successes, attempts) == (50, 50) and states_list[0].shape == (4713, 6)
        # With the drop-in the orbit's STM pass runs on the 42-state path too, which agrees with the reference to
        # ~1e-13 (libm pow / 42-element np.dot order), so the tube's initial conditions agree to 1e-12 and the end
        # states to that times the manifold's instability.
        assert np.abs(np.stack([s[0] for s in states_list]) - g["x0W"]).max() <= 1e-12
        assert np.abs(np.stack([s[-1] for s in states_list]) - g["yf"]).max() <= 1e-6
        assert times_list[0][-1] == -float(g["tf"])
        smap = SynodicMap(manifold)
        smap.compute(section_axis="y", section_offset=0.0, plane_coords=("x", "z"), direction=-1)
        pts = np.asarray(smap.get_points())
        assert pts.shape == (121, 2)
        ref_pts = g["hit_point"]
        d = np.abs(np.sort(pts[:, 0]) - np.sort(ref_pts[:, 0])).max()
        assert d <= 1e-6
    finally:
        hiten_b200.uninstall()


def test_event_recognition_and_unknown_callables_fall_to_reference(ref, monkeypatch):
    import fake_gpu
    import hiten_b200
    from hiten_b200 import dropin
    from hiten.algorithms.poincare.singlehit.backend import _g_y0, _get_cached_plane_event_fn
    from hiten.algorithms.dynamics.rhs import create_rhs_system
    system, l1, halo = ref
    assert dropin.recognise_event(_g_y0) == (1, 0.0)
    assert dropin.recognise_event(_get_cached_plane_event_fn(0, 0.75)) == (0, 0.75)
    assert dropin.recognise_event(lambda t, y: y[0]) is None
    assert dropin.recognise_system(system.dynsys)[:2] == (6, float(system.mu))
    assert dropin.recognise_system(system.var_dynsys)[0] == 42
    user = create_rhs_system(lambda t, y: -y, dim=2, name="decay")
    assert dropin.recognise_system(user) is None


def test_centre_manifold_map_through_the_public_api(ref, monkeypatch):
    """cm.poincare_map(E).compute("p3") with the drop-in == without it, bit for bit (the CM path is bit-exact)."""
    import fake_gpu
    import hiten_b200
    from hiten.algorithms.poincare.centermanifold.options import CenterManifoldMapOptions
    from hiten.algorithms.poincare.core.options import IterationOptions, SeedingOptions
    from hiten.algorithms.types.options import IntegrationOptions, WorkerOptions
    system, l1, halo = ref
    cm = l1.get_center_manifold(degree=6)
    cm.compute()

    def run():
        pm = cm.poincare_map(energy=0.7)
        opts = CenterManifoldMapOptions(
            integration=IntegrationOptions(dt=0.01, order=4, c_omega_heuristic=20, max_steps=2000),
            iteration=IterationOptions(n_iter=2), seeding=SeedingOptions(n_seeds=20), workers=WorkerOptions(n_workers=1))
        pm.compute(section_coord="p3", options=opts)
        return np.asarray(pm.get_points(section_coord="p3"))

    want = run()
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    try:
        got = run()
    finally:
        hiten_b200.uninstall()
    assert got.shape == want.shape and got.shape[0] > 0
    assert np.array_equal(got, want)


def test_connections_backend_through_the_drop_in(ref, monkeypatch):
    """_ConnectionsBackend.run with the drop-in installed returns the reference's own result objects, identical to
    the unpatched backend."""
    import fake_gpu
    import hiten_b200
    from hiten.algorithms.connections.backends import _ConnectionsBackend
    from hiten.algorithms.connections.types import ConnectionsBackendRequest
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "connections.npz"))
    req = ConnectionsBackendRequest(points_u=g["b_pu"], points_s=g["b_ps"], states_u=g["b_Xu"], states_s=g["b_Xs"],
                                    traj_indices_u=g["b_tu"], traj_indices_s=g["b_ts"], eps=float(g["b_eps"]),
                                    dv_tol=float(g["b_dv_tol"]), bal_tol=float(g["b_bal_tol"]), metadata={"tag": 1})
    want = _ConnectionsBackend().run(req)
    fake_gpu.patch(monkeypatch)
    hiten_b200.install()
    try:
        got = _ConnectionsBackend().run(req)
    finally:
        hiten_b200.uninstall()
    assert got.metadata == want.metadata and len(got.results) == len(want.results) > 100
    for a, b in zip(got.results, want.results):
        assert (a.kind, a.delta_v, a.point2d, a.index_u, a.index_s, a.trajectory_index_u, a.trajectory_index_s) == \
               (b.kind, b.delta_v, b.point2d, b.index_u, b.index_s, b.trajectory_index_u, b.trajectory_index_s)
        assert np.array_equal(a.state_u, b.state_u) and np.array_equal(a.state_s, b.state_s)


def test_correct_many_applies_results_like_the_reference(ref, monkeypatch):
    """corrector.correct_many on reference orbit objects: options read from each orbit's own correction config,
    one batch per configuration, results applied through the orbit's correction service."""
    import fake_gpu
    from hiten_b200 import corrector
    system, l1, _ = ref
    fake_gpu.patch(monkeypatch)
    mk = lambda: [l1.create_orbit("halo", amplitude_z=az, zenith="southern") for az in (0.1, 0.25)] + \
        [l1.create_orbit("lyapunov", amplitude_x=0.02)]
    a, b = mk(), mk()
    op = corrector.opts_from_reference(a[0])
    assert tuple(op.ctrl) == (0, 4) and tuple(op.res) == (3, 5) and op.event_idx == 1 and op.halo_quadratic == 1
    assert corrector.opts_from_reference(a[2]).halo_quadratic == 0
    ref_res = [o.correct() for o in a]
    res = corrector.correct_many(b)
    for o_ref, o_new, r_ref, r_new in zip(a, b, ref_res, res):
        assert np.abs(np.asarray(o_new.initial_state) - np.asarray(o_ref.initial_state)).max() <= 1e-10
        assert abs(o_new.period - o_ref.period) <= 1e-10
        assert r_new.converged and abs(r_new.iterations - r_ref.iterations) <= 1
        assert abs(r_new.half_period - r_ref.half_period) <= 1e-10


def test_batched_corrector_rebinds_orbit_correct(ref, monkeypatch):
    """install(corrector="batched"): halo.correct() / lyapunov.correct() are ONE hb_correct_orbits call each, with the
    reference's return values, caching and orbit update."""
    import fake_gpu
    import hiten_b200
    from hiten.algorithms.types.services.orbits import _OrbitCorrectionService
    system, l1, _ = ref
    a = [l1.create_orbit("halo", amplitude_z=0.15, zenith="northern"), l1.create_orbit("lyapunov", amplitude_x=0.02)]
    b = [l1.create_orbit("halo", amplitude_z=0.15, zenith="northern"), l1.create_orbit("lyapunov", amplitude_x=0.02)]
    ref_res = [o.correct() for o in a]
    orig = _OrbitCorrectionService.correct
    hiten_b200.install(corrector="batched")
    fake_gpu.patch(monkeypatch)
    calls = []
    import hiten_b200.corrector as corr
    inner = corr.correct_orbits
    monkeypatch.setattr(corr, "correct_orbits", lambda *a_, **k: (calls.append(1), inner(*a_, **k))[1])
    try:
        assert _OrbitCorrectionService.correct is not orig
        for o_ref, o_new, r_ref in zip(a, b, ref_res):
            r_new = o_new.correct()
            assert r_new.converged and abs(r_new.iterations - r_ref.iterations) <= 1
            assert np.abs(np.asarray(o_new.initial_state) - np.asarray(o_ref.initial_state)).max() <= 1e-10
            assert abs(o_new.period - o_ref.period) <= 1e-10
            assert o_new.correct() is r_new                 # cached like the reference
        assert len(calls) == 2
    finally:
        hiten_b200.uninstall()
    assert _OrbitCorrectionService.correct is orig
    with pytest.raises(ValueError):
        hiten_b200.install(corrector="nope")


def test_rk45_and_fixed_step_classes_and_propagate_variants(ref, monkeypatch):
    """_RK45.integrate / _FixedStepRK.integrate and _propagate_dynsys(method="fixed" | order=5) with the drop-in:
    the same arrays as the reference alone, bit for bit (grid integration and plane events)."""
    import fake_gpu
    import hiten_b200
    import hiten.algorithms.dynamics.base as dbase
    from hiten.algorithms.integrators.rk import RungeKutta
    from hiten.algorithms.poincare.singlehit.backend import _g_y0
    from hiten.algorithms.types.configs import EventConfig
    system, l1, halo = ref
    x0 = np.asarray(halo.initial_state, dtype=np.float64)
    dyn = system.dynsys
    grid = np.linspace(0.0, 1.5, 151)
    cfg = EventConfig(direction=0, terminal=True)
    y1 = dbase._propagate_dynsys(dyn, x0, 0.0, 0.05, steps=2).states[-1]

    def run_all():
        out = {}
        for order in (4, 6, 8, 45):
            integ = RungeKutta(order=order) if order != 45 else RungeKutta(order=45, rtol=1e-10, atol=1e-10)
            sol = integ.integrate(dyn, x0, grid)
            out[f"dense{order}"] = (sol.times.copy(), sol.states.copy())
            ev = integ.integrate(dyn, y1, np.linspace(0.0, 2.0, 2001), event_fn=_g_y0, event_cfg=cfg)
            out[f"event{order}"] = (ev.times.copy(), ev.states.copy())
        for method, order in (("fixed", 4), ("fixed", 8), ("adaptive", 5)):
            sol = dbase._propagate_dynsys(dyn, x0, 0.0, 1.0, forward=-1, steps=101, method=method, order=order)
            out[f"prop_{method}{order}"] = (sol.times.copy(), sol.states.copy())
        return out

    want = run_all()
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    calls = []
    import hiten_b200.propagate as prop
    d0, e0 = prop.cr3bp_dense, prop.cr3bp_event
    monkeypatch.setattr(prop, "cr3bp_dense", lambda *a, **k: (calls.append("dense"), d0(*a, **k))[1])
    monkeypatch.setattr(prop, "cr3bp_event", lambda *a, **k: (calls.append("event"), e0(*a, **k))[1])
    try:
        got = run_all()
    finally:
        hiten_b200.uninstall()
    assert calls.count("dense") == 7 and calls.count("event") == 4          # every call went through the GPU entry points
    for k in want:
        assert np.array_equal(got[k][0], want[k][0]), k
        assert np.array_equal(got[k][1], want[k][1]), k


def test_symplectic_class_and_propagate_symplectic(ref, monkeypatch):
    """_ExtendedSymplectic.integrate (grid + plane events, both time directions) and
    _propagate_dynsys(method="symplectic") with the drop-in: the same arrays as the reference alone, bit for bit --
    including the reference's double application of the direction sign to the times of a backward propagation."""
    import fake_gpu
    import hiten_b200
    import hiten.algorithms.dynamics.base as dbase
    from hiten.algorithms.dynamics.base import _DirectedSystem
    from hiten.algorithms.integrators.symplectic import _ExtendedSymplectic
    from hiten.algorithms.poincare.singlehit.backend import _get_cached_plane_event_fn
    from hiten.algorithms.types.configs import EventConfig
    system, l1, halo = ref
    cm = l1.get_center_manifold(degree=6)
    cm.compute()
    hamsys = cm.poincare_map(energy=0.7).dynamics.hamsys
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "symplectic.npz"))
    y0 = g["y0"][1]
    fn = _get_cached_plane_event_fn(2, 0.0)

    def run_all():
        out = {}
        for order, fwd in ((4, 1), (2, -1), (6, 1)):
            integ = _ExtendedSymplectic(order=order)
            sol = integ.integrate(_DirectedSystem(hamsys, fwd), y0.copy(), np.linspace(0.0, 1.0, 81))
            out[f"grid{order}{fwd}"] = (sol.times.copy(), sol.states.copy())
            ev = integ.integrate(_DirectedSystem(hamsys, fwd), y0.copy(), np.linspace(0.0, 6.0, 601), event_fn=fn,
                                 event_cfg=EventConfig(direction=0, terminal=True))
            out[f"event{order}{fwd}"] = (ev.times.copy(), ev.states.copy())
        nohit = _ExtendedSymplectic(order=4).integrate(hamsys, y0.copy(), np.linspace(0.0, 0.05, 6),
                                                       event_fn=_get_cached_plane_event_fn(2, 10.0),
                                                       event_cfg=EventConfig(direction=0, terminal=True))
        out["nohit"] = (nohit.times.copy(), nohit.states.copy())
        for fwd in (1, -1):
            sol = dbase._propagate_dynsys(hamsys, y0.copy(), 0.0, 1.0, forward=fwd, steps=41, method="symplectic", order=4)
            out[f"prop{fwd}"] = (sol.times.copy(), sol.states.copy())
        return out

    want = run_all()
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    calls = []
    import hiten_b200.symplectic as symp
    d0, e0 = symp.integrate_symplectic, symp.integrate_symplectic_until_event
    monkeypatch.setattr(symp, "integrate_symplectic", lambda *a, **k: (calls.append("grid"), d0(*a, **k))[1])
    monkeypatch.setattr(symp, "integrate_symplectic_until_event", lambda *a, **k: (calls.append("event"), e0(*a, **k))[1])
    try:
        got = run_all()
        # a user-defined event callable is not expressible on the GPU path: the reference's own method runs
        n_before = len(calls)
        _ExtendedSymplectic(order=4).integrate(hamsys, y0.copy(), np.linspace(0.0, 0.1, 11),
                                               event_fn=lambda t, y: y[2] - 10.0,
                                               event_cfg=EventConfig(direction=0, terminal=True))
        assert len(calls) == n_before
    finally:
        hiten_b200.uninstall()
    assert calls.count("grid") == 5 and calls.count("event") == 4
    assert want["prop-1"][0][-1] > 0                                      # the reference's sign quirk
    for k in want:
        assert np.array_equal(got[k][0], want[k][0]), k
        assert np.array_equal(got[k][1], want[k][1]), k


def test_fixed_step_rk_classes_on_the_hamiltonian_system(ref, monkeypatch):
    """RungeKutta(order=4|6|8).integrate(hamsys, ...) -- the `_ham` kernels of _FixedStepRK (grid with derivatives, plane
    events, no-hit) -- with the drop-in: the same arrays as the reference alone, bit for bit; a _DirectedSystem around the
    Hamiltonian system is left to the reference's own method (which raises, with or without the drop-in)."""
    import fake_gpu
    import hiten_b200
    from hiten.algorithms.dynamics.base import _DirectedSystem
    from hiten.algorithms.integrators.rk import AdaptiveRK, RungeKutta
    from hiten.algorithms.poincare.singlehit.backend import _get_cached_plane_event_fn
    from hiten.algorithms.types.configs import EventConfig
    system, l1, halo = ref
    cm = l1.get_center_manifold(degree=6)
    cm.compute()
    hamsys = cm.poincare_map(energy=0.7).dynamics.hamsys
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "symplectic.npz"))
    y0 = g["y0"][2]
    cfg = EventConfig(direction=0, terminal=True)

    def run_all():
        out = {}
        for order in (4, 6, 8):
            sol = RungeKutta(order=order).integrate(hamsys, y0.copy(), np.linspace(0.0, 1.0, 51))
            out[f"grid{order}"] = (sol.times.copy(), sol.states.copy(), sol.derivatives.copy())
            ev = RungeKutta(order=order).integrate(hamsys, y0.copy(), np.linspace(0.0, 6.0, 301),
                                                   event_fn=_get_cached_plane_event_fn(2, 0.0), event_cfg=cfg)
            out[f"event{order}"] = (ev.times.copy(), ev.states.copy())
        nh = RungeKutta(order=4).integrate(hamsys, y0.copy(), np.linspace(0.0, 0.05, 6),
                                           event_fn=_get_cached_plane_event_fn(2, 10.0), event_cfg=cfg)
        out["nohit"] = (nh.times.copy(), nh.states.copy())
        for order in (5, 8):                                           # _RK45 / _DOP853 `_ham` kernels
            integ = AdaptiveRK(order=order, rtol=1e-11, atol=1e-12)
            sol = integ.integrate(hamsys, y0.copy(), np.linspace(0.0, 2.0, 41))
            out[f"agrid{order}"] = (sol.times.copy(), sol.states.copy(), sol.derivatives.copy())
            ev = integ.integrate(hamsys, y0.copy(), np.linspace(0.0, 6.0, 301),
                                 event_fn=_get_cached_plane_event_fn(2, 0.0), event_cfg=cfg)
            out[f"aevent{order}"] = (ev.times.copy(), ev.states.copy())
            nh = integ.integrate(hamsys, y0.copy(), np.linspace(0.0, 0.05, 6),
                                 event_fn=_get_cached_plane_event_fn(2, 10.0), event_cfg=cfg)
            out[f"anohit{order}"] = (nh.times.copy(), nh.states.copy())
        return out

    want = run_all()
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    calls = []
    import hiten_b200.symplectic as symp
    d0, e0 = symp.integrate_rk_ham, symp.integrate_rk_ham_until_event
    monkeypatch.setattr(symp, "integrate_rk_ham", lambda *a, **k: (calls.append("grid"), d0(*a, **k))[1])
    monkeypatch.setattr(symp, "integrate_rk_ham_until_event", lambda *a, **k: (calls.append("event"), e0(*a, **k))[1])
    d1, e1 = symp.integrate_adaptive_ham, symp.integrate_adaptive_ham_until_event
    monkeypatch.setattr(symp, "integrate_adaptive_ham", lambda *a, **k: (calls.append("agrid"), d1(*a, **k))[1])
    monkeypatch.setattr(symp, "integrate_adaptive_ham_until_event", lambda *a, **k: (calls.append("aevent"), e1(*a, **k))[1])
    try:
        got = run_all()
        with pytest.raises(Exception):
            RungeKutta(order=4).integrate(_DirectedSystem(hamsys, -1), y0.copy(), np.linspace(0.0, 1.0, 11))
    finally:
        hiten_b200.uninstall()
    assert calls.count("grid") == 3 and calls.count("event") == 4
    assert calls.count("agrid") == 2 and calls.count("aevent") == 4
    for k in want:
        for a, b in zip(got[k], want[k]):
            assert np.array_equal(a, b), k


def test_centre_manifold_seeding_is_batched_and_identical(ref, monkeypatch):
    """Every seeding strategy with the drop-in: the candidates are lifted in ONE batch (no per-point Brent solve in
    `_build_seed`, none in the engine's lifting loop), and the seeds the strategy returns / the states the engine lifts
    are identical to the reference alone.  With cm_seeds_from_options the options' n_seeds reaches the strategy."""
    import fake_gpu
    import hiten_b200
    from hiten_b200 import centermanifold as cmod
    from hiten.algorithms.poincare.centermanifold.config import CenterManifoldMapConfig
    from hiten.algorithms.poincare.centermanifold.interfaces import _CenterManifoldInterface
    from hiten.algorithms.poincare.centermanifold.strategies import _make_strategy
    system, l1, halo = ref
    cm = l1.get_center_manifold(degree=6)
    cm.compute()
    pm = cm.poincare_map(energy=0.7)
    hamsys = pm.dynamics.hamsys
    H_blocks, clmo = hamsys.poly_H(), hamsys.clmo_table
    iface = _CenterManifoldInterface()
    solve = lambda var, fixed: iface.solve_missing_coord(var, fixed, h0=0.7, H_blocks=H_blocks, clmo_table=clmo)
    turn = lambda name: iface.find_turning(name, h0=0.7, H_blocks=H_blocks, clmo_table=clmo)

    def run(strategy, section):
        from hiten.algorithms.types.configs import IntegrationConfig
        cfg = CenterManifoldMapConfig(section_coord=section, seed_strategy=strategy,
                                      seed_axis=("q2" if section in ("q3", "p3") else "q3") if strategy == "single" else None,
                                      integration=IntegrationConfig(method="symplectic"))
        st = _make_strategy(cfg)
        pts = st.generate(h0=0.7, H_blocks=H_blocks, clmo_table=clmo, solve_missing_coord_fn=solve, find_turning_fn=turn)
        lifted = [iface.lift_plane_point(p, section_coord=section, h0=0.7, H_blocks=H_blocks, clmo_table=clmo) for p in pts]
        return pts, lifted

    cases = [("single", "p3"), ("axis_aligned", "q3"), ("level_sets", "p3"), ("radial", "q2")]
    want = {c: run(*c) for c in cases}
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    calls, brent = [], []
    inner = cmod.lift_plane_points
    monkeypatch.setattr(cmod, "lift_plane_points", lambda *a, **k: (calls.append(len(a[2])), inner(*a, **k))[1])
    try:
        for c in cases:
            n_before = len(calls)
            got = run(*c)
            assert len(calls) == n_before + 1 and calls[-1] >= len(got[0])   # one batch per generate()
            assert got[0] == want[c][0] and got[1] == want[c][1], c
        # random seeding: reproducible inside one generate() (probe and replay see the same draws), valid seeds only
        pts, lifted = run("random", "p3")
        assert len(pts) == 20 and all(s is not None for s in lifted)
        # n_seeds from the options reaches the strategies only when asked for
        from hiten.algorithms.poincare.centermanifold.options import CenterManifoldMapOptions
        from hiten.algorithms.poincare.core.options import IterationOptions, SeedingOptions
        from hiten.algorithms.types.options import IntegrationOptions, WorkerOptions
        opts = CenterManifoldMapOptions(
            integration=IntegrationOptions(dt=0.01, order=4, c_omega_heuristic=20, max_steps=2000),
            iteration=IterationOptions(n_iter=1), seeding=SeedingOptions(n_seeds=64), workers=WorkerOptions(n_workers=1))
        pm0 = cm.poincare_map(energy=0.66)                  # an energy no other test used: the reference caches sections
        pm0.compute(section_coord="p3", options=opts)       # by option NAMES only (services/base.py:155-172)
        n_default = len(np.asarray(pm0.get_points(section_coord="p3")))
        hiten_b200.install(cm_seeds_from_options=True)
        pm2 = cm.poincare_map(energy=0.65)                  # another energy: nothing cached
        pm2.compute(section_coord="p3", options=opts)
        n_more = len(np.asarray(pm2.get_points(section_coord="p3")))
        assert n_default <= 20 < n_more <= 64
        # the engine's lifting loop of a compute("p3") call is answered from a batch as well, although the strategy
        # validates its candidates on ITS config's section (q3): no per-seed Brent solve is left (only turning points)
        solves = []
        orig_solve = _CenterManifoldInterface.solve_missing_coord
        monkeypatch.setattr(_CenterManifoldInterface, "solve_missing_coord",
                            lambda self, *a, **k: (solves.append(1), orig_solve(self, *a, **k))[1])
        pm3 = cm.poincare_map(energy=0.64)
        n_batches = len(calls)
        pm3.compute(section_coord="p3", options=opts)
        assert len(np.asarray(pm3.get_points(section_coord="p3"))) > 20
        assert len(calls) == n_batches + 2 and len(solves) <= 4, (len(calls) - n_batches, len(solves))
    finally:
        hiten_b200.uninstall()


def test_cubic_synodic_request_goes_through_the_drop_in(ref, monkeypatch):
    """A request with interp_kind="cubic" (backend.py:762; never issued by the shipped SynodicMap) is served by the
    rebound backend too: same hits as the reference's own cubic branch, bit for bit."""
    import fake_gpu
    import hiten_b200
    import oracle_lib as O
    from hiten.algorithms.poincare.synodic.backend import _SynodicDetectionBackend
    from hiten.algorithms.poincare.synodic.types import SynodicBackendRequest
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synodic_cubic.npz"))
    tf, steps, fwd = float(g["l2_tf"]), int(g["l2_steps"]), int(g["l2_forward"])
    t_eval = np.linspace(0.0, tf, steps)
    dense, _ = O.batch_dense(O.system(O.SYS_CR3BP6, 0.012154535289174722, fwd=fwd, flip=(0, 6)), O.DOP853,
                             O.default_tol(), g["l2_x0W"][:6], t_eval, 4)
    normal = np.zeros(6); normal[1] = 1.0
    import dataclasses
    fields = {f.name for f in dataclasses.fields(SynodicBackendRequest)}
    kw = dict(trajectories=[(fwd * t_eval, d) for d in dense], trajectory_indices=list(range(len(dense))), normal=normal,
              offset=0.0, plane_coords=("x", "z"), interp_kind="cubic", segment_refine=50, tol_on_surface=1e-6,
              dedup_time_tol=1e-9, dedup_point_tol=1e-6, max_hits_per_traj=None, newton_max_iter=10, direction=-1)
    req = SynodicBackendRequest(**{k: v for k, v in kw.items() if k in fields})
    want = _SynodicDetectionBackend().run(req)
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    try:
        got = _SynodicDetectionBackend().run(req)
    finally:
        hiten_b200.uninstall()
    assert len(want.times) > 0 and np.array_equal(got.times, want.times)
    assert np.array_equal(got.states, want.states) and np.array_equal(got.points, want.points)
    assert np.array_equal(got.trajectory_indices, want.trajectory_indices)
    sel = g["l2_r50_dm_traj"] < 6
    assert np.array_equal(got.times, g["l2_r50_dm_time"][sel]) and np.array_equal(got.states, g["l2_r50_dm_state"][sel])


def test_compute_stm_with_the_other_integrators_goes_through_the_drop_in(ref, monkeypatch):
    """_compute_stm(..., method="fixed", order=4 | 8) and (method="adaptive", order=5) reach the 42-state kernels under
    install() (rtbp.py:258-340 -> _propagate_dynsys); results agree with the reference alone within the 42-state
    tolerance and with the golden rows."""
    import fake_gpu
    import hiten_b200
    from hiten.algorithms.dynamics import rtbp
    system, l1, halo = ref
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, "golden", "stm_variants.npz"))
    fam = np.load(os.path.join(here, "golden", "stm_family.npz"))
    x0, T = fam["x0"][0], float(fam["period"][0])
    calls = []
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    import hiten_b200.propagate as prop
    orig_dense = prop.cr3bp_stm_dense

    def spy(*a, **kw):
        calls.append(kw["integ"].method)
        return orig_dense(*a, **kw)

    monkeypatch.setattr(prop, "cr3bp_stm_dense", spy)
    try:
        for name in ("rk4_fwd", "rk8_fwd", "rk45_bwd"):
            kind, order, steps, fwd, frac = g[f"case_{name}"]
            x, times, phiT, PHI = rtbp._compute_stm(system.var_dynsys, x0, float(frac) * T, steps=int(steps), forward=int(fwd),
                                                    method="fixed" if kind else "adaptive", order=int(order))
            ref_rows = g[f"{name}_m0_PHI"]
            rows = np.asarray(PHI)[g[f"{name}_idx"]]
            scale = np.abs(ref_rows[:, :36]).max(axis=1, keepdims=True)
            assert (np.abs(rows[:, :36] - ref_rows[:, :36]) / scale).max() <= 1e-9
            assert times[-1] == g[f"{name}_m0_tlast"]
    finally:
        hiten_b200.uninstall()
    assert calls == [4, 8, 45]


def test_invariant_torus_stm_pass_runs_on_the_42_state_path(ref, monkeypatch):
    """_TorusDynamicsService.prepare (types/services/torus.py:215) calls _compute_stm(var_dynsys, x0, T, steps=n_theta1):
    under install() that is ONE dense 42-state launch; the torus grid equals the reference's own within the STM tolerance."""
    import fake_gpu
    import hiten_b200
    import hiten_b200.propagate as prop
    from hiten import InvariantTori
    system, l1, halo = ref
    halo.propagate()
    want = np.asarray(InvariantTori(halo).compute(epsilon=1e-3, n_theta1=64, n_theta2=16))
    calls = []
    hiten_b200.install()
    fake_gpu.patch(monkeypatch)
    orig_dense = prop.cr3bp_stm_dense

    def spy(x0, mu, t_eval, **kw):
        calls.append(len(t_eval))
        return orig_dense(x0, mu, t_eval, **kw)

    monkeypatch.setattr(prop, "cr3bp_stm_dense", spy)
    try:
        got = np.asarray(InvariantTori(halo).compute(epsilon=1e-3, n_theta1=64, n_theta2=16))
    finally:
        hiten_b200.uninstall()
    assert 64 in calls
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-9
