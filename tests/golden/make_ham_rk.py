"""Golden vectors for the `_ham` variants of the RK integrator classes (SURVEY 8a7-a9, a12, a13 "+_ham"): the reference's
RungeKutta / AdaptiveRK classes applied to its polynomial `_HamiltonianSystem` (algorithms/integrators/rk.py:
`_integrate_fixed_rk_ham` :592, `_integrate_fixed_rk_until_event_ham` :722, `_integrate_rk45_ham` :1403,
`_integrate_rk45_until_event_ham` :1589, `_integrate_dop853_ham` :2553, `_integrate_dop853_until_event_ham` :2807).

System: the degree-6 Earth-Moon L1 centre-manifold Hamiltonian of tests/golden/cm_map.npz (same term tables, asserted).
Also records that the reference RAISES for a `_DirectedSystem(hamsys, ...)` in the RK classes (and therefore for
`_propagate_dynsys(hamsys, method="fixed" | "adaptive")`): the `_ham` kernels are reached by direct class use only.
Writes tests/golden/ham_rk.npz.   Run: python tests/golden/make_ham_rk.py   (~4 min)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import _refenv  # noqa: E402

_refenv.enable()

from hiten import System  # noqa: E402
from hiten.algorithms.dynamics.base import _DirectedSystem, _propagate_dynsys  # noqa: E402
from hiten.algorithms.integrators.rk import AdaptiveRK, RungeKutta  # noqa: E402
from hiten.algorithms.poincare.singlehit.backend import _get_cached_plane_event_fn  # noqa: E402
from hiten.algorithms.types.configs import EventConfig  # noqa: E402

from make_cm_map import sparse_terms  # noqa: E402


def main():
    here = os.path.dirname(__file__)
    cmg = np.load(os.path.join(here, "cm_map.npz"))
    system = System.from_bodies("earth", "moon")
    cm = system.get_libration_point(1).get_center_manifold(degree=6)
    cm.compute()
    hamsys = cm.poincare_map(energy=0.7).dynamics.hamsys
    coefs = np.concatenate([sparse_terms(hamsys.jac_H[p], hamsys.clmo_table)[1] for p in range(6)])
    assert np.array_equal(coefs, cmg["jac_coef"]), "term tables differ from cm_map.npz"
    seeds = cmg["seeds_p3"][:6]
    y0 = np.zeros((6, 6))
    y0[:, 1], y0[:, 4], y0[:, 2], y0[:, 5] = seeds[:, 0], seeds[:, 1], seeds[:, 2], seeds[:, 3]
    out = {"y0": y0}
    cfg = EventConfig(direction=0, terminal=True)
    fn = _get_cached_plane_event_fn(2, 0.0)

    def integ(order):
        return RungeKutta(order=order) if order in (4, 6, 8) else AdaptiveRK(order={853: 8, 45: 5}[order], rtol=1e-11, atol=1e-12)

    for order in (4, 6, 8, 45, 853):
        grid = np.linspace(0.25, 2.25, 201) if order in (4, 6, 8) else np.linspace(0.25, 4.25, 81)
        ev_grid = np.linspace(0.0, 6.0, 601)
        dense, derivs, events = [], [], []
        for i in range(4):
            sol = integ(order).integrate(hamsys, y0[i].copy(), grid)
            assert np.array_equal(sol.times, grid)
            dense.append(sol.states); derivs.append(sol.derivatives)
            ev = integ(order).integrate(hamsys, y0[i].copy(), ev_grid, event_fn=fn, event_cfg=cfg)
            assert ev.times.shape == (2,)
            events.append(np.concatenate([[ev.times[1]], ev.states[1]]))
        out[f"grid_{order}"] = grid
        out[f"dense_{order}"] = np.array(dense)
        out[f"derivs_{order}"] = np.array(derivs)
        out[f"event_{order}"] = np.array(events)
        # no-hit case: the class returns [t0, t_last] and [y0, y_last]
        nh = integ(order).integrate(hamsys, y0[0].copy(), np.linspace(0.0, 0.05, 6),
                                    event_fn=_get_cached_plane_event_fn(2, 10.0), event_cfg=cfg)
        out[f"nohit_{order}"] = np.concatenate([[nh.times[-1]], nh.states[-1]])
        print(order, "dense", out[f"dense_{order}"].shape, "event t", [f"{e[0]:.6f}" for e in events], flush=True)

    # A _DirectedSystem(hamsys, +-1) is not a _HamiltonianSystemProtocol instance: the RK classes then take the generic
    # closure path, which Numba cannot type for this system -- so _propagate_dynsys(hamsys, method="fixed" | "adaptive")
    # raises in the reference, and the `_ham` kernels are reachable only through direct class use with the bare system.
    raised = []
    for call in (lambda: RungeKutta(order=4).integrate(_DirectedSystem(hamsys, -1), y0[0].copy(), np.linspace(0, 1, 11)),
                 lambda: _propagate_dynsys(hamsys, y0[1].copy(), 0.0, 1.0, forward=1, steps=11, method="fixed", order=4)):
        try:
            call()
            raised.append(False)
        except Exception as e:                                  # noqa: BLE001
            raised.append(True)
            print("reference raises:", type(e).__name__, flush=True)
    out["directed_or_propagate_raises"] = np.array(raised)
    path = os.path.join(here, "ham_rk.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
