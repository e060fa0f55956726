"""Golden vectors for the reference's other integrators on the CR3BP 6-state system (API parity rows a12/a13):
AdaptiveRK(order=5) = RK45 and RungeKutta(order=4/6/8) fixed-step, dense grids and terminal plane events.
Reference: algorithms/dynamics/base.py:346 (_propagate_dynsys method="adaptive"/"fixed"),
algorithms/integrators/rk.py:1138-1266 (_RK45.integrate), :422-529 (_FixedStepRK.integrate).
Writes tests/golden/rk_variants.npz.   Run: python tests/golden/make_rk_variants.py   (~4 min of JIT)
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
from hiten.algorithms.poincare.singlehit.backend import _g_y0, _get_cached_plane_event_fn  # noqa: E402
from hiten.algorithms.types.configs import EventConfig  # noqa: E402
from hiten.algorithms.types.options import EventOptions  # noqa: E402


def main():
    system = System.from_bodies("earth", "moon")
    g = np.load(os.path.join(os.path.dirname(__file__), "c1_manifold.npz"))
    x0 = g["halo_x0"].copy()
    T = float(g["halo_period"])
    dyn = system.dynsys
    out = {"mu": np.float64(system.mu), "x0": x0, "T": np.float64(T)}
    # dense / final, adaptive order 5
    for name, fwd in (("fwd", 1), ("bwd", -1)):
        sol = _propagate_dynsys(dyn, x0, 0.0, 2.0, forward=fwd, steps=41, method="adaptive", order=5,
                                flip_indices=slice(0, 6))
        out[f"rk45_dense_{name}"] = sol.states
        sol2 = _propagate_dynsys(dyn, x0, 0.0, 2.0, forward=fwd, steps=2, method="adaptive", order=5,
                                 flip_indices=slice(0, 6))
        out[f"rk45_final_{name}"] = sol2.states[-1]
    # fixed-step orders 4, 6, 8 on a 201-point grid
    for order in (4, 6, 8):
        sol = _propagate_dynsys(dyn, x0, 0.0, 1.0, forward=1, steps=201, method="fixed", order=order)
        out[f"rk{order}_dense"] = sol.states[::10]
        out[f"rk{order}_final"] = sol.states[-1]
        solb = _propagate_dynsys(dyn, x0, 0.0, 1.0, forward=-1, steps=101, method="fixed", order=order,
                                 flip_indices=slice(0, 6))
        out[f"rk{order}_final_bwd"] = solb.states[-1]
    # terminal events: y = 0 with direction -1 after leaving the plane, and x = offset plane
    y1 = _propagate_dynsys(dyn, x0, 0.0, 0.1 * T, steps=2).states[-1]
    out["y1"] = y1
    cfg = EventConfig(direction=-1, terminal=True)
    opt = EventOptions(xtol=1e-12, gtol=1e-12)
    sol = AdaptiveRK(order=5, max_step=1e4, rtol=1e-12, atol=1e-12).integrate(
        dyn, y1, np.array([0.0, T]), event_fn=_g_y0, event_cfg=cfg, event_options=opt)
    out["rk45_event_t"] = np.float64(sol.times[-1]); out["rk45_event_y"] = sol.states[-1]
    gx = _get_cached_plane_event_fn(0, float(x0[0]) + 0.003)
    sol = AdaptiveRK(order=5, max_step=1e4, rtol=1e-12, atol=1e-12).integrate(
        dyn, y1, np.array([0.0, T]), event_fn=gx, event_cfg=EventConfig(direction=0, terminal=True), event_options=opt)
    out["rk45_eventx_off"] = np.float64(float(x0[0]) + 0.003)
    out["rk45_eventx_t"] = np.float64(sol.times[-1]); out["rk45_eventx_y"] = sol.states[-1]
    tv = np.linspace(0.0, T, 1501)
    for order in (4, 8):
        sol = RungeKutta(order=order).integrate(dyn, y1, tv, event_fn=_g_y0, event_cfg=cfg, event_options=opt)
        out[f"rk{order}_event_t"] = np.float64(sol.times[-1]); out[f"rk{order}_event_y"] = sol.states[-1]
    # no-hit case: event never reached within a short window
    sol = RungeKutta(order=4).integrate(dyn, y1, np.linspace(0.0, 0.05, 11), event_fn=_g_y0, event_cfg=cfg, event_options=opt)
    out["rk4_nohit_t"] = np.float64(sol.times[-1]); out["rk4_nohit_y"] = sol.states[-1]
    path = os.path.join(os.path.dirname(__file__), "rk_variants.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
