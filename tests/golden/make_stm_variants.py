"""Golden 42-state (state + STM) propagations with the reference's OTHER integrators, from the reference:
_compute_stm(var_dynsys, x0, tf, steps, forward, method, order) (algorithms/dynamics/rtbp.py:258-340) ->
_propagate_dynsys(method="adaptive", order=5) = RK45 (rk.py:1138-1399) and method="fixed", order=4 | 6 | 8 = RungeKutta
on the linspace grid (rk.py:422-588), flip_indices = slice(36, 42).
Orbit: member 0 and member 60 of the halo family of stm_family.npz (x0, period).
Writes tests/golden/stm_variants.npz: per case PHI rows at `dense_idx` (and the last row).
Run: python tests/golden/make_stm_variants.py   (~2 min)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import _refenv  # noqa: E402

_refenv.enable()

from hiten import System  # noqa: E402
from hiten.algorithms.dynamics.rtbp import _compute_stm  # noqa: E402

# name: (method, order, steps, forward, fraction of the period)
CASES = {
    "rk45_fwd": ("adaptive", 5, 200, 1, 1.0),
    "rk45_bwd": ("adaptive", 5, 200, -1, 1.0),
    "rk4_fwd": ("fixed", 4, 2001, 1, 1.0),
    "rk6_fwd": ("fixed", 6, 1001, 1, 1.0),
    "rk8_fwd": ("fixed", 8, 401, 1, 1.0),
    "rk8_bwd": ("fixed", 8, 401, -1, 0.5),
    "rk4_bwd": ("fixed", 4, 1001, -1, 0.5),
}


def main():
    here = os.path.dirname(__file__)
    system = System.from_bodies("earth", "moon")
    var_sys = system.var_dynsys
    g = np.load(os.path.join(here, "stm_family.npz"))
    out = {"mu": np.float64(system.mu), "case_names": np.array(list(CASES)), "members": np.array([0, 60])}
    for name, (method, order, steps, fwd, frac) in CASES.items():
        out[f"case_{name}"] = np.array([{"adaptive": 0, "fixed": 1}[method], order, steps, fwd, frac], dtype=np.float64)
        idx = np.unique(np.concatenate([np.arange(0, steps, max(1, steps // 25)), [1, steps - 2, steps - 1]]))
        out[f"{name}_idx"] = idx
        for mem in (0, 60):
            x0, T = g["x0"][mem], float(g["period"][mem])
            x, times, phiT, PHI = _compute_stm(var_sys, x0, frac * T, steps=steps, forward=fwd, method=method, order=order)
            out[f"{name}_m{mem}_PHI"] = np.asarray(PHI, float)[idx]
            out[f"{name}_m{mem}_tlast"] = np.float64(times[-1])
            print(name, mem, "trace", np.trace(phiT))
    np.savez_compressed(os.path.join(here, "stm_variants.npz"), **out)


if __name__ == "__main__":
    main()
