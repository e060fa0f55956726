"""Golden STMs for BASELINE config 4 (halo family, 42-dim state+STM propagation), from the reference.

Family: examples/orbit_family.py (Earth-Moon L1 southern halo, continuation in z, max_members=100).
For every member (x0, T):   _compute_stm(var_dynsys, x0, T, steps=2000, forward=+1)  -> PHI[-1] (42 values)
(algorithms/dynamics/rtbp.py:258-340).  For the seed orbit additionally the BACKWARD STM the stable-manifold
service uses (forward=-1, only the state block flipped) and dense rows of PHI.
Writes tests/golden/stm_family.npz.   Run: python tests/golden/make_stm_family.py  (~3 min)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import _refenv  # noqa: E402

_refenv.enable()

from hiten import System  # noqa: E402
from hiten.algorithms.continuation.options import OrbitContinuationOptions  # noqa: E402
from hiten.algorithms.dynamics.rtbp import _compute_stm  # noqa: E402
from hiten.algorithms.types.states import SynodicState  # noqa: E402


def main():
    system = System.from_bodies("earth", "moon")
    l1 = system.get_libration_point(1)
    seed = l1.create_orbit("halo", amplitude_z=0.2, zenith="southern")
    seed.correct()
    seed.propagate()
    options = OrbitContinuationOptions(
        target=([seed.initial_state[SynodicState.Z], seed.initial_state[SynodicState.Y]],
                [seed.initial_state[SynodicState.Z] + 2.0, seed.initial_state[SynodicState.Y] - 1.0]),
        step=((1 - seed.initial_state[SynodicState.Z]) / (100 - 1), (1 - seed.initial_state[SynodicState.Y]) / (100 - 1)),
        max_members=100, max_retries_per_step=50, step_min=1e-10, step_max=1.0, shrink_policy=None,
        extra_params=seed.correction_options,
    )
    result = seed.generate(options)
    family = list(result.family)
    print("members:", len(family))
    var_sys = seed.dynamics.var_dynsys if hasattr(seed.dynamics, "var_dynsys") else seed.system.var_dynsys
    x0s = np.stack([np.asarray(o.initial_state, float) for o in family])
    Ts = np.array([float(o.period) for o in family])
    PHI_end = np.empty((len(family), 42))
    for i, (x0, T) in enumerate(zip(x0s, Ts)):
        _, _, _, PHI = _compute_stm(var_sys, x0, T, steps=2000, forward=1)
        PHI_end[i] = PHI[-1]
    x, times, phiT, PHI = _compute_stm(var_sys, x0s[0], Ts[0], steps=2000, forward=1)
    xb, timesb, phiTb, PHIb = _compute_stm(var_sys, x0s[0], Ts[0], steps=2000, forward=-1)
    idx = np.unique(np.concatenate([np.arange(0, 2000, 40), [1, 1998, 1999]]))
    out = os.path.join(os.path.dirname(__file__), "stm_family.npz")
    np.savez_compressed(out, mu=np.float64(system.mu), x0=x0s, period=Ts, PHI_end=PHI_end, dense_idx=idx,
                        PHI_fwd_dense=PHI[idx], PHI_bwd_dense=PHIb[idx], times_bwd_last=np.float64(timesb[-1]))
    print("wrote", out, "trace(Phi_T[0]) =", np.trace(phiT))


if __name__ == "__main__":
    main()
