"""Golden vectors mirroring the reference's OWN test-suite of the Tao integrator
(hiten/algorithms/integrators/_tests/test_symplectic.py): the Taylor pendulum
H = p1^2/2 - (1 - q1^2/2 + q1^4/24 - q1^6/720) as a degree-6 polynomial Hamiltonian built with the reference's polynomial
tools (test_symplectic.py:42-103), integrated with `_ExtendedSymplectic.integrate` in the four configurations its tests
use (energy conservation :106-127, reversibility :130-156 -- a DESCENDING time grid --, final-state error :159-186,
comparison with solve_ivp :189-270).  Stores the sparse term tables (gradient and H) and the trajectories.
Writes tests/golden/pendulum.npz.   Run: python tests/golden/make_pendulum.py   (~1 min)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import _refenv  # noqa: E402

_refenv.enable()

from numba.typed import List  # noqa: E402

from hiten.algorithms.dynamics.hamiltonian import create_hamiltonian_system  # noqa: E402
from hiten.algorithms.integrators.rk import AdaptiveRK, RungeKutta  # noqa: E402
from hiten.algorithms.integrators.symplectic import _ExtendedSymplectic  # noqa: E402
from hiten.algorithms.polynomial.base import (_create_encode_dict_from_clmo, _encode_multiindex,  # noqa: E402
                                              _init_index_tables)

from make_cm_map import sparse_terms  # noqa: E402

DEG = 6


def main():
    psi, clmo = _init_index_tables(DEG)
    enc = _create_encode_dict_from_clmo(clmo)
    H = [np.zeros(psi[6, d], dtype=np.complex128) for d in range(DEG + 1)]

    def put(var, power, coef):
        k = np.zeros(6, dtype=np.int64)
        if power:
            k[var] = power
        H[power][_encode_multiindex(k, power, enc)] += coef

    put(3, 2, 0.5); put(0, 0, -1.0); put(0, 2, 0.5); put(0, 4, -1.0 / 24.0); put(0, 6, 1.0 / 720.0)
    Hn = List()
    for a in H:
        Hn.append(a.copy())
    hamsys = create_hamiltonian_system(H_blocks=Hn, degree=DEG, psi_table=psi, clmo_table=clmo, encode_dict_list=enc,
                                       n_dof=3, name="Test Pendulum System")
    out = {}
    ptr, degs, coefs, exps = [0], [], [], []
    for p in range(6):
        d, c, e, _ = sparse_terms(hamsys.jac_H[p], clmo)
        degs.append(d); coefs.append(c); exps.append(e); ptr.append(ptr[-1] + len(d))
    out.update(jac_ptr=np.array(ptr, dtype=np.int64), jac_deg=np.concatenate(degs), jac_coef=np.concatenate(coefs),
               jac_exp=np.concatenate(exps).reshape(-1, 6))
    d, c, e, _ = sparse_terms(Hn, clmo)
    out.update(H_deg=d, H_coef=c, H_exp=e)
    print("gradient terms per partial:", np.diff(ptr), "H terms:", len(d))

    def run(y0, times, order, c_om):
        return _ExtendedSymplectic(order=order, c_omega_heuristic=c_om).integrate(hamsys, np.asarray(y0, float), times).states

    # test_energy_conservation
    out["energy_traj"] = run([np.pi / 2, 0, 0, 0, 0, 0], np.linspace(0, 20.0, 2000), 6, 20.0)
    # test_reversibility (forward, then a descending grid from the forward end state)
    fwd = run([0.5, 0, 0, 0.3, 0, 0], np.linspace(0, 1.5, 150), 4, 5.0)
    out["rev_fwd"] = fwd
    out["rev_bwd"] = run(fwd[-1].copy(), np.linspace(1.5, 0, 150), 4, 5.0)
    # test_final_state_error
    out["fse_200"] = run([np.pi / 4, 0, 0, 0, 0, 0], np.linspace(0, np.pi, 200), 6, 5.0)
    out["fse_800"] = run([np.pi / 4, 0, 0, 0, 0, 0], np.linspace(0, np.pi, 800), 6, 5.0)
    # test_vs_solve_ivp
    out["ivp_traj"] = run([0.1, 0, 0, 0, 0, 0], np.linspace(0, 100.0, 10000), 6, 20.0)
    # ---- the same fixture through the RK classes: integrators/_tests/test_rk.py:96-262 (the `_ham` kernels) ----
    def rk(make, y0, times):
        sol = make().integrate(hamsys, np.asarray(y0, float), times)
        return sol.states

    out["rk_energy_traj"] = rk(lambda: RungeKutta(order=8), [np.pi / 6, 0, 0, 0, 0, 0], np.linspace(0, 10.0, 10000))
    f = rk(lambda: RungeKutta(order=8), [0.3, 0, 0, 0.2, 0, 0], np.linspace(0, 1.0, 1000))
    out["rk_rev_fwd"] = f
    out["rk_rev_bwd"] = rk(lambda: RungeKutta(order=8), f[-1].copy(), np.linspace(1.0, 0, 1000))
    out["rk_fse_100"] = rk(lambda: RungeKutta(order=6), [np.pi / 4, 0, 0, 0, 0, 0], np.linspace(0, np.pi / 2, 100))
    out["rk_fse_1600"] = rk(lambda: RungeKutta(order=6), [np.pi / 4, 0, 0, 0, 0, 0], np.linspace(0, np.pi / 2, 1600))
    t_eval = np.linspace(0.0, 20.0, 4000)
    for label, make in (("4", lambda: RungeKutta(order=4)), ("6", lambda: RungeKutta(order=6)),
                        ("8", lambda: RungeKutta(order=8)), ("45", lambda: AdaptiveRK(order=5)),
                        ("853", lambda: AdaptiveRK(order=8))):
        out["rk_ivp_" + label] = rk(make, [0.1, 0, 0, 0, 0, 0], t_eval)
    a = AdaptiveRK(order=8)
    out["adaptive_defaults"] = np.array([a._rtol, a._atol, a._max_step, a._min_step])
    print("AdaptiveRK defaults rtol, atol, max_step, min_step:", out["adaptive_defaults"])
    path = os.path.join(os.path.dirname(__file__), "pendulum.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
