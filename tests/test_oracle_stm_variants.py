"""CPU: the oracle's RK45 and fixed-step RK4 / RK6 / RK8 on the 42-state (state + STM) system against the reference's
_compute_stm(method=..., order=...) (tests/golden/stm_variants.npz, make_stm_variants.py).  Tolerance statement, like
every 42-state comparison: the reference's Jacobian uses libm pow and a SIMD dot product (test_oracle_stm.py)."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "stm_variants.npz"))
F = np.load(os.path.join(HERE, "golden", "stm_family.npz"))
CASES = [str(c) for c in G["case_names"]]
METHOD = {5: O.RK45, 4: O.RK4, 6: O.RK6, 8: O.RK8}


def case(name):
    kind, order, steps, fwd, frac = G[f"case_{name}"]
    return int(kind), int(order), int(steps), int(fwd), float(frac)


def check(name, mem, rows, tol_phi, tol_x):
    ref = G[f"{name}_m{mem}_PHI"]
    scale = np.abs(ref[:, :36]).max(axis=1, keepdims=True)
    e_phi = (np.abs(rows[:, :36] - ref[:, :36]) / scale).max()
    e_x = np.abs(rows[:, 36:] - ref[:, 36:]).max()
    assert e_phi <= tol_phi and e_x <= tol_x, (name, mem, e_phi, e_x)
    return e_phi, e_x


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mem", [0, 60])
def test_oracle_42_state_variants_vs_reference(name, mem):
    kind, order, steps, fwd, frac = case(name)
    s = O.system(O.SYS_VAR42, float(G["mu"]), fwd=fwd, flip=(36, 42))
    y0 = np.concatenate([np.eye(6).ravel(), F["x0"][mem]])
    t_eval = np.linspace(0.0, frac * float(F["period"][mem]), steps)
    if kind == 0:
        d, _ = O.adaptive_dense(s, METHOD[order], O.default_tol(), y0, t_eval)
    else:
        d = O.fixed_dense(s, METHOD[order], y0, t_eval)
    assert G[f"{name}_m{mem}_tlast"] == fwd * t_eval[-1]
    check(name, mem, d[G[f"{name}_idx"]], 1e-9, 1e-10)
