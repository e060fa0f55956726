"""GPU: RK45 and fixed-step RK4 / RK6 / RK8 on the 42-state (state + STM) system -- hb_cr3bp_stm / hb_cr3bp_stm_dense with
integ.method != DOP853 -- against the reference's _compute_stm(method=..., order=...) (tests/golden/stm_variants.npz) within
the 42-state tolerance (STMs <= 1e-8 of |Phi|, BASELINE.json; measured ~1e-11), and against the oracle on a batch."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "stm_variants.npz"))
F = np.load(os.path.join(HERE, "golden", "stm_family.npz"))
CASES = [str(c) for c in G["case_names"]]


def hb_method(order):
    from hiten_b200 import _lib as L
    return {5: L.HB_RK45, 4: L.HB_RK4, 6: L.HB_RK6, 8: L.HB_RK8}[order]


def case(name):
    kind, order, steps, fwd, frac = G[f"case_{name}"]
    return int(kind), int(order), int(steps), int(fwd), float(frac)


@pytest.mark.parametrize("arith", ["parity", "fast"])
@pytest.mark.parametrize("name", CASES)
def test_gpu_42_state_variants_vs_reference(name, arith):
    import hiten_b200 as hb
    kind, order, steps, fwd, frac = case(name)
    mu = float(G["mu"])
    x0 = F["x0"][[0, 60]]
    worst = 0.0
    for j, mem in enumerate((0, 60)):
        t_eval = np.linspace(0.0, frac * float(F["period"][mem]), steps)
        integ = hb.make_integ(method=hb_method(order), arith=arith, n_fixed_steps=steps - 1)
        res = hb.cr3bp_stm_dense(x0[j:j + 1], mu, t_eval, forward=fwd, flip=(36, 42), integ=integ)
        assert int(res.status[0]) == 0
        rows = res.states[0][G[f"{name}_idx"]]
        ref = G[f"{name}_m{mem}_PHI"]
        scale = np.abs(ref[:, :36]).max(axis=1, keepdims=True)
        e_phi = (np.abs(rows[:, :36] - ref[:, :36]) / scale).max()
        e_x = np.abs(rows[:, 36:] - ref[:, 36:]).max()
        assert e_phi <= 1e-8 and e_x <= 1e-9, (name, mem, e_phi, e_x)
        worst = max(worst, e_phi)
        # end-state form (hb_cr3bp_stm): the same numbers as the last dense row
        fin = hb.cr3bp_stm(x0[j:j + 1], mu, float(t_eval[-1]), forward=fwd, flip=(36, 42), integ=integ)
        last = res.states[0][-1]
        assert np.abs(fin.states[0] - last).max() <= 1e-9 * max(1.0, np.abs(last).max())
    print(f"[parity] 42-state {name} {arith}: max |dPhi|/|Phi| = {worst:.2e}")


def test_batch_of_family_members_vs_oracle():
    """All 100 family members in one launch (RK45 dense, RK8 fixed) against the CPU oracle's generic integrators."""
    import hiten_b200 as hb
    mu = float(F["mu"])
    x0 = F["x0"]
    T = float(F["period"][0])
    for order, steps in ((5, 50), (8, 201)):
        t_eval = np.linspace(0.0, 0.5 * T, steps)
        integ = hb.make_integ(method=hb_method(order), n_fixed_steps=steps - 1)
        res = hb.cr3bp_stm_dense(x0, mu, t_eval, forward=1, flip=(36, 42), integ=integ)
        assert np.all(res.status == 0)
        s = O.system(O.SYS_VAR42, mu, fwd=1, flip=(36, 42))
        for i in range(0, 100, 7):
            y0 = np.concatenate([np.eye(6).ravel(), x0[i]])
            if order == 5:
                d, cnt = O.adaptive_dense(s, O.RK45, O.default_tol(), y0, t_eval)
                assert abs(int(res.n_acc[i]) - int(cnt[0])) <= 2
            else:
                d = O.fixed_dense(s, O.RK8, y0, t_eval)
            scale = np.abs(d[:, :36]).max(axis=1, keepdims=True)
            assert (np.abs(res.states[i][:, :36] - d[:, :36]) / scale).max() <= 1e-9
            assert np.abs(res.states[i][:, 36:] - d[:, 36:]).max() <= 1e-10
