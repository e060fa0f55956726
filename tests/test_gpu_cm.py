"""GPU parity: centre-manifold Poincare map kernel vs the reference's _poincare_map (BASELINE config 3)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from test_oracle_cm import CASES

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _table(g):
    from hiten_b200.centermanifold import PolyTable
    return PolyTable(g["jac_ptr"], g["jac_deg"], g["jac_coef"], g["jac_exp"])


@pytest.mark.parametrize("jit", [True, False], ids=["specialised", "table"])
@pytest.mark.parametrize("name,order,symp,sec", CASES)
def test_map_vs_reference(name, order, symp, sec, jit):
    from hiten_b200 import centermanifold as cmod
    g = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    ref = g[name]
    seeds = g["seeds_" + sec][: len(ref)]
    opts = cmod.make_opts(float(g["dt"]), int(g["max_steps"]), "symplectic" if symp else "fixed", order, sec,
                          float(g["c_omega"]))
    f, o, t = cmod.poincare_map(_table(g), seeds, opts, jit=jit)
    assert np.array_equal(f, ref[:, 0].astype(np.int64))              # identical crossing flags
    d = np.abs(o - ref[:, 1:5]).max()
    print(f"[parity] CM {name}: {len(ref)} seeds, |d state| max {d:.2e}, |d t| max {np.abs(t - ref[:, 5]).max():.2e}, "
          f"bit-exact {np.array_equal(o, ref[:, 1:5])}")
    assert np.array_equal(o, ref[:, 1:5]) and np.array_equal(t, ref[:, 5])      # bit-exact


def test_fast_variant_and_failures():
    from hiten_b200 import centermanifold as cmod
    g = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    tab = _table(g)
    ref = g["rk4_p3_maxsteps200"]
    f, o, t = cmod.poincare_map(tab, g["seeds_p3"][:64], cmod.make_opts(0.01, 200, "fixed", 4, "p3"))
    assert np.array_equal(f, ref[:, 0].astype(np.int64)) and (o[f == 0] == 0).all()
    ref = g["tao4_p3"]
    f, o, t = cmod.poincare_map(tab, g["seeds_p3"][:256], cmod.make_opts(0.01, 2000, "symplectic", 4, "p3", arith="fast"))
    assert np.array_equal(f, ref[:, 0].astype(np.int64))
    assert np.abs(o - ref[:, 1:5]).max() <= 1e-9
    with pytest.raises(NotImplementedError):
        cmod.make_opts(method="adaptive")
    e = cmod.poincare_map(tab, np.empty((0, 4)), cmod.make_opts())
    assert e[0].size == 0 and e[1].shape == (0, 4)


def test_large_batch_returns_on_section():
    """1e5 seeds (BASELINE size): every returned point lies on the section and conserves energy."""
    from hiten_b200 import centermanifold as cmod
    g = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    tab = _table(g)
    rng = np.random.default_rng(1)
    base = g["seeds_p3"]
    seeds = base[rng.integers(0, len(base), 100_000)]
    f, o, t = cmod.poincare_map(tab, seeds, cmod.make_opts(0.01, 2000, "symplectic", 4, "p3"))
    assert f.all()
    # p3 ~ 0 on the section: the reference's hit is a Hermite interpolant at the LINEAR alpha, so it is only
    # near the section (the golden hits reach |p3| = 1.2e-5 as well)
    assert np.abs(o[:, 3]).max() < 1e-4
    # identical seeds -> identical results regardless of lane / queue order
    idx = np.where((seeds == base[0]).all(axis=1))[0]
    assert len(idx) > 1 and (o[idx] == o[idx[0]]).all()


def test_iterated_map_on_device_matches_the_engine_loop():
    """poincare_map_iterate = the _worker loop of _CenterManifoldEngine.solve (engine.py:163-191): five iterations of
    the Tao-4 map with the hits fed back as seeds, against the same loop over the oracle (bit-exact vs the reference)."""
    from hiten_b200 import centermanifold as cmod
    g = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    tab = _table(g)
    sec = "p3"
    seeds = g["seeds_" + sec][:256]
    opts = cmod.make_opts(float(g["dt"]), int(g["max_steps"]), "symplectic", 4, sec, float(g["c_omega"]))
    st, tt, it = cmod.poincare_map_iterate(tab, seeds, opts, 5, sec)
    ham = O.PolyHam(tab.ptr, tab.deg, tab.coef, tab.exp)
    cur, rs, rt, ri = seeds, [], [], []
    for k in range(5):
        f, o, t = O.cm_poincare_map(ham, cur, float(g["dt"]), 4, int(g["max_steps"]), True, sec, float(g["c_omega"]), 4)
        ok = f == 1
        nxt = o[ok].copy()
        if not len(nxt):
            break
        nxt[:, 3] = 0.0
        rs.append(nxt); rt.append(t[ok]); ri.append(np.full(len(nxt), k))
        cur = nxt
    assert np.array_equal(st.cpu().numpy(), np.vstack(rs))
    assert np.array_equal(tt.cpu().numpy(), np.concatenate(rt))
    assert np.array_equal(it.cpu().numpy(), np.concatenate(ri))
    assert len(rs) == 5 and st.shape[0] > 1000


@pytest.mark.parametrize("order,symp", [(4, True), (4, False), (8, False)])
def test_split_and_plain_kernels_hand_rounds_over_bit_exactly(order, symp):
    """40000 DISTINCT seeds: round 0 runs in cm_map (the list is longer than the split threshold), the later rounds -- the
    seeds that need more than 160 steps -- in cm_map_split (four warps per 32 seeds); 3000 seeds run in cm_map_split from
    the start.  Both against the table-driven kernel (one thread per seed, no rounds): identical flags, states and times."""
    from hiten_b200 import centermanifold as cmod
    g = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    tab = _table(g)
    rng = np.random.default_rng(7)
    base = g["seeds_p3"]
    for n in (40_000, 3_000):
        seeds = base[rng.integers(0, len(base), n)] * (1.0 - 0.5 * rng.random((n, 1)))
        opts = cmod.make_opts(0.01, 2000, "symplectic" if symp else "fixed", order, "p3", 20.0)
        f1, o1, t1 = cmod.poincare_map(tab, seeds, opts, jit=True)
        f0, o0, t0 = cmod.poincare_map(tab, seeds, opts, jit=False)
        assert f1.sum() > 0.99 * n and (t1 / 0.01).max() > 400          # several rounds were needed
        assert np.array_equal(f1, f0) and np.array_equal(o1, o0) and np.array_equal(t1, t0)
