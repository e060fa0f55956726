"""GPU connection search (hb_connections) against _ConnectionsBackend.run outputs and the oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("case", ["a", "b"])
def test_connections_bit_exact_vs_reference(case):
    from hiten_b200 import connections as cn
    g = np.load(os.path.join(HERE, "golden", "connections.npz"))
    r = cn.find_connections(g[f"{case}_pu"], g[f"{case}_ps"], g[f"{case}_Xu"], g[f"{case}_Xs"], float(g[f"{case}_eps"]),
                            float(g[f"{case}_dv_tol"]), float(g[f"{case}_bal_tol"]), traj_indices_u=g[f"{case}_tu"],
                            traj_indices_s=g[f"{case}_ts"])
    assert r.pairs_considered == int(g[f"{case}_pairs_considered"])
    assert np.array_equal(r.index_u, g[f"{case}_iu"]) and np.array_equal(r.index_s, g[f"{case}_is"])
    assert np.array_equal(r.kind, g[f"{case}_kind"])
    assert np.array_equal(r.delta_v, g[f"{case}_dv"])
    assert np.array_equal(r.point2d, g[f"{case}_pt"])
    assert np.array_equal(r.state_u, g[f"{case}_su"]) and np.array_equal(r.state_s, g[f"{case}_ss"])
    assert np.array_equal(r.trajectory_index_u, g[f"{case}_tiu"]) and np.array_equal(r.trajectory_index_s, g[f"{case}_tis"])


def test_connections_medium_batch_vs_oracle_and_edge_cases():
    from hiten_b200 import connections as cn
    rng = np.random.default_rng(3)
    n, m, eps = 20000, 17000, 1.5e-3
    pu = rng.uniform(-0.4, 0.4, (n, 2))
    ps = np.vstack((pu[rng.choice(n, m // 2, replace=False)] + rng.uniform(-1, 1, (m // 2, 2)) * eps * 0.8,
                    rng.uniform(-0.4, 0.4, (m - m // 2, 2))))
    Xu, Xs = rng.normal(0, 0.2, (n, 6)), rng.normal(0, 0.2, (m, 6))
    Xs[: m // 2, 3:] = Xu[:m // 2, 3:] * (1 + 1e-3 * rng.normal(size=(m // 2, 3)))
    r = cn.find_connections(pu, ps, Xu, Xs, eps, 0.6, 1e-3)
    o = O.connections(pu, ps, Xu, Xs, eps, 0.6, 1e-3)
    assert r.pairs_considered == o["pairs_considered"] and len(r.delta_v) == len(o["dv"]) > 1000
    assert np.array_equal(r.index_u, o["iu"]) and np.array_equal(r.index_s, o["is_"])
    assert np.array_equal(r.delta_v, o["dv"]) and np.array_equal(r.point2d, o["pt"])
    assert np.array_equal(r.state_u, o["su"]) and np.array_equal(r.state_s, o["ss"])
    # nothing within eps / empty inputs / single points
    assert len(cn.find_connections(pu[:100], ps[:100] + 10.0, Xu[:100], Xs[:100], eps, 1.0, 1e-3).delta_v) == 0
    assert len(cn.find_connections(pu[:0], ps[:10], Xu[:0], Xs[:10], eps, 1.0, 1e-3).delta_v) == 0
    one = cn.find_connections(pu[:1], pu[:1] + 1e-5, Xu[:1], Xu[:1], eps, 1.0, 1e-3)
    assert len(one.delta_v) == 1 and one.delta_v[0] == 0.0 and one.kind[0] == 0
