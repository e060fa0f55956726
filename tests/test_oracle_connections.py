"""Oracle connection search against _ConnectionsBackend.run outputs (tests/golden/connections.npz)."""
import os

import numpy as np
import pytest

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("case", ["a", "b"])
def test_connections_bit_exact(case):
    g = np.load(os.path.join(HERE, "golden", "connections.npz"))
    r = O.connections(g[f"{case}_pu"], g[f"{case}_ps"], g[f"{case}_Xu"], g[f"{case}_Xs"], float(g[f"{case}_eps"]),
                      float(g[f"{case}_dv_tol"]), float(g[f"{case}_bal_tol"]))
    assert r["pairs_considered"] == int(g[f"{case}_pairs_considered"])
    assert len(r["dv"]) == len(g[f"{case}_dv"]) > 100
    assert np.array_equal(r["iu"], g[f"{case}_iu"]) and np.array_equal(r["is_"], g[f"{case}_is"])
    assert np.array_equal(r["kind"], g[f"{case}_kind"]) and 0 < r["kind"].sum() < len(r["kind"])
    assert np.array_equal(r["dv"], g[f"{case}_dv"])
    assert np.array_equal(r["pt"], g[f"{case}_pt"])
    assert np.array_equal(r["su"], g[f"{case}_su"]) and np.array_equal(r["ss"], g[f"{case}_ss"])
