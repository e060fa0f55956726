"""GPU: hiten_b200.sharded.ShardedTubeSection -- one process driving several devices (per-device runners, workspaces,
streams; every kernel attribute and the NVRTC cache are per device).  With one visible device the two shards run on
the same GPU (the sharding / merging logic is the same); with more they run on distinct devices."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("kw", [dict(steps_capacity=192), dict(steps_capacity=0, pool_records=8)])
def test_sharded_equals_single_device(kw):
    import torch
    from hiten_b200 import sharded, synodic
    g = np.load(os.path.join(HERE, "golden", "synodic_c2.npz"))
    mu, fwd = float(g["mu"]), int(g["forward"])
    t_eval = np.linspace(0.0, float(g["tf"]), int(g["steps"]))
    sec = synodic.make_section("y", float(g["req_offset"]), ("x", "z"), int(g["req_direction"]))
    x0 = g["x0W"][:199]                                             # odd: the shards differ in size
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    sh = sharded.ShardedTubeSection(len(x0), mu, t_eval, sec, forward=fwd, flip=(0, 6), devices=devices, **kw)
    sh.launch(x0)
    hits, yf, status = sh.gather()
    assert (status == 0).all()
    sel = g["hit_traj"] < 199
    assert np.array_equal(hits.trajectory_indices, g["hit_traj"][sel])
    assert np.array_equal(hits.times, g["hit_time"][sel]) and np.array_equal(hits.states, g["hit_state"][sel])
    assert np.array_equal(yf, g["yf"][:199])
    assert int(hits.hits_per_traj.sum()) == int(sel.sum())


def test_second_device_runs_every_kernel_family():
    """The per-device state an earlier version kept per process (opt-in shared-memory attributes, the NVRTC function
    cache): the record pipeline, the filter kernel and the specialised CM map on the LAST visible device after the
    first one has used them."""
    import torch
    from hiten_b200 import centermanifold as cm
    from hiten_b200 import synodic
    ndev = torch.cuda.device_count()
    g = np.load(os.path.join(HERE, "golden", "synodic_c1.npz"))
    t_eval = np.linspace(0.0, float(g["tf"]), int(g["steps"]))
    sec = synodic.make_section("y", 0.0, ("x", "z"), -1)
    gm = np.load(os.path.join(HERE, "golden", "cm_map.npz"))
    tab = cm.PolyTable(gm["jac_ptr"], gm["jac_deg"], gm["jac_coef"], gm["jac_exp"])
    opts = cm.make_opts(0.01, 2000, "symplectic", 4, "p3", 20.0, "parity")
    out = []
    for d in (0, ndev - 1):
        dev = torch.device("cuda", d)
        with torch.cuda.device(dev):
            r = synodic.TubeSectionRunner(50, float(g["mu"]), t_eval, sec, forward=-1, flip=(0, 6), steps_capacity=160,
                                          filters=(3.318e-05, 9.04e-06, 1e-7), device=dev)
            r.launch(torch.from_numpy(np.ascontiguousarray(g["x0W"].T)).to(dev))
            h = r.sorted_hits()
            f, st, tt = cm.poincare_map(tab, torch.from_numpy(gm["seeds_p3"][:64]).to(dev), opts, device=dev)
            out.append((h.times.copy(), r.filter_result()[0].cpu().numpy(), st.cpu().numpy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1], equal_nan=True)
    assert np.array_equal(out[0][2], out[1][2])
    ref = gm["tao4_p3"][:64]
    assert np.array_equal(out[1][2][ref[:, 0] == 1], ref[ref[:, 0] == 1][:, 1:5])
