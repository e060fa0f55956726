"""Layout of the configs[4] bench batches (hiten_b200/workloads.py), CPU only: bench.py's pilot cost model and
tools/sim_launch_order.py rely on trajectory index = displacement_row * 2000 + orbit_node."""
import numpy as np

from hiten_b200 import workloads as W


def test_c5_batch_is_displacement_major_node_minor():
    ics, mu = W.c5_batch(2 * 6000)
    t = W.c5_nodes()
    disp = np.logspace(-7.0, -5.0, 3)
    for key in ("l1", "l2"):
        x = ics[key]
        assert x.shape == (6000, 6)
        xn, man = t[f"{key}_x_node"], t[f"{key}_man"]
        assert xn.shape == (2000, 6)
        mag = np.linalg.norm(man[:, 0:3], axis=1)
        for row in range(3):
            block = x[row * 2000:(row + 1) * 2000]
            want = xn + (disp[row] / mag)[:, None] * man
            want[np.abs(want[:, 2]) < 1e-15, 2] = 0.0
            want[np.abs(want[:, 5]) < 1e-15, 5] = 0.0
            assert np.allclose(block, want, rtol=1e-13, atol=1e-18)      # (the norm is taken per node there: last-bit differences)
            if row:
                assert not np.allclose(block, x[(row - 1) * 2000:row * 2000], rtol=1e-9, atol=0.0)


def test_c5_batch_shards_interleave_by_index():
    whole, _ = W.c5_batch(2 * 4000)
    for rank in range(2):
        part, _ = W.c5_batch(2 * 4000, rank, 2)
        for key in ("l1", "l2"):
            assert np.array_equal(part[key], whole[key][rank::2])
