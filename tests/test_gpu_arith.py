"""GPU: the restated div.rn / sqrt.rn fast paths of the parity variant are IEEE-correctly rounded."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_shared_reciprocal_division_and_sqrt_are_correctly_rounded():
    import torch
    from hiten_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(7)
    n = 4_000_000
    # magnitudes the propagation produces: numerators 1e-30..1e6 (and exact zeros), denominators 1e-14..1e8
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 6, n)
    a[::1000] = 0.0
    b = np.abs(rng.standard_normal(n)) * 10.0 ** rng.uniform(-14, 8, n) + 1e-300
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(5)]
    rc = lib.hb_selftest_arith(ta.data_ptr(), tb.data_ptr(), n, *[o.data_ptr() for o in outs],
                               L.vp(torch.cuda.current_stream().cuda_stream))
    L.check(rc, "hb_selftest_arith")
    torch.cuda.synchronize()
    div_shared, div_ref, sqrt_fast, sqrt_ref, _ = [o.cpu().numpy() for o in outs]
    assert np.array_equal(div_ref, a / b)                    # the intrinsic is IEEE (sanity)
    assert np.array_equal(div_shared, a / b)                 # ours too, bit for bit
    assert np.array_equal(sqrt_ref, np.sqrt(b))
    assert np.array_equal(sqrt_fast, np.sqrt(b))


def test_pow_matches_libm_bit_for_bit():
    """Controller exponents: the device restatement of glibc pow() vs the host libm the reference calls."""
    import torch
    from hiten_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(8)
    n = 300_000
    x = 10.0 ** rng.uniform(-6, 3, n)
    y = rng.choice(np.array([-1.0 / 9.0, 0.4 * (1.0 / 9.0), -1.0 / 8.0, -1.0 / 6.0, 0.4 * (1.0 / 6.0), -0.2, 1.5, 2.5]), n)
    ty, tx = torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda()
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(5)]
    rc = lib.hb_selftest_arith(ty.data_ptr(), tx.data_ptr(), n, *[o.data_ptr() for o in outs],
                               L.vp(torch.cuda.current_stream().cuda_stream))
    L.check(rc, "hb_selftest_arith")
    got = outs[4].cpu().numpy()
    import math
    ref = np.array([math.pow(a, b) for a, b in zip(x.tolist(), y.tolist())])   # glibc pow, what Numba calls
    # (np.power on arrays is NumPy's own SIMD kernel and differs from glibc in ~5 % of arguments)
    ulp = np.abs(got - ref) / np.spacing(ref)
    print(f"[parity] pow vs libm: exact {np.mean(ulp == 0):.6f}, max {ulp.max():.1f} ulp")
    assert np.array_equal(got, ref)
