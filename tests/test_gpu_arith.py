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


def test_pow_squares_match_python_float_pow():
    """`x ** 2` on Python floats (the weights of the cubic detector's hit state, backend.py:632-635) is libm pow(x, 2.0):
    the device restatement must give the same bits for arguments in (0, 1], tiny ones included."""
    import math
    import torch
    from hiten_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(9)
    n = 300_000
    x = np.concatenate([rng.uniform(0.0, 1.0, n // 2), 10.0 ** rng.uniform(-17, 0, n // 4),
                        1.0 - 10.0 ** rng.uniform(-17, -1, n - n // 2 - n // 4)])
    x = x[(x > 0.0) & (x < 1.0)]
    n = len(x)
    y = np.full(n, 2.0)
    ty, tx = torch.from_numpy(y).cuda(), torch.from_numpy(x).cuda()
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(5)]
    L.check(lib.hb_selftest_arith(ty.data_ptr(), tx.data_ptr(), n, *[o.data_ptr() for o in outs],
                                  L.vp(torch.cuda.current_stream().cuda_stream)), "hb_selftest_arith")
    got = outs[4].cpu().numpy()
    ref = np.array([v ** 2 for v in x.tolist()])
    assert np.array_equal(ref, np.array([math.pow(v, 2.0) for v in x.tolist()]))
    print(f"[parity] pow(x, 2) vs libm: exact {np.mean(got == ref):.6f}; libm == x*x on {np.mean(ref == x * x):.6f}")
    assert np.array_equal(got, ref)
