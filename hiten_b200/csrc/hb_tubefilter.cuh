// hb_tubefilter.cuh -- per-sample quantities of the manifold trajectory filters, shared by hb_tube_filter (stored
// tubes, hb_manifold.cu) and hb_section2_filter (step records, hb_section_scan.cu) so that both give the same bits.
//
// Reference: safe-radius expressions of _run_compute (hiten/algorithms/types/services/manifold.py:412-424) and
// _max_rel_energy_error (hiten/algorithms/common/energy.py:27-76).  IEEE division / square root (the compiler's
// div.rn / sqrt.rn with their special-case paths: samples may be NaN or sit on a primary).
#pragma once
#include <math_constants.h>

#include "hb_common.cuh"

// Jacobi constant (energy.py:51-54) and, as a by-product, the squared distance to the primary that the safe-radius
// test uses too ((x + mu)^2 + y^2 + z^2 is the same expression in both places)
HB_DEV double jacobi_ref(const double *s, double mu1, double mu2, double &s1)
{
    const double a = __dadd_rn(s[0], mu2), b = __dsub_rn(s[0], mu1);
    const double yy = __dmul_rn(s[1], s[1]), zz = __dmul_rn(s[2], s[2]);
    s1 = __dadd_rn(__dadd_rn(__dmul_rn(a, a), yy), zz);
    const double r1 = __dsqrt_rn(s1);
    const double r2 = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(b, b), yy), zz));
    const double pot = __dmul_rn(2.0, __dadd_rn(__ddiv_rn(mu1, r1), __ddiv_rn(mu2, r2)));
    const double kin = __dadd_rn(__dadd_rn(__dmul_rn(s[3], s[3]), __dmul_rn(s[4], s[4])), __dmul_rn(s[5], s[5]));
    return __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(s[0], s[0]), yy), pot), kin);
}
HB_DEV double jacobi_ref(const double *s, double mu1, double mu2)
{
    double s1;
    return jacobi_ref(s, mu1, mu2, s1);
}

struct NanMin {
    double v;
    bool nan;
    HB_DEV void take(double x)
    {
        if (x != x) nan = true;
        else if (x < v) v = x;
    }
    HB_DEV void warp_reduce()
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            if (ov < v) v = ov;
        }
        nan = __any_sync(0xffffffffu, nan);
    }
    HB_DEV double result() const { return nan ? CUDART_NAN : v; }
};

// Running reductions over the samples one lane sees.  Correctly rounded sqrt and division by a positive constant are
// monotone, so min_k sqrt(s_k) = sqrt(min_k s_k) and max_k (d_k / |C0|) = (max_k d_k) / |C0| bit for bit: the
// accumulators hold the squared distances and the absolute drift, and the two square roots and the division are taken
// once per trajectory instead of once per sample.
struct TubeFilterAcc {
    NanMin m1, m2;      // min of the SQUARED distances to the primaries (manifold.py:415-416 before the sqrt)
    double mx;          // max |C_k - C_0|
    HB_DEV TubeFilterAcc()
    {
        m1.v = m2.v = CUDART_INF;
        m1.nan = m2.nan = false;
        mx = 0.0;
    }
    // k = sample index (sample 0 defines C0 and is not compared, energy.py:62)
    HB_DEV void sample(const double *s, int k, double mu, double mu1, double mu2, double C0)
    {
        double s1;
        const double Ck = jacobi_ref(s, mu1, mu2, s1);       // s1 = (x + mu)**2 + y**2 + z**2
        const double b = __dadd_rn(__dsub_rn(s[0], 1.0), mu);   // (x - 1 + mu): not the Jacobi routine's x - (1 - mu)
        m1.take(s1);
        m2.take(__dadd_rn(__dadd_rn(__dmul_rn(b, b), __dmul_rn(s[1], s[1])), __dmul_rn(s[2], s[2])));
        if (k > 0) {                                         // energy.py:62-73
            const double dC = fabs(__dsub_rn(Ck, C0));
            if (dC > mx) mx = dC;
        }
        (void)mu;
    }
    HB_DEV void warp_reduce()
    {
        m1.warp_reduce();
        m2.warp_reduce();
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, mx, sh);
            if (ov > mx) mx = ov;
        }
    }
    // lane 0: write {min r1, min r2, max drift} and the keep / discard decision of _run_compute
    HB_DEV void store(const hb_tube_filter_opts &o, long long traj, double absC0, double *out, int *keep) const
    {
        const double r1 = __dsqrt_rn(m1.result()), r2 = __dsqrt_rn(m2.result());
        const double drift = absC0 > 1e-14 ? __ddiv_rn(mx, absC0) : mx;
        out[3 * traj + 0] = r1;
        out[3 * traj + 1] = r2;
        out[3 * traj + 2] = drift;
        if (keep) keep[traj] = !((r1 < o.safe_r1) || (r2 < o.safe_r2)) && !(drift > o.energy_tol);
    }
};
