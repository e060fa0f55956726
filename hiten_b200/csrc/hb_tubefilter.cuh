// hb_tubefilter.cuh -- per-sample quantities of the manifold trajectory filters, shared by hb_tube_filter (stored
// tubes, hb_manifold.cu) and hb_section2_filter (step records, hb_section_scan.cu) so that both give the same bits.
//
// Reference: safe-radius expressions of _run_compute (hiten/algorithms/types/services/manifold.py:412-424) and
// _max_rel_energy_error (hiten/algorithms/common/energy.py:27-76).  IEEE division / square root (the compiler's
// div.rn / sqrt.rn with their special-case paths: samples may be NaN or sit on a primary).
#pragma once
#include <math_constants.h>

#include "hb_common.cuh"

HB_DEV double jacobi_ref(const double *s, double mu1, double mu2)
{
    const double a = __dadd_rn(s[0], mu2), b = __dsub_rn(s[0], mu1);
    const double yy = __dmul_rn(s[1], s[1]), zz = __dmul_rn(s[2], s[2]);
    const double r1 = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, a), yy), zz));
    const double r2 = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(b, b), yy), zz));
    const double pot = __dmul_rn(2.0, __dadd_rn(__ddiv_rn(mu1, r1), __ddiv_rn(mu2, r2)));
    const double kin = __dadd_rn(__dadd_rn(__dmul_rn(s[3], s[3]), __dmul_rn(s[4], s[4])), __dmul_rn(s[5], s[5]));
    return __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(s[0], s[0]), yy), pot), kin);
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

// running reductions over the samples one lane sees
struct TubeFilterAcc {
    NanMin m1, m2;
    double mx;
    HB_DEV TubeFilterAcc()
    {
        m1.v = m2.v = CUDART_INF;
        m1.nan = m2.nan = false;
        mx = 0.0;
    }
    // k = sample index (sample 0 defines C0 and is not compared, energy.py:62)
    HB_DEV void sample(const double *s, int k, double mu, double mu1, double mu2, double C0, double absC0)
    {
        // manifold.py:415-416: np.sqrt((x + mu)**2 + y**2 + z**2), np.sqrt((x - 1 + mu)**2 + y**2 + z**2)
        const double a = __dadd_rn(s[0], mu), b = __dadd_rn(__dsub_rn(s[0], 1.0), mu);
        const double yy = __dmul_rn(s[1], s[1]), zz = __dmul_rn(s[2], s[2]);
        m1.take(__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, a), yy), zz)));
        m2.take(__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(b, b), yy), zz)));
        if (k > 0) {                                         // energy.py:62-73
            const double dC = fabs(__dsub_rn(jacobi_ref(s, mu1, mu2), C0));
            const double rel = absC0 > 1e-14 ? __ddiv_rn(dC, absC0) : dC;
            if (rel > mx) mx = rel;
        }
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
    HB_DEV void store(const hb_tube_filter_opts &o, long long traj, double *out, int *keep) const
    {
        const double r1 = m1.result(), r2 = m2.result();
        out[3 * traj + 0] = r1;
        out[3 * traj + 1] = r2;
        out[3 * traj + 2] = mx;
        if (keep) keep[traj] = !((r1 < o.safe_r1) || (r2 < o.safe_r2)) && !(mx > o.energy_tol);
    }
};
