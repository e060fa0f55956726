/* ho_manifold.c -- CPU restatement of the manifold-tube host steps around the propagation loop
 * (TEST INFRASTRUCTURE, see hiten_oracle.h).
 *
 * Reference (paths relative to src/hiten/):
 *   _ManifoldDynamicsService._totime                    algorithms/types/services/manifold.py:539-573
 *   _ManifoldDynamicsService._compute_manifold_section  algorithms/types/services/manifold.py:470-537
 *   safe-radius filter of _run_compute                  algorithms/types/services/manifold.py:412-424
 *   _max_rel_energy_error                               algorithms/common/energy.py:27-76
 *
 * Platform arithmetic that had to be reproduced (measured against the reference run in the build container,
 * tests/golden/make_manifold_ics.py; 668/668 initial conditions bit-identical):
 *   - `phi_frac @ eigvec` multiplies a float64 6x6 by a COMPLEX128 vector (get_real_eigenvectors returns complex
 *     arrays with zero imaginary part), i.e. numpy promotes to complex and calls OpenBLAS zgemv (transposed,
 *     Haswell kernel): elements 0..3 of each row run through two FMA lanes (even / odd elements), the lanes are
 *     added, elements 4..5 are a separately rounded mul + add tail, and head + tail is the result.  The imaginary
 *     products are exact zeros.
 *   - np.linalg.norm of the complex 3-vector = sqrt(re.dot(re) + im.dot(im)); the 3-element dot is an
 *     FMA-accumulated sequential sum (OpenBLAS ddot tail).
 *   - everything else is numpy elementwise arithmetic (one rounding per operation, left to right).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "hiten_oracle.h"

/* _totime (manifold.py:566-571): index of the first minimum of |target - |t[k]||. */
int64_t ho_totime(const double *tt, int64_t n_samples, double target)
{
    int64_t best = 0;
    double bd = fabs(target - fabs(tt[0]));
    for (int64_t k = 1; k < n_samples; ++k) {
        double d = fabs(target - fabs(tt[k]));
        if (d < bd) { bd = d; best = k; }          /* np.argmin: first occurrence; NaN never enters here */
    }
    return best;
}

/* One row of `phi_frac @ eigvec` (real part), OpenBLAS zgemv_t order (see header). */
static double zgemv_row6(const double *a, const double *x)
{
    double l0 = fma(a[0], x[0], 0.0);
    double l1 = fma(a[1], x[1], 0.0);
    l0 = fma(a[2], x[2], l0);
    l1 = fma(a[3], x[3], l1);
    double head = l0 + l1;
    double tail = 0.0 + a[4] * x[4];
    tail = tail + a[5] * x[5];
    return head + tail;
}

/* _compute_manifold_section for one (sample index, displacement): x0W[6]. */
void ho_manifold_section(const double *phi_row42, const double *eigvec_re, int direction, double displacement,
                         double *x0w)
{
    double man[6];
    for (int r = 0; r < 6; ++r) man[r] = (double)direction * zgemv_row6(phi_row42 + 6 * r, eigvec_re);
    double sq = 0.0;
    for (int c = 0; c < 3; ++c) sq = fma(man[c], man[c], sq);
    sq = sq + 0.0;                                   /* + x_imag.dot(x_imag) */
    double mag = sqrt(sq);
    if (mag < 1e-14) mag = 1.0;                      /* manifold.py:516-521 */
    double d = displacement / mag;
    for (int c = 0; c < 6; ++c) x0w[c] = phi_row42[36 + c] + d * man[c];
    if (fabs(x0w[2]) < 1.0e-15) x0w[2] = 0.0;        /* manifold.py:531-534 */
    if (fabs(x0w[5]) < 1.0e-15) x0w[5] = 0.0;
}

/* All initial conditions of a tube: fractions[K] x displacements[D], displacement-major rows
 * x0w[(j*K + k)][6]; node_idx[K] (optional) receives the STM sample index of each fraction.
 * phi_dense[S][42] is the reference's PHI (state in columns 36..41 = xx), tt[S] its times. */
void ho_manifold_ics(const double *phi_dense, const double *tt, int64_t n_samples, double period,
                     const double *eigvec_re, int direction, const double *fractions, int64_t K,
                     const double *displacements, int64_t D, double *x0w, int64_t *node_idx)
{
    for (int64_t k = 0; k < K; ++k) {
        int64_t idx = ho_totime(tt, n_samples, fractions[k] * period);
        if (node_idx) node_idx[k] = idx;
        for (int64_t j = 0; j < D; ++j)
            ho_manifold_section(phi_dense + 42 * idx, eigvec_re, direction, displacements[j], x0w + 6 * (j * K + k));
    }
}

static double jacobi(const double *s, double mu1, double mu2)
{
    double x = s[0], y = s[1], z = s[2], vx = s[3], vy = s[4], vz = s[5];
    double a = x + mu2, b = x - mu1;
    double r1 = sqrt(a * a + y * y + z * z);        /* (...) ** 0.5: LLVM lowers pow(., 0.5) to sqrt */
    double r2 = sqrt(b * b + y * y + z * z);
    return x * x + y * y + 2.0 * (mu1 / r1 + mu2 / r2) - (vx * vx + vy * vy + vz * vz);
}

/* Per-trajectory filter quantities of _run_compute on a stored tube states[m][6]:
 *   out[0] = r1.min(), out[1] = r2.min()   with the numpy expressions of manifold.py:415-416
 *   out[2] = _max_rel_energy_error(states, mu)                                            */
void ho_tube_filter(const double *states, int m, double mu, double *out)
{
    double mn1 = INFINITY, mn2 = INFINITY;
    for (int i = 0; i < m; ++i) {
        const double *s = states + 6 * (int64_t)i;
        double a = s[0] + mu, b = s[0] - 1 + mu;
        double r1 = sqrt(a * a + s[1] * s[1] + s[2] * s[2]);
        double r2 = sqrt(b * b + s[1] * s[1] + s[2] * s[2]);
        if (r1 < mn1 || isnan(r1)) mn1 = isnan(mn1) ? mn1 : r1;   /* np.min propagates NaN */
        if (r2 < mn2 || isnan(r2)) mn2 = isnan(mn2) ? mn2 : r2;
    }
    double mu1 = 1.0 - mu, mu2 = mu;
    double C0 = jacobi(states, mu1, mu2), absC0 = fabs(C0), mx = 0.0;
    for (int i = 1; i < m; ++i) {
        double Ci = jacobi(states + 6 * (int64_t)i, mu1, mu2);
        double rel = absC0 > 1e-14 ? fabs(Ci - C0) / absC0 : fabs(Ci - C0);
        if (rel > mx) mx = rel;
    }
    out[0] = mn1;
    out[1] = mn2;
    out[2] = mx;
}

void ho_batch_tube_filter(const double *states, int64_t n, int m, double mu, double *out /* [n][3] */)
{
    for (int64_t i = 0; i < n; ++i) ho_tube_filter(states + (int64_t)i * m * 6, m, mu, out + 3 * i);
}
