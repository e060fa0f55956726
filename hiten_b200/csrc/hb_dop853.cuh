// hb_dop853.cuh -- DOP853 machinery on a lane-local 6-component slice, tableau resolved at compile time.
//
// Shared by the 6-state kernel (one trajectory per thread: the slice IS the state) and the 42-state
// state+STM kernel (8-lane group per trajectory: each lane carries one STM column or the state).
// RHS is a functor  void operator()(const double (&y)[6], double (&dy)[6]) const .
// Reference: hiten/algorithms/integrators/rk.py  dop853_step_jit_kernel :1637-1708,
// _dop853_build_dense_cache :1791-1875, _dop853_eval_dense :1962-2003.
#pragma once
#include "hb_common.cuh"

// ---------------------------------------------------------------------------------------------
// DOP853 stages, tableau resolved at compile time.
//   y_stage = y; for j < i with a_ij != 0: y_stage += (h * a_ij) * k_j      (rk.py:1674-1678)
// ---------------------------------------------------------------------------------------------
template <class AR, int I, int J>
HB_DEV void stage_acc(double (&ys)[6], const double (&k)[13][6], double h)
{
    if constexpr (J < I) {
        if constexpr (HB_DOP853_A[I][J] != 0.0) {
            constexpr double a = HB_DOP853_A[I][J];
            const double ha = AR::mul(h, a);
#pragma unroll
            for (int d = 0; d < 6; ++d) ys[d] = AR::madd(ha, k[J][d], ys[d]);
        }
        stage_acc<AR, I, J + 1>(ys, k, h);
    }
}

template <class AR, class RHS, int I>
HB_DEV void run_stages(const double (&y)[6], double (&k)[13][6], double h, const RHS &rhs)
{
    if constexpr (I < 12) {
        double ys[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) ys[d] = y[d];
        stage_acc<AR, I, 0>(ys, k, h);
        rhs(ys, k[I]);
        run_stages<AR, RHS, I + 1>(y, k, h, rhs);
    }
}

template <class AR, int J>
HB_DEV void high_acc(double (&yh)[6], const double (&k)[13][6], double h)
{
    if constexpr (J < 12) {
        if constexpr (HB_DOP853_B[J] != 0.0) {
            constexpr double b = HB_DOP853_B[J];
            const double hb = AR::mul(h, b);
#pragma unroll
            for (int d = 0; d < 6; ++d) yh[d] = AR::madd(hb, k[J][d], yh[d]);
        }
        high_acc<AR, J + 1>(yh, k, h);
    }
}

// err5 += E5_j * k_j ; err3 += E3_j * k_j     (rk.py:1691-1697)
template <class AR, int J>
HB_DEV void err_acc(double (&e5)[6], double (&e3)[6], const double (&k)[13][6])
{
    if constexpr (J < 13) {
        if constexpr (HB_DOP853_E5[J] != 0.0) {
            constexpr double c = HB_DOP853_E5[J];
#pragma unroll
            for (int d = 0; d < 6; ++d) e5[d] = AR::madd(c, k[J][d], e5[d]);
        }
        if constexpr (HB_DOP853_E3[J] != 0.0) {
            constexpr double c = HB_DOP853_E3[J];
#pragma unroll
            for (int d = 0; d < 6; ++d) e3[d] = AR::madd(c, k[J][d], e3[d]);
        }
        err_acc<AR, J + 1>(e5, e3, k);
    }
}

// One attempted step, stage part.  k[0] must hold f(t, y); fills k[1..12] and y_high.
template <class AR, class RHS>
HB_DEV void dop853_stages(const double (&y)[6], double (&k)[13][6], double h, double (&yh)[6], const RHS &rhs)
{
    run_stages<AR, RHS, 1>(y, k, h, rhs);
#pragma unroll
    for (int d = 0; d < 6; ++d) yh[d] = y[d];
    high_acc<AR, 0>(yh, k, h);
    rhs(yh, k[12]);
}

// Error sums of this lane's 6 components (rk.py:1686-1699, 2457-2462):
//   n5 = dot(e5/scale, e5/scale), n3 = dot(e3/scale, e3/scale), sequential FMA accumulation like np.dot.
template <class AR>
HB_DEV void dop853_err_sums(const double (&y)[6], const double (&yh)[6], const double (&k)[13][6], double h,
                            double rtol, double atol, double &n5, double &n3)
{
    double e5[6], e3[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) { e5[d] = 0.0; e3[d] = 0.0; }
    err_acc<AR, 0>(e5, e3, k);
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        const double sc = AR::madd(rtol, fmax(fabs(y[d]), fabs(yh[d])), atol);
        double a, b;
        if constexpr (AR::parity) {
            const double isc = hb_rcp_refined(sc);                 // both quotients correctly rounded
            a = hb_div_with(AR::mul(e5[d], h), sc, isc);
            b = hb_div_with(AR::mul(e3[d], h), sc, isc);
        } else {
            const double hs = h * hb_rcp_approx(sc);
            a = e5[d] * hs;
            b = e3[d] * hs;
        }
        n5 = fma(a, a, n5);
        n3 = fma(b, b, n3);
    }
}

// SciPy-style combined error norm (rk.py:2463-2467): err = |h| * n5 / sqrt((n5 + 0.01 n3) * n)
template <class AR>
HB_DEV double dop853_err_norm(double n5, double n3, double h, double ndim)
{
    if (n5 == 0.0 && n3 == 0.0) return 0.0;
    const double denom = AR::madd(0.01, n3, n5);
    if constexpr (AR::parity) return AR::div(AR::mul(fabs(h), n5), AR::sqrt(AR::mul(denom, ndim)));
    else return fabs(h) * n5 * hb_rsqrt_fast(denom * ndim);
}

// ---------------------------------------------------------------------------------------------
// Dense output of one accepted segment (rk.py:1836-1875): three extra stages (rows 13..15 of the
// extended tableau), then F[0..6].  k[0] = f_old, k[12] = f_new.
// ---------------------------------------------------------------------------------------------
template <class AR, int S, int R>
HB_DEV void ext_acc(double (&acc)[6], const double (&k)[13][6], const double (&kx)[3][6])
{
    if constexpr (R < S) {
        if constexpr (HB_DOP853_A[S][R] != 0.0) {
            constexpr double a = HB_DOP853_A[S][R];
#pragma unroll
            for (int d = 0; d < 6; ++d) {
                const double kv = (R < 13) ? k[R < 13 ? R : 0][d] : kx[R >= 13 ? R - 13 : 0][d];
                acc[d] = AR::madd(a, kv, acc[d]);
            }
        }
        ext_acc<AR, S, R + 1>(acc, k, kx);
    }
}

template <class AR, class RHS, int S>
HB_DEV void ext_stage(const double (&y_old)[6], double h, const double (&k)[13][6], double (&kx)[3][6],
                      const RHS &rhs)
{
    double acc[6], ys[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) acc[d] = 0.0;
    ext_acc<AR, S, 0>(acc, k, kx);
#pragma unroll
    for (int d = 0; d < 6; ++d) ys[d] = AR::madd(h, acc[d], y_old[d]);
    rhs(ys, kx[S - 13]);
}

template <class AR, int I, int R>
HB_DEV void d_acc(double (&acc)[6], const double (&k)[13][6], const double (&kx)[3][6])
{
    if constexpr (R < 16) {
        if constexpr (HB_DOP853_D[I][R] != 0.0) {
            constexpr double c = HB_DOP853_D[I][R];
#pragma unroll
            for (int d = 0; d < 6; ++d) {
                const double kv = (R < 13) ? k[R < 13 ? R : 0][d] : kx[R >= 13 ? R - 13 : 0][d];
                acc[d] = AR::madd(c, kv, acc[d]);
            }
        }
        d_acc<AR, I, R + 1>(acc, k, kx);
    }
}

template <class AR, int I>
HB_DEV void d_rows(double (&F)[7][6], double h, const double (&k)[13][6], const double (&kx)[3][6])
{
    if constexpr (I < 4) {
        double acc[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) acc[d] = 0.0;
        d_acc<AR, I, 0>(acc, k, kx);
#pragma unroll
        for (int d = 0; d < 6; ++d) F[3 + I][d] = AR::mul(h, acc[d]);
        d_rows<AR, I + 1>(F, h, k, kx);
    }
}

template <class AR, class RHS>
HB_DEV void dense_cache(const double (&y_old)[6], const double (&y_new)[6], double h,
                        const double (&k)[13][6], double (&F)[7][6], const RHS &rhs)
{
    double kx[3][6];
    ext_stage<AR, RHS, 13>(y_old, h, k, kx, rhs);
    ext_stage<AR, RHS, 14>(y_old, h, k, kx, rhs);
    ext_stage<AR, RHS, 15>(y_old, h, k, kx, rhs);
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        const double dy = AR::sub(y_new[d], y_old[d]);
        F[0][d] = dy;
        F[1][d] = AR::sub(AR::mul(h, k[0][d]), dy);
        F[2][d] = AR::sub(AR::mul(2.0, dy), AR::mul(h, AR::add(k[12][d], k[0][d])));
    }
    d_rows<AR, 0>(F, h, k, kx);
}

// The same interpolant restricted to ONE component c, with the stage rows STREAMED in ascending order into the
// accumulators of the three extra stages and of the four D rows (each accumulator still sees its terms in the
// reference's order, so the numbers are those of dense_cache).  `row(R, kr)` delivers k[R] for R = 5..12;
// k[0] = f(y_old) is evaluated here.  Used by the section-scan kernel, which needs ~1/6 of F and little state.
template <class AR, int R>
HB_DEV void stream_row(const double (&kr)[6], double kc, double (&a13)[6], double (&a14)[6], double (&a15)[6],
                       double (&da)[4])
{
    constexpr double c13 = HB_DOP853_A[13][R], c14 = HB_DOP853_A[14][R], c15 = HB_DOP853_A[15][R];
    constexpr double d0 = HB_DOP853_D[0][R], d1 = HB_DOP853_D[1][R], d2 = HB_DOP853_D[2][R], d3 = HB_DOP853_D[3][R];
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        if constexpr (c13 != 0.0) a13[d] = AR::madd(c13, kr[d], a13[d]);
        if constexpr (R < 14 && c14 != 0.0) a14[d] = AR::madd(c14, kr[d], a14[d]);
        if constexpr (R < 15 && c15 != 0.0) a15[d] = AR::madd(c15, kr[d], a15[d]);
    }
    if constexpr (d0 != 0.0) da[0] = AR::madd(d0, kc, da[0]);
    if constexpr (d1 != 0.0) da[1] = AR::madd(d1, kc, da[1]);
    if constexpr (d2 != 0.0) da[2] = AR::madd(d2, kc, da[2]);
    if constexpr (d3 != 0.0) da[3] = AR::madd(d3, kc, da[3]);
}

template <class AR, class RHS, class ROW, class PICK>
HB_DEV void dense_component(const double (&y_old)[6], const double (&y_new)[6], double h, ROW row, PICK pick,
                            const RHS &rhs, double (&f)[7])
{
    static_assert(HB_DOP853_A[13][13] == 0.0 && HB_DOP853_A[13][14] == 0.0 && HB_DOP853_A[14][14] == 0.0, "tableau");
    double a13[6], a14[6], a15[6], da[4] = {0.0, 0.0, 0.0, 0.0}, kr[6], ys[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) { a13[d] = 0.0; a14[d] = 0.0; a15[d] = 0.0; }
    rhs(y_old, kr);
    const double k0c = pick(kr);
    stream_row<AR, 0>(kr, k0c, a13, a14, a15, da);
    row(5, kr);  stream_row<AR, 5>(kr, pick(kr), a13, a14, a15, da);
    row(6, kr);  stream_row<AR, 6>(kr, pick(kr), a13, a14, a15, da);
    row(7, kr);  stream_row<AR, 7>(kr, pick(kr), a13, a14, a15, da);
    row(8, kr);  stream_row<AR, 8>(kr, pick(kr), a13, a14, a15, da);
    row(9, kr);  stream_row<AR, 9>(kr, pick(kr), a13, a14, a15, da);
    row(10, kr); stream_row<AR, 10>(kr, pick(kr), a13, a14, a15, da);
    row(11, kr); stream_row<AR, 11>(kr, pick(kr), a13, a14, a15, da);
    row(12, kr);
    const double k12c = pick(kr);
    stream_row<AR, 12>(kr, k12c, a13, a14, a15, da);
#pragma unroll
    for (int d = 0; d < 6; ++d) ys[d] = AR::madd(h, a13[d], y_old[d]);
    rhs(ys, kr);
    stream_row<AR, 13>(kr, pick(kr), a13, a14, a15, da);
#pragma unroll
    for (int d = 0; d < 6; ++d) ys[d] = AR::madd(h, a14[d], y_old[d]);
    rhs(ys, kr);
    stream_row<AR, 14>(kr, pick(kr), a13, a14, a15, da);
#pragma unroll
    for (int d = 0; d < 6; ++d) ys[d] = AR::madd(h, a15[d], y_old[d]);
    rhs(ys, kr);
    stream_row<AR, 15>(kr, pick(kr), a13, a14, a15, da);
    const double dy = AR::sub(pick(y_new), pick(y_old));
    f[0] = dy;
    f[1] = AR::sub(AR::mul(h, k0c), dy);
    f[2] = AR::sub(AR::mul(2.0, dy), AR::mul(h, AR::add(k12c, k0c)));
#pragma unroll
    for (int i = 0; i < 4; ++i) f[3 + i] = AR::mul(h, da[i]);
}

template <int I, int R>
HB_DEV void d_acc1(double &acc, const double (&kc)[16])
{
    if constexpr (R < 16) {
        if constexpr (HB_DOP853_D[I][R] != 0.0) {
            constexpr double c = HB_DOP853_D[I][R];
            acc = fma(c, kc[R], acc);
        }
        d_acc1<I, R + 1>(acc, kc);
    }
}

// Screening for the sparse step records of hb_cr3bp_section2 (records = "near"): can the dense interpolant of this
// accepted step come anywhere near the section plane?  Evaluates component C of the interpolant in FAST arithmetic
// (FMA-contracted, approximate reciprocal square roots: |difference to the separately rounded values| < 1e-12 for O(1)
// states) and applies the quiet-step bound of the scan kernel -- |p(x) - (y0 + x F0)| <= sum_{i>=1}|F_i| / 4 on [0, 1] --
// with the margin widened by 1e-8 (1 + |y_C| + |offset|), four orders above that difference.  Returns false only when
// every point of the interpolant on [0, 1] stays on one side of the plane, farther from it than the on-surface
// tolerance: such a step has no sample on or across the section and its record is not needed by the scan.
// RHSF: the vector field in ArFast arithmetic.
template <int C, class RHSF>
HB_DEV bool dop853_step_near_plane(const double (&y_old)[6], const double (&y_new)[6], double h,
                                   const double (&k)[13][6], const RHSF &rhs, double offset, double tol)
{
    double kx[3][6];
    ext_stage<ArFast, RHSF, 13>(y_old, h, k, kx, rhs);
    ext_stage<ArFast, RHSF, 14>(y_old, h, k, kx, rhs);
    ext_stage<ArFast, RHSF, 15>(y_old, h, k, kx, rhs);
    double S = 0.0;
    {
        // rows 3..6 of F, component C only: h * sum_R D[i][R] k_R[C]
        double kc[16];
#pragma unroll
        for (int R = 0; R < 13; ++R) kc[R] = k[R][C];
#pragma unroll
        for (int R = 13; R < 16; ++R) kc[R] = kx[R - 13][C];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        d_acc1<0, 0>(a0, kc);
        d_acc1<1, 0>(a1, kc);
        d_acc1<2, 0>(a2, kc);
        d_acc1<3, 0>(a3, kc);
        S = fabs(h * a0) + fabs(h * a1) + fabs(h * a2) + fabs(h * a3);
    }
    const double dy = y_new[C] - y_old[C];
    const double f1 = fma(h, k[0][C], -dy);
    const double f2 = fma(-h, k[12][C] + k[0][C], 2.0 * dy);
    S += fabs(f1) + fabs(f2);
    const double g_old = y_old[C] - offset, g_new = y_new[C] - offset;
    const double margin = 0.25 * S + tol + 1.0e-8 * (1.0 + fabs(y_old[C]) + fabs(dy) + fabs(offset));
    const bool same = (g_old > 0.0 && g_new > 0.0) || (g_old < 0.0 && g_new < 0.0);
    return !(same && fmin(fabs(g_old), fabs(g_new)) > margin);
}

// _dop853_eval_dense (rk.py:1989-2003): alternating x / (1-x) Horner form.
template <class AR>
HB_DEV void dense_eval(const double (&y_old)[6], const double (&F)[7][6], double x, double (&out)[6])
{
    const double omx = AR::sub(1.0, x);
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        double v = 0.0;
#pragma unroll
        for (int i = 6; i >= 0; --i) {
            v = AR::add(v, F[i][d]);
            v = AR::mul(v, ((6 - i) % 2 == 0) ? x : omx);
        }
        out[d] = AR::add(v, y_old[d]);
    }
}

