// hb_rkgen.cuh -- explicit Runge-Kutta stage machinery for an arbitrary compile-time tableau on a 6-vector.
// TAB provides  static constexpr int S;  static constexpr double a(i,j), b(i);  RHS is a functor
// void operator()(const double (&y)[6], double (&dy)[6]) const.  Zero tableau entries cost nothing.
// Reference: rk_embedded_step_jit_kernel (hiten/algorithms/integrators/rk.py:155-215) and
// _integrate_rk_ham (hiten/algorithms/poincare/centermanifold/backend.py:144-184):
//   y_stage = y; for j < i with a_ij != 0: y_stage += (h * a_ij) * k_j ;  y_new = y + sum_j (h * b_j) * k_j
#pragma once
#include "hb_common.cuh"

struct TabRK4 { static constexpr int S = 4; static constexpr double a(int i, int j) { return HB_RK4_A[i][j]; } static constexpr double b(int i) { return HB_RK4_B[i]; } };
struct TabRK6 { static constexpr int S = 7; static constexpr double a(int i, int j) { return HB_RK6_A[i][j]; } static constexpr double b(int i) { return HB_RK6_B[i]; } };
struct TabRK8 { static constexpr int S = 13; static constexpr double a(int i, int j) { return HB_RK8_A[i][j]; } static constexpr double b(int i) { return HB_RK8_B[i]; } };

template <class AR, class TAB, int I, int J>
HB_DEV void g_stage_acc(double (&ys)[6], const double (&k)[TAB::S][6], double h)
{
    if constexpr (J < I) {
        if constexpr (TAB::a(I, J) != 0.0) {
            constexpr double a = TAB::a(I, J);
            const double ha = AR::mul(h, a);
#pragma unroll
            for (int d = 0; d < 6; ++d) ys[d] = AR::madd(ha, k[J][d], ys[d]);
        }
        g_stage_acc<AR, TAB, I, J + 1>(ys, k, h);
    }
}
// is stage I referenced by a later stage or by B?  (the 7th DOPRI5 stage of "RK6" is not)
template <class TAB, int I, int R>
constexpr bool g_stage_used()
{
    if constexpr (R >= TAB::S) return TAB::b(I) != 0.0;
    else return (TAB::a(R, I) != 0.0) || g_stage_used<TAB, I, R + 1>();
}
template <class AR, class TAB, class RHS, int I>
HB_DEV void g_run_stages(const RHS &rhs, const double (&y)[6], double (&k)[TAB::S][6], double h)
{
    if constexpr (I < TAB::S) {
        if constexpr (g_stage_used<TAB, I, I + 1>()) {
            double ys[6];
#pragma unroll
            for (int d = 0; d < 6; ++d) ys[d] = y[d];
            g_stage_acc<AR, TAB, I, 0>(ys, k, h);
            rhs(ys, k[I]);
        }
        g_run_stages<AR, TAB, RHS, I + 1>(rhs, y, k, h);
    }
}
template <class AR, class TAB, int J>
HB_DEV void g_high_acc(double (&yn)[6], const double (&k)[TAB::S][6], double h)
{
    if constexpr (J < TAB::S) {
        if constexpr (TAB::b(J) != 0.0) {
            constexpr double b = TAB::b(J);
            const double hb = AR::mul(h, b);
#pragma unroll
            for (int d = 0; d < 6; ++d) yn[d] = AR::madd(hb, k[J][d], yn[d]);
        }
        g_high_acc<AR, TAB, J + 1>(yn, k, h);
    }
}

