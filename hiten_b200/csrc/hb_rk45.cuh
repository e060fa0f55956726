// hb_rk45.cuh -- DOPRI5(4) pieces shared by the 6-state CR3BP kernel (hb_cr3bp_rk.cu) and the polynomial-Hamiltonian
// kernel (hb_cm.cu): tableau view, error accumulation, SciPy-style dense-output cache and evaluation.
// Reference: hiten/algorithms/integrators/rk.py  rk45_step_jit_kernel :842-898, _rk45_build_Q_cache :971-1000,
// _rk45_eval_dense :1021-1034.
#pragma once
#include "hb_common.cuh"

struct Tab45 { static constexpr int S = 6; static constexpr double a(int i, int j) { return j < 5 ? HB_RK45_A[i][j] : 0.0; } static constexpr double b(int i) { return HB_RK45_B[i]; } };

// err_vec += (h * E_j) * k_j   (rk.py:892-896); k6 = f(t+h, y_high)
template <class AR, int J>
HB_DEV void rk45_err_acc(double (&ev)[6], const double (&k)[6][6], const double (&k6)[6], double h)
{
    if constexpr (J < 7) {
        if constexpr (HB_RK45_E[J] != 0.0) {
            constexpr double c = HB_RK45_E[J];
            const double hc = AR::mul(h, c);
#pragma unroll
            for (int d = 0; d < 6; ++d) ev[d] = AR::madd(hc, (J < 6) ? k[J < 6 ? J : 0][d] : k6[d], ev[d]);
        }
        rk45_err_acc<AR, J + 1>(ev, k, k6, h);
    }
}
// Q[d][c] = sum_r P[r][c] K[r][d]   (rk.py:988-997)
template <class AR, int R, int Cc>
HB_DEV void rk45_q_acc(double (&Q)[6][4], const double (&k)[6][6], const double (&k6)[6])
{
    if constexpr (Cc < 4) {
        if constexpr (R < 7) {
            if constexpr (HB_RK45_P[R][Cc] != 0.0) {
                constexpr double pc = HB_RK45_P[R][Cc];
#pragma unroll
                for (int d = 0; d < 6; ++d) Q[d][Cc] = AR::madd(pc, (R < 6) ? k[R < 6 ? R : 0][d] : k6[d], Q[d][Cc]);
            }
            rk45_q_acc<AR, R + 1, Cc>(Q, k, k6);
        } else {
            rk45_q_acc<AR, 0, Cc + 1>(Q, k, k6);
        }
    }
}
// _rk45_eval_dense (rk.py:1021-1034)
template <class AR>
HB_DEV void rk45_eval(const double (&y_old)[6], const double (&Q)[6][4], double x, double hseg, double (&out)[6])
{
    double pw[4], val = x;
#pragma unroll
    for (int c = 0; c < 4; ++c) { pw[c] = val; val = AR::mul(val, x); }
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc = AR::madd(Q[d][c], pw[c], acc);
        out[d] = AR::madd(hseg, acc, y_old[d]);
    }
}
