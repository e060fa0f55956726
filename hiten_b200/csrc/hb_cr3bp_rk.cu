// hb_cr3bp_rk.cu -- the other integrators of the reference's factory on the 6-state CR3BP system:
// adaptive RK45 (DOPRI5) and fixed-step RK4 / "RK6" (= 7-stage DOPRI5) / RK8 (Prince-Dormand 8(7)13M),
// each in the three modes of hb_cr3bp.cu (end state, dense grid, terminal plane event).
// One trajectory per thread, persistent work queue, tableau resolved at compile time (hb_rkgen.cuh).
//
// Reference routines (paths relative to hiten/):
//   rk45_step_jit_kernel           algorithms/integrators/rk.py:842-898
//   _RK45._integrate_rk45 / _until_event                   rk.py:1269-1399, 1460-1633
//   _rk45_build_Q_cache / _rk45_eval_dense / _rk45_refine_in_step   rk.py:971-1092
//   rk_embedded_step_jit_kernel    rk.py:155-215
//   _FixedStepRK._integrate_fixed_rk / _until_event         rk.py:533-588, 657-718
//   _hermite_eval_dense / _hermite_refine_in_step           rk.py:273-391
//   RungeKutta / AdaptiveRK / FixedRK factories             rk.py:2871-2978
#include "hb_cr3bp_common.cuh"
#include "hb_rkgen.cuh"
#include "hb_rk45.cuh"

namespace {
using namespace hbc;

enum { RMODE_FINAL = 0, RMODE_DENSE = 1, RMODE_EVENT = 2 };

struct Tab4 { static constexpr int S = 4; static constexpr double a(int i, int j) { return HB_RK4_A[i][j]; } static constexpr double b(int i) { return HB_RK4_B[i]; } };
struct Tab6 { static constexpr int S = 7; static constexpr double a(int i, int j) { return HB_RK6_A[i][j]; } static constexpr double b(int i) { return HB_RK6_B[i]; } };
struct Tab8 { static constexpr int S = 13; static constexpr double a(int i, int j) { return HB_RK8_A[i][j]; } static constexpr double b(int i) { return HB_RK8_B[i]; } };

struct RkExtra {           // fixed-step grid (np.linspace(t0, tf, n_fixed + 1)) when no explicit t_eval is given
    int n_fixed;
};

// _hermite_eval_dense (rk.py:296-312)
template <class AR>
HB_DEV void hermite_eval(const double (&y0)[6], const double (&f0)[6], const double (&y1)[6], const double (&f1)[6],
                         double x, double h, double (&out)[6])
{
    const double x2 = AR::mul(x, x), x3 = AR::mul(x2, x);
    const double H00 = AR::add(AR::sub(AR::mul(2.0, x3), AR::mul(3.0, x2)), 1.0);
    const double H10 = AR::add(AR::sub(x3, AR::mul(2.0, x2)), x);
    const double H01 = AR::add(AR::mul(-2.0, x3), AR::mul(3.0, x2));
    const double H11 = AR::sub(x3, x2);
#pragma unroll
    for (int d = 0; d < 6; ++d)
        out[d] = AR::add(AR::add(AR::add(AR::mul(H00, y0[d]), AR::mul(H10, AR::mul(h, f0[d]))), AR::mul(H01, y1[d])),
                         AR::mul(H11, AR::mul(h, f1[d])));
}

// In-step bisection on x in [0,1] (rk.py:1072-1092 / 373-391): EVAL(x, out) evaluates the interpolant.
template <class AR, class EVAL>
HB_DEV double refine_bisect(const PropParams &p, double g_left, double h, EVAL eval, double (&ym)[6])
{
    double a = 0.0, b = 1.0;
    for (int it = 0; it < 128; ++it) {
        const double mid = AR::mul(0.5, AR::add(a, b));
        eval(mid, ym);
        const double g_mid = AR::sub(pick6(ym, p.ev_idx), p.ev_off);
        if (fabs(g_mid) <= p.gtol) return mid;
        if (hb_crossed_direction(g_left, g_mid, p.ev_dir)) b = mid;
        else { a = mid; g_left = g_mid; }
        if (AR::mul(AR::sub(b, a), fabs(h)) <= p.xtol) break;
    }
    eval(b, ym);
    return b;
}

// ---------------------------------------------------------------------------------------------
// adaptive RK45
// ---------------------------------------------------------------------------------------------
template <class AR, int MODE>
__global__ void __launch_bounds__(256) k_rk45_6(const PropParams p)
{
    const Cr3bpRhs<AR, 2> rhs{p};
    for (;;) {
        const long long idx = hb_fetch_index(p.ws);
        if (idx >= p.n) break;
        double y[6], yh[6], k[6][6], k6[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) y[d] = p.y0[(long long)d * p.n + idx];
        rhs(y, k[0]);
        double t = p.t0;
        const double tf = p.tf_arr ? p.tf_arr[idx] : p.tf;
        double h = initial_step<AR>(y, k[0], p);
        double err_prev = -1.0, g_prev = 0.0;
        if (MODE == RMODE_EVENT) g_prev = AR::sub(pick6(y, p.ev_idx), p.ev_off);
        int nacc = 0, nrej = 0, cursor = 0, fin = -1;
        long long attempts = 0;
        bool wrote = false;
        while ((t - tf) < 0.0 && fin < 0) {
            h = hb_clamp_step(h, p.max_step, p.min_step);
            if (AR::add(t, h) > tf) h = fabs(AR::sub(tf, t));
            g_run_stages<AR, Tab45, Cr3bpRhs<AR, 2>, 1>(rhs, y, k, h);
#pragma unroll
            for (int d = 0; d < 6; ++d) yh[d] = y[d];
            g_high_acc<AR, Tab45, 0>(yh, k, h);
            rhs(yh, k6);
            double ev[6];
#pragma unroll
            for (int d = 0; d < 6; ++d) ev[d] = 0.0;
            rk45_err_acc<AR, 0>(ev, k, k6, h);
#pragma unroll
            for (int d = 0; d < 6; ++d) {
                const double sc = AR::madd(p.rtol, fmax(fabs(y[d]), fabs(yh[d])), p.atol);
                ev[d] = AR::div(ev[d], sc);
            }
            const double err = AR::div(norm2_ext6(ev), AR::sqrt(6.0));       // np.linalg.norm / sqrt(n), rk.py:1333
            ++attempts;
            if (err <= 1.0) {
                const double t_new = AR::add(t, h);
                ++nacc;
                const bool last = !((t_new - tf) < 0.0);
                if (MODE == RMODE_EVENT) {
                    const double g_new = AR::sub(pick6(yh, p.ev_idx), p.ev_off);
                    if (hb_event_crossed(g_prev, g_new, p.ev_dir)) {
                        double Q[6][4] = {}, ym[6];
                        rk45_q_acc<AR, 0, 0>(Q, k, k6);
                        auto eval = [&](double x, double (&o)[6]) { rk45_eval<AR>(y, Q, x, h, o); };
                        const double xh = refine_bisect<AR>(p, g_prev, h, eval, ym);
                        p.t_hit[idx] = AR::madd(xh, h, t);
#pragma unroll
                        for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = ym[d];
                        fin = HB_TRAJ_HIT;
                        wrote = true;
                    } else {
                        g_prev = g_new;
                    }
                } else if (MODE == RMODE_DENSE || last) {
                    const double hseg = AR::sub(t_new, t);
                    double Q[6][4] = {}, yo[6];
                    rk45_q_acc<AR, 0, 0>(Q, k, k6);
                    if (MODE == RMODE_DENSE) {
                        while (cursor < p.m) {
                            const double tq = p.t_eval[cursor];
                            if (!(last || tq < t_new)) break;
                            rk45_eval<AR>(y, Q, AR::div(AR::sub(tq, t), hseg), hseg, yo);
                            double *o = p.dense_out + ((long long)idx * p.m + cursor) * 6;
#pragma unroll
                            for (int d = 0; d < 6; ++d) o[d] = yo[d];
                            ++cursor;
                        }
                    } else {
                        rk45_eval<AR>(y, Q, AR::div(AR::sub(tf, t), hseg), hseg, yo);
#pragma unroll
                        for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = yo[d];
                        wrote = true;
                    }
                }
                t = t_new;
#pragma unroll
                for (int d = 0; d < 6; ++d) { y[d] = yh[d]; k[0][d] = k6[d]; }
                h = AR::mul(h, hb_pi_accept_factor<AR>(err, err_prev, 5.0));
                err_prev = err;
            } else {
                ++nrej;
                h = AR::mul(h, hb_pi_reject_factor<AR>(err, 5.0));
                h = hb_clamp_step(h, p.max_step, p.min_step);
            }
            if (fin < 0) {
                if (!(h == h) || !(err == err)) fin = HB_TRAJ_NONFINITE;
                else if (attempts >= p.max_attempts) fin = HB_TRAJ_MAXSTEPS;
            }
        }
        if (fin < 0) fin = HB_TRAJ_OK;
        if (MODE == RMODE_EVENT && fin != HB_TRAJ_HIT) p.t_hit[idx] = t;
        if (MODE != RMODE_DENSE && !wrote) {
#pragma unroll
            for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = y[d];
        }
        if (MODE == RMODE_DENSE) {
            for (; cursor < p.m; ++cursor) {       // zero-length span or early termination: hold the last state
                double *o = p.dense_out + ((long long)idx * p.m + cursor) * 6;
#pragma unroll
                for (int d = 0; d < 6; ++d) o[d] = y[d];
            }
        }
        p.nacc[idx] = nacc; p.nrej[idx] = nrej; p.status[idx] = fin;
    }
}

// ---------------------------------------------------------------------------------------------
// fixed-step RK4 / RK6 / RK8 over a grid: t_eval[m] when given, else linspace(t0, tf, n_fixed + 1)
// ---------------------------------------------------------------------------------------------
template <class AR, class TAB, int MODE>
__global__ void __launch_bounds__(256) k_rkfixed_6(const PropParams p, const RkExtra x)
{
    const Cr3bpRhs<AR, 2> rhs{p};
    const int npts = p.t_eval ? p.m : x.n_fixed + 1;
    for (;;) {
        const long long idx = hb_fetch_index(p.ws);
        if (idx >= p.n) break;
        double y[6], yn[6], fp[6], k[TAB::S][6];
#pragma unroll
        for (int d = 0; d < 6; ++d) y[d] = p.y0[(long long)d * p.n + idx];
        const double tf = p.tf_arr ? p.tf_arr[idx] : p.tf;
        const double lin_step = (x.n_fixed > 0) ? AR::div(AR::sub(tf, p.t0), (double)x.n_fixed) : 0.0;
        auto grid = [&](int i) -> double {
            if (p.t_eval) return p.t_eval[i];
            return (i == x.n_fixed) ? tf : AR::madd((double)i, lin_step, p.t0);    // numpy.linspace
        };
        rhs(y, fp);
        double g_prev = 0.0;
        if (MODE == RMODE_EVENT) g_prev = AR::sub(pick6(y, p.ev_idx), p.ev_off);
        if (MODE == RMODE_DENSE) {
            double *o = p.dense_out + (long long)idx * p.m * 6;
#pragma unroll
            for (int d = 0; d < 6; ++d) o[d] = y[d];
        }
        int fin = HB_TRAJ_OK;
        double t_end = grid(npts - 1);
        for (int i = 0; i + 1 < npts; ++i) {
            const double tn = grid(i);
            const double h = AR::sub(grid(i + 1), tn);
#pragma unroll
            for (int d = 0; d < 6; ++d) k[0][d] = fp[d];
            g_run_stages<AR, TAB, Cr3bpRhs<AR, 2>, 1>(rhs, y, k, h);
#pragma unroll
            for (int d = 0; d < 6; ++d) yn[d] = y[d];
            g_high_acc<AR, TAB, 0>(yn, k, h);
            double fn[6];
            rhs(yn, fn);
            if (MODE == RMODE_EVENT) {
                const double g_new = AR::sub(pick6(yn, p.ev_idx), p.ev_off);
                if (hb_event_crossed(g_prev, g_new, p.ev_dir)) {
                    double ym[6];
                    auto eval = [&](double xx, double (&o)[6]) { hermite_eval<AR>(y, fp, yn, fn, xx, h, o); };
                    const double xh = refine_bisect<AR>(p, g_prev, h, eval, ym);
                    t_end = AR::madd(xh, h, tn);
#pragma unroll
                    for (int d = 0; d < 6; ++d) y[d] = ym[d];
                    fin = HB_TRAJ_HIT;
                    break;
                }
                g_prev = g_new;
            }
#pragma unroll
            for (int d = 0; d < 6; ++d) { y[d] = yn[d]; fp[d] = fn[d]; }
            if (MODE == RMODE_DENSE) {
                double *o = p.dense_out + ((long long)idx * p.m + i + 1) * 6;
#pragma unroll
                for (int d = 0; d < 6; ++d) o[d] = y[d];
            }
            if (!(y[0] == y[0])) { fin = HB_TRAJ_NONFINITE; }
        }
        if (MODE != RMODE_DENSE) {
#pragma unroll
            for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = y[d];
        }
        if (MODE == RMODE_EVENT) p.t_hit[idx] = t_end;
        p.nacc[idx] = npts - 1; p.nrej[idx] = 0; p.status[idx] = fin;
    }
}

template <class AR, int MODE>
int launch_method(const PropParams &p, int method, int n_fixed, cudaStream_t st)
{
    HB_CUDA_TRY(cudaMemsetAsync(p.ws, 0, sizeof(HbWorkspace), st));
    const int threads = 256;
    long long blocks = (p.n + threads - 1) / threads;
    const long long cap = 2LL * sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const RkExtra x{n_fixed};
    const unsigned g = (unsigned)blocks;
    switch (method) {
    case HB_RK45: k_rk45_6<AR, MODE><<<g, threads, 0, st>>>(p); break;
    case HB_RK4: k_rkfixed_6<AR, Tab4, MODE><<<g, threads, 0, st>>>(p, x); break;
    case HB_RK6: k_rkfixed_6<AR, Tab6, MODE><<<g, threads, 0, st>>>(p, x); break;
    case HB_RK8: k_rkfixed_6<AR, Tab8, MODE><<<g, threads, 0, st>>>(p, x); break;
    default: return HB_ERR_UNSUPPORTED;
    }
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

}  // namespace

// Called by hb_cr3bp_propagate / _dense / _event (hb_cr3bp.cu) for every method other than DOP853.
// mode: 0 end state, 1 dense grid, 2 terminal event.
int hb_rk_dispatch(const hbc::PropParams &p, int method, int arith, int mode, int n_fixed, cudaStream_t st)
{
    const bool fixed = method == HB_RK4 || method == HB_RK6 || method == HB_RK8;
    if (!fixed && method != HB_RK45) return HB_ERR_UNSUPPORTED;
    if (fixed && mode != 1 && n_fixed < 1) return HB_ERR_BADARG;       // a fixed-step run needs its step count
    if (arith == HB_ARITH_PARITY) {
        if (mode == 0) return launch_method<ArParity, RMODE_FINAL>(p, method, n_fixed, st);
        if (mode == 1) return launch_method<ArParity, RMODE_DENSE>(p, method, n_fixed, st);
        return launch_method<ArParity, RMODE_EVENT>(p, method, n_fixed, st);
    }
    if (mode == 0) return launch_method<ArFast, RMODE_FINAL>(p, method, n_fixed, st);
    if (mode == 1) return launch_method<ArFast, RMODE_DENSE>(p, method, n_fixed, st);
    return launch_method<ArFast, RMODE_EVENT>(p, method, n_fixed, st);
}
