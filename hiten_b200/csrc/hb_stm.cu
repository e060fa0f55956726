// hb_stm.cu -- batched 42-dimensional state + STM (variational) propagation: DOP853 (tuned), RK45 and fixed-step RK4/6/8.
//
// A lane group of 8 per trajectory (4 trajectories per warp): lane j < 6 carries column j of the
// state-transition matrix Phi (6 values), lane 6 carries the state x (6 values), lane 7 idles.
// dPhi/dt = F(x) Phi decomposes by columns, so a lane only needs F -- i.e. the position (x,y,z) of the
// current stage vector, broadcast from lane 6 with three group shuffles -- and its own column.  Every
// lane therefore runs the same 6-component DOP853 machinery as the 6-state kernel (13 x 6 stage values
// in registers, no shared memory, no spills); the only other communication is the butterfly reduction
// of the error norm over the group.  The Jacobian entries are evaluated redundantly by all lanes: under
// SIMT that costs exactly what a single lane would.
//
// Reference routines (paths relative to hiten/):
//   _jacobian_crtbp   algorithms/dynamics/rtbp.py:77-165     _var_equations  rtbp.py:168-255
//   _compute_stm      algorithms/dynamics/rtbp.py:258-340    (PHI0 = [I6 row-major, x0], flip_indices = slice(36,42))
// PHI_vec[6*i + j] = Phi[i][j]  ->  lane j, slot i;   PHI_vec[36 + i] = x[i] -> lane 6, slot i.
//
// Note on r2**1.5 / r2**2.5: the reference calls libm pow (not correctly rounded);
// here r^3 = r2*sqrt(r2) and r^5 = (r2*r2)*sqrt(r2), which agrees with it to an ulp or two.  STM parity
// is therefore a tolerance statement (1e-8 relative to |Phi|), not a bit-exact one.
#include "hb_dop853.cuh"
#include "hb_rkgen.cuh"
#include "hb_rk45.cuh"

namespace {

enum { SMODE_FINAL = 0, SMODE_DENSE = 1 };

struct StmParams {
    double mu, om;
    int neg_phi, neg_state;   // fwd == -1 wrapper: which derivative blocks are negated
    double rtol, atol, max_step, min_step;
    long long max_attempts;
    long long n;
    const double *x0;         // SoA [6][n]
    double t0, tf;
    const double *tf_arr;
    double *phi_out;          // FINAL: [n][42] (the reference's flat PHI row at tf)
    int *nacc, *nrej, *status;
    HbWorkspace *ws;
    const double *t_eval;     // DENSE: [m] shared or [n][m]
    int m, t_eval_per_traj;
    double *dense_out;        // [n][m][42]
};

#ifndef HB_STM_PARITY_FIELD
#define HB_STM_PARITY_FIELD 0
#endif
// The same argument applies to the stage sums, error sums and dense output of the 42-state system: with a vector field
// that cannot reproduce the reference's bits, separately rounded multiply-adds (two FP64 pipe slots each) only halve the
// throughput.  Both variants therefore integrate with FMA-contracted arithmetic (AS below); `parity` keeps what decides
// the STEP SEQUENCE in the reference's arithmetic -- the controller's restated libm pow, the initial step, the time
// updates -- so step counts still track the reference (44 accepted steps on member 0 of the halo family).  Measured
// against the reference on the 100-member family: <= 2e-11 of |Phi| either way (criterion: 1e-8).
#ifndef HB_STM_PARITY_STAGES
#define HB_STM_PARITY_STAGES 0
#endif
#ifndef HB_STM_RHS_NOINLINE
#define HB_STM_RHS_NOINLINE HB_STM_PARITY_FIELD       // the called form only pays for the large (old parity) field
#endif
struct StmV6 { double a, b, c, d, e, f; };

template <class AR>
struct StmRhsImpl {
    const StmParams &p;
    unsigned gmask;           // the 8 lanes of this group
    bool is_state;            // lane 6
    HB_DEV void operator()(const double (&v)[6], double (&dv)[6]) const
    {
        const double x = __shfl_sync(gmask, v[0], 6, 8);
        const double y = __shfl_sync(gmask, v[1], 6, 8);
        const double z = __shfl_sync(gmask, v[2], 6, 8);
        const double mu = p.mu, mu2 = p.om;
        double oxx, oyy, ozz, oxy, oxz, oyz, ax, ay, az;
        // The reference evaluates these entries with libm pow (r2**1.5, r2**2.5: not correctly rounded) and sums F @ Phi
        // with a SIMD dot product, so NO arithmetic reproduces its bits here (header note): the 42-state path is a
        // tolerance statement in both variants.  The separately rounded form below (2 sqrt.rn, 4 refined reciprocals,
        // 10 correctly rounded quotients: ~230 instructions, executed by all 8 lanes of a group for ONE trajectory) buys
        // nothing over the FMA / rsqrt form (~60 instructions, |difference| ~ 1e-15) -- measured against the reference:
        // 2.0e-11 vs 1.2e-11 of |Phi| on the 100-member halo family -- so both variants use the latter
        // (HB_STM_PARITY_FIELD=1 brings the old form back).  Stage sums, error norm and controller stay in AR.
        if constexpr (AR::parity && HB_STM_PARITY_FIELD) {
            const double xm = AR::add(x, mu), xo = AR::sub(x, mu2);
            const double xm2 = AR::mul(xm, xm), xo2 = AR::mul(xo, xo), yy = AR::mul(y, y), zz = AR::mul(z, z);
            const double r2 = AR::add(AR::add(xm2, yy), zz);
            const double R2 = AR::add(AR::add(xo2, yy), zz);
            const double s1 = AR::sqrt(r2), s2 = AR::sqrt(R2);
            const double r3 = AR::mul(r2, s1), r5 = AR::mul(AR::mul(r2, r2), s1);
            const double R3 = AR::mul(R2, s2), R5 = AR::mul(AR::mul(R2, R2), s2);
            const double ir3 = hb_rcp_refined(r3), ir5 = hb_rcp_refined(r5);
            const double iR3 = hb_rcp_refined(R3), iR5 = hb_rcp_refined(R5);
            const double a5 = hb_div_with(mu2, r5, ir5), b5 = hb_div_with(mu, R5, iR5);   // mu2/r5, mu/R5
            const double a3 = hb_div_with(mu2, r3, ir3), b3 = hb_div_with(mu, R3, iR3);   // mu2/r3, mu/R3
            const double common = AR::add(a3, b3);
            oxx = AR::sub(AR::add(AR::add(1.0, AR::mul(AR::mul(a5, 3.0), xm2)), AR::mul(AR::mul(b5, 3.0), xo2)), common);
            oyy = AR::sub(AR::add(AR::add(1.0, AR::mul(AR::mul(a5, 3.0), yy)), AR::mul(AR::mul(b5, 3.0), yy)), common);
            ozz = AR::sub(AR::add(AR::add(0.0, AR::mul(AR::mul(a5, 3.0), zz)), AR::mul(AR::mul(b5, 3.0), zz)), common);
            const double cross = AR::add(hb_div_with(AR::mul(mu2, xm), r5, ir5), hb_div_with(AR::mul(mu, xo), R5, iR5));
            oxy = AR::mul(AR::mul(3.0, y), cross);
            oxz = AR::mul(AR::mul(3.0, z), cross);
            oyz = AR::mul(AR::mul(AR::mul(3.0, y), z), AR::add(a5, b5));
            // accelerations, rtbp.py:233-242 (uses lane-local velocities; only lane 6 keeps them)
            ax = AR::add(AR::sub(AR::sub(x, AR::mul(mu2, hb_div_with(xm, r3, ir3))), AR::mul(mu, hb_div_with(xo, R3, iR3))),
                         AR::mul(2.0, v[4]));
            ay = AR::sub(AR::sub(AR::sub(y, AR::mul(mu2, hb_div_with(y, r3, ir3))), AR::mul(mu, hb_div_with(y, R3, iR3))),
                         AR::mul(2.0, v[3]));
            az = AR::sub(AR::mul(-mu2, hb_div_with(z, r3, ir3)), AR::mul(mu, hb_div_with(z, R3, iR3)));
        } else {
            const double xm = x + mu, xo = x - mu2;
            const double yz = fma(y, y, z * z);
            const double r2 = fma(xm, xm, yz), R2 = fma(xo, xo, yz);
            const double i1 = hb_rsqrt_fast(r2), i2 = hb_rsqrt_fast(R2);
            const double i1s = i1 * i1, i2s = i2 * i2;
            const double a3 = mu2 * (i1s * i1), b3 = mu * (i2s * i2);          // mu2/r^3, mu/R^3
            const double a5 = 3.0 * a3 * i1s, b5 = 3.0 * b3 * i2s;             // 3 mu2/r^5, 3 mu/R^5
            const double common = a3 + b3, s5 = a5 + b5;
            const double cross = fma(a5, xm, b5 * xo);
            oxx = fma(a5, xm * xm, fma(b5, xo * xo, 1.0 - common));
            oyy = fma(s5, y * y, 1.0 - common);
            ozz = fma(s5, z * z, -common);
            oxy = y * cross;
            oxz = z * cross;
            oyz = y * z * s5;
            ax = fma(2.0, v[4], x) - fma(a3, xm, b3 * xo);
            ay = fma(-2.0, v[3], y) - common * y;
            az = -common * z;
        }
        double o0, o1, o2, o3, o4, o5;
        if (is_state) {
            o0 = v[3]; o1 = v[4]; o2 = v[5]; o3 = ax; o4 = ay; o5 = az;
        } else {
            // column of F*Phi in the reference's k-order (rtbp.py:219-225); zero entries of F add exact zeros
            o0 = v[3]; o1 = v[4]; o2 = v[5];
            o3 = AR::madd(2.0, v[4], AR::madd(oxz, v[2], AR::madd(oxy, v[1], AR::mul(oxx, v[0]))));
            o4 = AR::madd(-2.0, v[3], AR::madd(oyz, v[2], AR::madd(oyy, v[1], AR::mul(oxy, v[0]))));
            o5 = AR::madd(ozz, v[2], AR::madd(oyz, v[1], AR::mul(oxz, v[0])));
        }
        const bool neg = is_state ? (p.neg_state != 0) : (p.neg_phi != 0);
        dv[0] = neg ? -o0 : o0; dv[1] = neg ? -o1 : o1; dv[2] = neg ? -o2 : o2;
        dv[3] = neg ? -o3 : o3; dv[4] = neg ? -o4 : o4; dv[5] = neg ? -o5 : o5;
    }
};

// ONE copy of the variational vector field per kernel: inlined into the 13 stages and the 3 dense-output stages it makes
// the step loop far larger than the instruction cache (ncu: `no_instruction` was the top stall reason, 1.4 cycles per
// issue); called, the loop fits: 4.41e8 -> 5.19e8 steps/s in the parity variant (same bits).
template <class AR>
__device__ __noinline__ StmV6 stm_rhs_call(double mu, double om, int neg_phi, int neg_state, unsigned gmask, bool is_state,
                                           double v0, double v1, double v2, double v3, double v4, double v5)
{
    StmParams q{};
    q.mu = mu; q.om = om; q.neg_phi = neg_phi; q.neg_state = neg_state;
    const StmRhsImpl<AR> impl{q, gmask, is_state};
    const double v[6] = {v0, v1, v2, v3, v4, v5};
    double dv[6];
    impl(v, dv);
    return StmV6{dv[0], dv[1], dv[2], dv[3], dv[4], dv[5]};
}

template <class AR>
struct StmRhs {
    const StmParams &p;
    unsigned gmask;
    bool is_state;
    HB_DEV void operator()(const double (&v)[6], double (&dv)[6]) const
    {
        if constexpr (HB_STM_RHS_NOINLINE && AR::parity) {
            const StmV6 r = stm_rhs_call<AR>(p.mu, p.om, p.neg_phi, p.neg_state, gmask, is_state, v[0], v[1], v[2], v[3],
                                             v[4], v[5]);
            dv[0] = r.a; dv[1] = r.b; dv[2] = r.c; dv[3] = r.d; dv[4] = r.e; dv[5] = r.f;
        } else {                          // the FMA-contracted field is small: inlined it is 13 % faster than called
            const StmRhsImpl<AR> impl{p, gmask, is_state};
            impl(v, dv);
        }
    }
};

HB_DEV double group_sum(double v, unsigned gmask)
{
    v += __shfl_xor_sync(gmask, v, 4, 8);
    v += __shfl_xor_sync(gmask, v, 2, 8);
    v += __shfl_xor_sync(gmask, v, 1, 8);
    return v;
}

template <class AR> struct StageArith { using type = ArFast; };
#if HB_STM_PARITY_STAGES
template <> struct StageArith<ArParity> { using type = ArParity; };
#endif

template <class AR, int MODE>
__global__ void __launch_bounds__(256, 1) k_dop853_stm(const StmParams p)
{
    const int lane = threadIdx.x & 31;
    const int role = lane & 7;
    const unsigned gmask = 0xFFu << (lane & 24);
    using AS = typename StageArith<AR>::type;             // arithmetic of stages / error sums / dense output
    const StmRhs<AS> rhs{p, gmask, role == 6};
    double y[6], yh[6], k[13][6];
    double t = 0.0, h = 0.0, err_prev = -1.0, tf = 0.0;
    long long idx = -1, attempts = 0;
    int nacc = 0, nrej = 0, cursor = 0;
    bool have = false, exhausted = false;

    for (;;) {
        if (!have && !exhausted) {
            long long got = 0;
            if (role == 0) got = hb_fetch_index(p.ws);
            idx = __shfl_sync(gmask, got, 0, 8);
            if (idx < p.n) {
                // PHI0 = [I6 row-major, x0]  (rtbp.py:316-318)
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    double v = (role == i) ? 1.0 : 0.0;
                    if (role == 6) v = p.x0[(long long)i * p.n + idx];
                    if (role == 7) v = 0.0;
                    y[i] = v;
                }
                rhs(y, k[0]);
                t = p.t0;
                tf = p.tf_arr ? p.tf_arr[idx] : p.tf;
                if (MODE == SMODE_DENSE) {
                    const double *te = p.t_eval + (p.t_eval_per_traj ? idx * (long long)p.m : 0);
                    t = te[0];
                    tf = te[p.m - 1];
                }
                // initial step (rk.py:2445-2448) with the 42-component norms reduced over the group
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int d = 0; d < 6; ++d) {
                    const double sc = AR::madd(p.rtol, fabs(y[d]), p.atol);
                    const double a = AR::div(y[d], sc), b = AR::div(k[0][d], sc);
                    s0 = fma(a, a, s0);
                    s1 = fma(b, b, s1);
                }
                s0 = group_sum(s0, gmask);
                s1 = group_sum(s1, gmask);
                const double sq = AR::sqrt(42.0);
                const double d0 = AR::div(AR::sqrt(s0), sq), d1 = AR::div(AR::sqrt(s1), sq);
                h = (d0 < 1.0e-5 || d1 < 1.0e-5) ? 1.0e-6 : AR::div(AR::mul(0.01, d0), d1);
                if (h > p.max_step) h = p.max_step;
                if (h < p.min_step) h = p.min_step;
                err_prev = -1.0;
                nacc = 0; nrej = 0; cursor = 0; attempts = 0;
                have = true;
                if (!((t - tf) < 0.0)) {
                    if (role < 7) {
                        if (MODE == SMODE_FINAL) {
#pragma unroll
                            for (int i = 0; i < 6; ++i)
                                p.phi_out[idx * 42 + (role == 6 ? 36 + i : 6 * i + role)] = y[i];
                        } else {
                            for (int c = 0; c < p.m; ++c)
#pragma unroll
                                for (int i = 0; i < 6; ++i)
                                    p.dense_out[(idx * (long long)p.m + c) * 42 + (role == 6 ? 36 + i : 6 * i + role)] = y[i];
                        }
                    }
                    if (role == 0) { p.nacc[idx] = 0; p.nrej[idx] = 0; p.status[idx] = HB_TRAJ_OK; }
                    have = false;
                }
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(0xffffffffu, !have && exhausted)) break;
        if (!have) continue;

        h = hb_clamp_step(h, p.max_step, p.min_step);
        if (AR::add(t, h) > tf) h = fabs(AR::sub(tf, t));
        dop853_stages<AS>(y, k, h, yh, rhs);
        double n5 = 0.0, n3 = 0.0;
        dop853_err_sums<AS>(y, yh, k, h, p.rtol, p.atol, n5, n3);
        n5 = group_sum(n5, gmask);
        n3 = group_sum(n3, gmask);
        const double err = dop853_err_norm<AS>(n5, n3, h, 42.0);
        ++attempts;
        int fin = -1;

        // both branches, one logarithm: err_prev ** alpha is carried from the step where err_prev was the current error
        // (hb_pi_factor_carried, same bits as hb_pi_factor; in the parity build `err_prev` holds that power)
        const double h_factor = hb_pi_factor_carried<AR>(err, err_prev, err <= 1.0, 8.0);
        if (err <= 1.0) {
            const double t_new = AR::add(t, h);
            ++nacc;
            const bool last = !((t_new - tf) < 0.0);
            if (MODE == SMODE_DENSE) {
                const double *te = p.t_eval + (p.t_eval_per_traj ? idx * (long long)p.m : 0);
                if (cursor < p.m && (last || te[cursor] < t_new)) {
                    const double hseg = AR::sub(t_new, t);
                    double F[7][6], yo[6];
                    if (hseg != 0.0) dense_cache<AS>(y, yh, hseg, k, F, rhs);
                    while (cursor < p.m) {
                        const double tq = te[cursor];
                        if (!(last || tq < t_new)) break;
                        if (hseg == 0.0) {
#pragma unroll
                            for (int d = 0; d < 6; ++d) yo[d] = y[d];
                        } else {
                            dense_eval<AS>(y, F, AR::div(AR::sub(tq, t), hseg), yo);
                        }
                        if (role < 7) {
                            double *o = p.dense_out + (idx * (long long)p.m + cursor) * 42;
#pragma unroll
                            for (int i = 0; i < 6; ++i) o[role == 6 ? 36 + i : 6 * i + role] = yo[i];
                        }
                        ++cursor;
                    }
                }
                if (last) fin = HB_TRAJ_OK;
            } else if (last) {
                const double hseg = AR::sub(t_new, t);
                double yo[6];
                if (hseg == 0.0) {
#pragma unroll
                    for (int d = 0; d < 6; ++d) yo[d] = y[d];
                } else {
                    const double x = AR::div(AR::sub(tf, t), hseg);
                    if (x == 1.0) {
#pragma unroll
                        for (int d = 0; d < 6; ++d) yo[d] = AR::add(AR::sub(yh[d], y[d]), y[d]);
                    } else {
                        double F[7][6];
                        dense_cache<AS>(y, yh, hseg, k, F, rhs);
                        dense_eval<AS>(y, F, x, yo);
                    }
                }
                if (role < 7) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) p.phi_out[idx * 42 + (role == 6 ? 36 + i : 6 * i + role)] = yo[i];
                }
                fin = HB_TRAJ_OK;
            }
            t = t_new;
#pragma unroll
            for (int d = 0; d < 6; ++d) { y[d] = yh[d]; k[0][d] = k[12][d]; }
            h = AR::mul(h, h_factor);
        } else {
            ++nrej;
            h = AR::mul(h, h_factor);
            h = hb_clamp_step(h, p.max_step, p.min_step);
        }
        if (fin < 0) {
            if (!(h == h) || !(err == err)) fin = HB_TRAJ_NONFINITE;
            else if (attempts >= p.max_attempts) fin = HB_TRAJ_MAXSTEPS;
            if (fin >= 0 && MODE == SMODE_FINAL && role < 7) {
#pragma unroll
                for (int i = 0; i < 6; ++i) p.phi_out[idx * 42 + (role == 6 ? 36 + i : 6 * i + role)] = y[i];
            }
        }
        if (fin >= 0) {
            if (role == 0) { p.nacc[idx] = nacc; p.nrej[idx] = nrej; p.status[idx] = fin; }
            have = false;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The reference's other integrators on the 42-state system (_compute_stm(..., method=, order=) -> _propagate_dynsys,
// rtbp.py:258-340): adaptive RK45 (rk45_step_jit_kernel rk.py:842-898, _integrate_rk45 :1269-1399, dense output
// :971-1034) and fixed-step RK4 / RK6 / RK8 over the grid (rk_embedded_step_jit_kernel :155-215, _integrate_fixed_rk
// :533-588).  Same lane-group mapping and vector field as k_dop853_stm; the stage / error / dense-output machinery is the
// 6-state kernels' (hb_rkgen.cuh, hb_rk45.cuh) on each lane's column, the error norm is reduced over the group.
// API-parity kernels (tolerance statement like every 42-state result), not tuned.
// ---------------------------------------------------------------------------------------------------------------------
template <class AR, int MODE>
__global__ void __launch_bounds__(256, 1) k_rk45_stm(const StmParams p)
{
    const int lane = threadIdx.x & 31;
    const int role = lane & 7;
    const unsigned gmask = 0xFFu << (lane & 24);
    using AS = typename StageArith<AR>::type;
    const StmRhs<AS> rhs{p, gmask, role == 6};
    double y[6], yh[6], k[6][6], k6[6];
    double t = 0.0, h = 0.0, err_prev = -1.0, tf = 0.0;
    long long idx = -1, attempts = 0;
    int nacc = 0, nrej = 0, cursor = 0;
    bool have = false, exhausted = false;
    auto put = [&](double *row, const double (&v)[6]) {
        if (role < 7) {
#pragma unroll
            for (int i = 0; i < 6; ++i) row[role == 6 ? 36 + i : 6 * i + role] = v[i];
        }
    };
    for (;;) {
        if (!have && !exhausted) {
            long long got = 0;
            if (role == 0) got = hb_fetch_index(p.ws);
            idx = __shfl_sync(gmask, got, 0, 8);
            if (idx < p.n) {
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    double v = (role == i) ? 1.0 : 0.0;
                    if (role == 6) v = p.x0[(long long)i * p.n + idx];
                    if (role == 7) v = 0.0;
                    y[i] = v;
                }
                rhs(y, k[0]);
                t = p.t0;
                tf = p.tf_arr ? p.tf_arr[idx] : p.tf;
                if (MODE == SMODE_DENSE) {
                    const double *te = p.t_eval + (p.t_eval_per_traj ? idx * (long long)p.m : 0);
                    t = te[0];
                    tf = te[p.m - 1];
                }
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int d = 0; d < 6; ++d) {
                    const double sc = AR::madd(p.rtol, fabs(y[d]), p.atol);
                    const double a = AR::div(y[d], sc), b = AR::div(k[0][d], sc);
                    s0 = fma(a, a, s0);
                    s1 = fma(b, b, s1);
                }
                s0 = group_sum(s0, gmask);
                s1 = group_sum(s1, gmask);
                const double sq = AR::sqrt(42.0);
                const double d0 = AR::div(AR::sqrt(s0), sq), d1 = AR::div(AR::sqrt(s1), sq);
                h = (d0 < 1.0e-5 || d1 < 1.0e-5) ? 1.0e-6 : AR::div(AR::mul(0.01, d0), d1);
                if (h > p.max_step) h = p.max_step;
                if (h < p.min_step) h = p.min_step;
                err_prev = -1.0;
                nacc = 0; nrej = 0; cursor = 0; attempts = 0;
                have = true;
                if (!((t - tf) < 0.0)) {
                    if (MODE == SMODE_FINAL) put(p.phi_out + idx * 42, y);
                    else for (int c = 0; c < p.m; ++c) put(p.dense_out + (idx * (long long)p.m + c) * 42, y);
                    if (role == 0) { p.nacc[idx] = 0; p.nrej[idx] = 0; p.status[idx] = HB_TRAJ_OK; }
                    have = false;
                }
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(0xffffffffu, !have && exhausted)) break;
        if (!have) continue;

        h = hb_clamp_step(h, p.max_step, p.min_step);
        if (AR::add(t, h) > tf) h = fabs(AR::sub(tf, t));
        g_run_stages<AS, Tab45, StmRhs<AS>, 1>(rhs, y, k, h);
#pragma unroll
        for (int d = 0; d < 6; ++d) yh[d] = y[d];
        g_high_acc<AS, Tab45, 0>(yh, k, h);
        rhs(yh, k6);
        double ev[6], ssq = 0.0;
#pragma unroll
        for (int d = 0; d < 6; ++d) ev[d] = 0.0;
        rk45_err_acc<AS, 0>(ev, k, k6, h);
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            const double sc = AS::madd(p.rtol, fmax(fabs(y[d]), fabs(yh[d])), p.atol);
            const double e = AS::div(ev[d], sc);
            ssq = fma(e, e, ssq);
        }
        ssq = group_sum(ssq, gmask);
        const double err = AR::div(AR::sqrt(ssq), AR::sqrt(42.0));          // norm(err / scale) / sqrt(n), rk.py:1333
        ++attempts;
        int fin = -1;
        const double h_factor = hb_pi_factor<AR>(err, err_prev, err <= 1.0, 5.0);
        if (err <= 1.0) {
            const double t_new = AR::add(t, h);
            ++nacc;
            const bool last = !((t_new - tf) < 0.0);
            if (MODE == SMODE_DENSE || last) {
                const double hseg = AR::sub(t_new, t);
                double Q[6][4] = {}, yo[6];
                rk45_q_acc<AS, 0, 0>(Q, k, k6);
                if (MODE == SMODE_DENSE) {
                    const double *te = p.t_eval + (p.t_eval_per_traj ? idx * (long long)p.m : 0);
                    while (cursor < p.m) {
                        const double tq = te[cursor];
                        if (!(last || tq < t_new)) break;
                        rk45_eval<AS>(y, Q, AR::div(AR::sub(tq, t), hseg), hseg, yo);
                        put(p.dense_out + (idx * (long long)p.m + cursor) * 42, yo);
                        ++cursor;
                    }
                } else {
                    rk45_eval<AS>(y, Q, AR::div(AR::sub(tf, t), hseg), hseg, yo);
                    put(p.phi_out + idx * 42, yo);
                }
                if (last) fin = HB_TRAJ_OK;
            }
            t = t_new;
#pragma unroll
            for (int d = 0; d < 6; ++d) { y[d] = yh[d]; k[0][d] = k6[d]; }
            h = AR::mul(h, h_factor);
            err_prev = err;
        } else {
            ++nrej;
            h = AR::mul(h, h_factor);
            h = hb_clamp_step(h, p.max_step, p.min_step);
        }
        if (fin < 0) {
            if (!(h == h) || !(err == err)) fin = HB_TRAJ_NONFINITE;
            else if (attempts >= p.max_attempts) fin = HB_TRAJ_MAXSTEPS;
            if (fin >= 0) {
                if (MODE == SMODE_FINAL) put(p.phi_out + idx * 42, y);
                else for (; cursor < p.m; ++cursor) put(p.dense_out + (idx * (long long)p.m + cursor) * 42, y);
            }
        }
        if (fin >= 0) {
            if (role == 0) { p.nacc[idx] = nacc; p.nrej[idx] = nrej; p.status[idx] = fin; }
            have = false;
        }
    }
}

// Fixed-step RK4 / RK6 / RK8: one step per grid interval -- t_eval[m] (DENSE) or linspace(t0, tf, n_fixed + 1) (FINAL).
// Every group of a batch takes the same number of steps, so the plain per-group loop stays convergent.
template <class AR, class TAB, int MODE>
__global__ void __launch_bounds__(256, 1) k_rkfixed_stm(const StmParams p, const int n_fixed)
{
    const int lane = threadIdx.x & 31;
    const int role = lane & 7;
    const unsigned gmask = 0xFFu << (lane & 24);
    using AS = typename StageArith<AR>::type;
    const StmRhs<AS> rhs{p, gmask, role == 6};
    const int npts = (MODE == SMODE_DENSE) ? p.m : n_fixed + 1;
    for (;;) {
        long long got = 0;
        if (role == 0) got = hb_fetch_index(p.ws);
        const long long idx = __shfl_sync(gmask, got, 0, 8);
        if (idx >= p.n) break;
        double y[6], yn[6], k[TAB::S][6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double v = (role == i) ? 1.0 : 0.0;
            if (role == 6) v = p.x0[(long long)i * p.n + idx];
            if (role == 7) v = 0.0;
            y[i] = v;
        }
        const double *te = (MODE == SMODE_DENSE) ? p.t_eval + (p.t_eval_per_traj ? idx * (long long)p.m : 0) : nullptr;
        const double tf = p.tf_arr ? p.tf_arr[idx] : p.tf;
        const double lin_step = (n_fixed > 0) ? AR::div(AR::sub(tf, p.t0), (double)n_fixed) : 0.0;
        auto grid = [&](int i) -> double {
            if (MODE == SMODE_DENSE) return te[i];
            return (i == n_fixed) ? tf : AR::madd((double)i, lin_step, p.t0);         // numpy.linspace
        };
        auto put = [&](double *row, const double (&v)[6]) {
            if (role < 7) {
#pragma unroll
                for (int i = 0; i < 6; ++i) row[role == 6 ? 36 + i : 6 * i + role] = v[i];
            }
        };
        if (MODE == SMODE_DENSE) put(p.dense_out + idx * (long long)p.m * 42, y);
        int fin = HB_TRAJ_OK;
        for (int i = 0; i + 1 < npts; ++i) {
            const double tn = grid(i);
            const double h = AR::sub(grid(i + 1), tn);
            rhs(y, k[0]);
            g_run_stages<AS, TAB, StmRhs<AS>, 1>(rhs, y, k, h);
#pragma unroll
            for (int d = 0; d < 6; ++d) yn[d] = y[d];
            g_high_acc<AS, TAB, 0>(yn, k, h);
#pragma unroll
            for (int d = 0; d < 6; ++d) y[d] = yn[d];
            if (MODE == SMODE_DENSE) put(p.dense_out + (idx * (long long)p.m + i + 1) * 42, y);
            const unsigned bad = __ballot_sync(gmask, !(y[0] == y[0])) & gmask;
            if (bad) fin = HB_TRAJ_NONFINITE;
        }
        if (MODE == SMODE_FINAL) put(p.phi_out + idx * 42, y);
        if (role == 0) { p.nacc[idx] = npts - 1; p.nrej[idx] = 0; p.status[idx] = fin; }
    }
}

int fill(const hb_cr3bp *sys, const hb_integ *integ, StmParams &p)
{
    if (!sys || !integ) return HB_ERR_BADARG;
    if (integ->method != HB_DOP853 && integ->method != HB_RK45 && integ->method != HB_RK4 && integ->method != HB_RK6 &&
        integ->method != HB_RK8)
        return HB_ERR_UNSUPPORTED;
    if (integ->arith != HB_ARITH_PARITY && integ->arith != HB_ARITH_FAST) return HB_ERR_BADARG;
    p.mu = sys->mu;
    p.om = 1.0 - sys->mu;
    p.neg_phi = 0; p.neg_state = 0;
    if (sys->fwd < 0) {
        if (sys->flip_lo < 0 || (sys->flip_lo == 0 && sys->flip_hi == 42)) { p.neg_phi = 1; p.neg_state = 1; }
        else if (sys->flip_lo == 36 && sys->flip_hi == 42) p.neg_state = 1;      // _compute_stm, rtbp.py:329
        else if (sys->flip_lo == 0 && sys->flip_hi == 36) p.neg_phi = 1;
        else return HB_ERR_UNSUPPORTED;
    }
    p.rtol = integ->rtol; p.atol = integ->atol;
    p.max_step = integ->max_step; p.min_step = integ->min_step;
    p.max_attempts = integ->max_attempts > 0 ? integ->max_attempts : 2147483647LL;
    return HB_OK;
}

template <class AR, int MODE>
int launch_rk(const StmParams &p, int method, int n_fixed, unsigned blocks, cudaStream_t st)
{
    switch (method) {
    case HB_RK45: k_rk45_stm<AR, MODE><<<blocks, 256, 0, st>>>(p); break;
    case HB_RK4: k_rkfixed_stm<AR, TabRK4, MODE><<<blocks, 256, 0, st>>>(p, n_fixed); break;
    case HB_RK6: k_rkfixed_stm<AR, TabRK6, MODE><<<blocks, 256, 0, st>>>(p, n_fixed); break;
    case HB_RK8: k_rkfixed_stm<AR, TabRK8, MODE><<<blocks, 256, 0, st>>>(p, n_fixed); break;
    default: return HB_ERR_UNSUPPORTED;
    }
    return HB_OK;
}

template <int MODE>
int launch(const StmParams &p, int arith, cudaStream_t st, int method = HB_DOP853, int n_fixed = 0)
{
    if (method != HB_DOP853 && method != HB_RK45 && MODE == SMODE_FINAL && n_fixed < 1) return HB_ERR_BADARG;
    HB_CUDA_TRY(cudaMemsetAsync(p.ws, 0, sizeof(HbWorkspace), st));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256;                         // 32 trajectories per CTA
    long long blocks = (p.n * 8 + threads - 1) / threads;
    if (blocks > sms) blocks = sms;                  // persistent: one CTA per SM
    if (blocks < 1) blocks = 1;
    if (method != HB_DOP853) {
        const int rc = (arith == HB_ARITH_PARITY) ? launch_rk<ArParity, MODE>(p, method, n_fixed, (unsigned)blocks, st)
                                                  : launch_rk<ArFast, MODE>(p, method, n_fixed, (unsigned)blocks, st);
        if (rc != HB_OK) return rc;
    } else if (arith == HB_ARITH_PARITY) k_dop853_stm<ArParity, MODE><<<(unsigned)blocks, threads, 0, st>>>(p);
    else k_dop853_stm<ArFast, MODE><<<(unsigned)blocks, threads, 0, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

}  // namespace

extern "C" int hb_cr3bp_stm(const hb_cr3bp *sys, const hb_integ *integ, int64_t n, const double *x0_soa, double t0,
                            double tf, const double *tf_per_traj, double *phi_out, int32_t *n_acc, int32_t *n_rej,
                            int32_t *status, void *workspace, void *stream)
{
    StmParams p{};
    int rc = fill(sys, integ, p);
    if (rc != HB_OK) return rc;
    if (n < 0 || !workspace || (n > 0 && (!x0_soa || !phi_out || !n_acc || !n_rej || !status))) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    p.n = n; p.x0 = x0_soa; p.t0 = t0; p.tf = tf; p.tf_arr = tf_per_traj; p.phi_out = phi_out;
    p.nacc = n_acc; p.nrej = n_rej; p.status = status; p.ws = (HbWorkspace *)workspace;
    return launch<SMODE_FINAL>(p, integ->arith, (cudaStream_t)stream, integ->method, integ->n_fixed_steps);
}

extern "C" int hb_cr3bp_stm_dense(const hb_cr3bp *sys, const hb_integ *integ, int64_t n, const double *x0_soa,
                                  const double *t_eval, int32_t m, int32_t t_eval_per_traj, double *phi_dense,
                                  int32_t *n_acc, int32_t *n_rej, int32_t *status, void *workspace, void *stream)
{
    StmParams p{};
    int rc = fill(sys, integ, p);
    if (rc != HB_OK) return rc;
    if (n < 0 || m < 2 || !workspace || !t_eval || (n > 0 && (!x0_soa || !phi_dense || !n_acc || !n_rej || !status)))
        return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    p.n = n; p.x0 = x0_soa; p.t_eval = t_eval; p.m = m; p.t_eval_per_traj = t_eval_per_traj ? 1 : 0;
    p.dense_out = phi_dense; p.nacc = n_acc; p.nrej = n_rej; p.status = status; p.ws = (HbWorkspace *)workspace;
    return launch<SMODE_DENSE>(p, integ->arith, (cudaStream_t)stream, integ->method, integ->n_fixed_steps);
}
