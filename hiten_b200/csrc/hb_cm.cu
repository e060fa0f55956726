// hb_cm.cu -- centre-manifold Poincare map: fixed-step RK4/6/8 or Tao extended-phase-space symplectic
// steps on a polynomial Hamiltonian, with in-loop section detection and cubic-Hermite hit interpolation.
//
// One seed per thread.  The polynomial gradient is evaluated from a sparse real term table
// {coef, 6 exponents, degree} held in shared memory (uniform across the block -> broadcast reads) and a
// per-thread power table x_v^e, also in shared memory, laid out [v][e][thread] (conflict-free).  Terms are
// visited in the reference's order (degree ascending, packed index ascending) and summed per degree
// first, so the parity variant reproduces its rounding.
//
// Reference routines (paths relative to hiten/):
//   _poly_evaluate / _polynomial_evaluate   algorithms/polynomial/algebra.py:403-461, operations.py:551-583
//   _hamiltonian_rhs                        algorithms/dynamics/hamiltonian.py:35-90
//   _eval_dH_dQ / _eval_dH_dP               algorithms/integrators/symplectic.py:105-178
//   _phi_H_a/_b/_phi_omega_H_c/_recursive_update_poly/_integrate_symplectic   symplectic.py:370-653
//   _integrate_rk_ham, _detect_crossing, _poincare_step, _poincare_map
//                                           algorithms/poincare/centermanifold/backend.py:35-382
//   _hermite_scalar                         algorithms/poincare/utils.py:54-97
#include "hb_cr3bp_common.cuh"
#include "hb_rkgen.cuh"
#include "hb_rk45.cuh"

#include <cmath>

namespace {

constexpr int CM_BLOCK = 128;

struct TermMeta {          // 16 bytes
    double coef;
    unsigned long long ex; // exponents of (q1,q2,q3,p1,p2,p3) in bytes 0..5, degree in bytes 6..7
};

struct CmParams {
    hb_cm_opts o;
    long long n;
    const double *seeds;   // [n][4] (q2,p2,q3,p3)
    int *flags;
    double *out;           // [n][4]
    double *t_out;
    HbWorkspace *ws;
    const TermMeta *terms; // device, all partials back to back
    int ptr[7];
    int n_terms;
    int D;                 // max exponent
};

// ---- polynomial gradient -----------------------------------------------------------------------
// pw points at this thread's column of the shared power table: pw[(v*(D+1)+e)*CM_BLOCK].
template <class AR, class PRM>
__device__ __noinline__ void cm_grad(const PRM &p, double *pw, const TermMeta *terms, const double (&pt)[6],
                                     double (&g)[6])
{
    const int D1 = p.D + 1;
#pragma unroll
    for (int v = 0; v < 6; ++v) {
        double w = 1.0;
        pw[(v * D1) * CM_BLOCK] = 1.0;
        for (int e = 1; e < D1; ++e) {
            w = AR::mul(w, pt[v]);
            pw[(v * D1 + e) * CM_BLOCK] = w;
        }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        double total = 0.0, acc = 0.0;
        int dcur = -1;
        for (int i = p.ptr[q]; i < p.ptr[q + 1]; ++i) {
            const TermMeta tm = terms[i];
            const int d = (int)(tm.ex >> 48);
            if (d != dcur) {
                if (dcur >= 0) total = AR::add(total, acc);
                acc = 0.0;
                dcur = d;
            }
            double t = 1.0;
            bool first = true;
#pragma unroll
            for (int v = 0; v < 6; ++v) {
                const int e = (int)((tm.ex >> (8 * v)) & 0xffu);
                if (e) {
                    const double f = pw[(v * D1 + e) * CM_BLOCK];
                    t = first ? f : AR::mul(t, f);
                    first = false;
                }
            }
            acc = AR::madd(tm.coef, t, acc);
        }
        if (dcur >= 0) total = AR::add(total, acc);
        g[q] = total;
    }
}

// rhs = [dH/dP, -dH/dQ]  (hamiltonian.py:80-90)
template <class AR, class PRM>
HB_DEV void cm_rhs(const PRM &p, double *pw, const TermMeta *terms, const double (&y)[6], double (&dy)[6])
{
    double g[6];
    cm_grad<AR>(p, pw, terms, y, g);
    dy[0] = g[3]; dy[1] = g[4]; dy[2] = g[5];
    dy[3] = -g[0]; dy[4] = -g[1]; dy[5] = -g[2];
}

// One _recursive_update_poly call (symplectic.py:509-560) flattened into its n_sub order-2 kernels
// phi_a(ts/2) phi_b(ts/2) phi_c(ts) phi_b(ts/2) phi_a(ts/2) on the extended state [Q, P, X, Y]; ts / cos / sin per kernel
// come from the host (hb_cm_prepare / hb_tao_grid_prepare: libm, as the reference evaluates them).
template <class AR, class PRM>
HB_DEV void tao_update(const PRM &p, double *pw, const TermMeta *terms, const double *sub_ts, const double *sub_cos,
                       const double *sub_sin, int n_sub, double (&Q)[3], double (&P)[3], double (&X)[3], double (&Y)[3])
{
    double g[6], pt[6];
    for (int j = 0; j < n_sub; ++j) {
        const double ts = sub_ts[j], hd = AR::mul(0.5, ts);
        const double c = sub_cos[j], s = sub_sin[j];
#pragma unroll 1
        for (int ph = 0; ph < 5; ++ph) {
            if (ph == 2) {                   // phi_omega_H_c (symplectic.py:486-506)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double qpx = AR::add(Q[i], X[i]), qmx = AR::sub(Q[i], X[i]);
                    const double ppy = AR::add(P[i], Y[i]), pmy = AR::sub(P[i], Y[i]);
                    Q[i] = AR::mul(0.5, AR::add(AR::add(qpx, AR::mul(c, qmx)), AR::mul(s, pmy)));
                    P[i] = AR::mul(0.5, AR::add(AR::sub(ppy, AR::mul(s, qmx)), AR::mul(c, pmy)));
                    X[i] = AR::mul(0.5, AR::sub(AR::sub(qpx, AR::mul(c, qmx)), AR::mul(s, pmy)));
                    Y[i] = AR::mul(0.5, AR::sub(AR::add(ppy, AR::mul(s, qmx)), AR::mul(c, pmy)));
                }
            } else if (ph == 0 || ph == 4) { // phi_H_a: gradient at (Q, Y); P -= d*dHdQ, X += d*dHdP
#pragma unroll
                for (int i = 0; i < 3; ++i) { pt[i] = Q[i]; pt[3 + i] = Y[i]; }
                cm_grad<AR>(p, pw, terms, pt, g);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    P[i] = AR::sub(P[i], AR::mul(hd, g[i]));
                    X[i] = AR::add(X[i], AR::mul(hd, g[3 + i]));
                }
            } else {                         // phi_H_b: gradient at (X, P); Q += d*dHdP, Y -= d*dHdQ
#pragma unroll
                for (int i = 0; i < 3; ++i) { pt[i] = X[i]; pt[3 + i] = P[i]; }
                cm_grad<AR>(p, pw, terms, pt, g);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    Q[i] = AR::add(Q[i], AR::mul(hd, g[3 + i]));
                    Y[i] = AR::sub(Y[i], AR::mul(hd, g[i]));
                }
            }
        }
    }
}

template <class AR>
struct CmRhs {
    const CmParams &p;
    double *pw;
    const TermMeta *terms;
    HB_DEV void operator()(const double (&y)[6], double (&dy)[6]) const { cm_rhs<AR>(p, pw, terms, y, dy); }
};

// _hermite_scalar (poincare/utils.py:93-97)
template <class AR>
HB_DEV double hermite_scalar(double s, double y0, double y1, double dy0, double dy1, double dt)
{
    const double oms = AR::sub(1.0, s);
    const double oms2 = AR::mul(oms, oms), s2 = AR::mul(s, s);
    const double h00 = AR::mul(AR::add(1.0, AR::mul(2.0, s)), oms2);
    const double h10 = AR::mul(s, oms2);
    const double h01 = AR::mul(s2, AR::sub(3.0, AR::mul(2.0, s)));
    const double h11 = AR::mul(s2, AR::sub(s, 1.0));
    return AR::add(AR::add(AR::add(AR::mul(h00, y0), AR::mul(AR::mul(h10, dy0), dt)), AR::mul(h01, y1)),
                   AR::mul(AR::mul(h11, dy1), dt));
}

// TAB = void selects the Tao symplectic integrator.
template <class AR, class TAB>
__global__ void __launch_bounds__(CM_BLOCK) k_cm_map(const CmParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    TermMeta *terms = reinterpret_cast<TermMeta *>(smem);
    double *pw_all = reinterpret_cast<double *>(smem + (((size_t)p.n_terms * sizeof(TermMeta) + 15) & ~(size_t)15));
    for (int i = threadIdx.x; i < p.n_terms; i += CM_BLOCK) terms[i] = p.terms[i];
    __syncthreads();
    double *pw = pw_all + threadIdx.x;
    constexpr bool TAO = std::is_same<TAB, void>::value;
    const int sec = p.o.section;                       // 0 q2, 1 p2, 2 q3, 3 p3
    const double dt = p.o.dt;

    for (;;) {
        const long long idx = hb_fetch_index(p.ws);
        if (idx >= p.n) break;
        const double *sd = p.seeds + idx * 4;
        double so[6] = {0.0, sd[0], sd[2], 0.0, sd[1], sd[3]}, sn[6], rn[6], ro[6];
        if (!TAO) cm_rhs<AR>(p, pw, terms, so, ro);     // stage 0 of the first step; later steps reuse rn
        double elapsed = 0.0;
        int flag = 0;
        double o4[4] = {0.0, 0.0, 0.0, 0.0}, tc = 0.0;
        for (int it = 0; it < p.o.max_steps; ++it) {
            if constexpr (TAO) {
                // q_ext = [Q, P, X = Q, Y = P] rebuilt every dt (backend.py:289-291, symplectic.py:636-640)
                double Q[3] = {so[0], so[1], so[2]}, P[3] = {so[3], so[4], so[5]};
                double X[3] = {so[0], so[1], so[2]}, Y[3] = {so[3], so[4], so[5]};
                tao_update<AR>(p, pw, terms, p.o.sub_ts, p.o.sub_cos, p.o.sub_sin, p.o.n_sub, Q, P, X, Y);
#pragma unroll
                for (int i = 0; i < 3; ++i) { sn[i] = Q[i]; sn[3 + i] = P[i]; }
            } else {
                double k[TAB::S][6];
#pragma unroll
                for (int d = 0; d < 6; ++d) k[0][d] = ro[d];
                g_run_stages<AR, TAB, CmRhs<AR>, 1>(CmRhs<AR>{p, pw, terms}, so, k, dt);
#pragma unroll
                for (int d = 0; d < 6; ++d) sn[d] = so[d];
                g_high_acc<AR, TAB, 0>(sn, k, dt);
            }
            cm_rhs<AR>(p, pw, terms, sn, rn);
            // _detect_crossing (backend.py:58-89)
            const double f_old = (sec == 0) ? so[1] : (sec == 1) ? so[4] : (sec == 2) ? so[2] : so[5];
            const double f_new = (sec == 0) ? sn[1] : (sec == 1) ? sn[4] : (sec == 2) ? sn[2] : sn[5];
            bool crossed = false;
            if (!(AR::mul(f_old, f_new) >= 0.0)) {
                crossed = (sec == 2) ? (sn[5] > 0.0) : (sec == 0) ? (sn[4] > 0.0) : (sec == 3) ? (rn[2] > 0.0) : (rn[1] > 0.0);
            }
            if (crossed) {
                const double alpha = AR::div(f_old, AR::sub(f_old, f_new));
                if (TAO) cm_rhs<AR>(p, pw, terms, so, ro);
                o4[0] = hermite_scalar<AR>(alpha, so[1], sn[1], ro[1], rn[1], dt);
                o4[1] = hermite_scalar<AR>(alpha, so[4], sn[4], ro[4], rn[4], dt);
                o4[2] = hermite_scalar<AR>(alpha, so[2], sn[2], ro[2], rn[2], dt);
                o4[3] = hermite_scalar<AR>(alpha, so[5], sn[5], ro[5], rn[5], dt);
                tc = AR::add(elapsed, AR::mul(alpha, dt));
                flag = 1;
                break;
            }
#pragma unroll
            for (int d = 0; d < 6; ++d) { so[d] = sn[d]; ro[d] = rn[d]; }
            elapsed = AR::add(elapsed, dt);
        }
        p.flags[idx] = flag;
        p.t_out[idx] = tc;
        double *o = p.out + idx * 4;
        o[0] = o4[0]; o[1] = o4[1]; o[2] = o4[2]; o[3] = o4[3];
    }
}

template <class AR, class TAB>
int launch_cm(const CmParams &p, size_t smem, cudaStream_t st)
{
    auto kern = k_cm_map<AR, TAB>;
    HB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CM_BLOCK, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (p.n + CM_BLOCK - 1) / CM_BLOCK;
    const long long cap = (long long)sms * per_sm;       // persistent: resident CTAs, seeds pulled from the queue
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, CM_BLOCK, smem, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

template <class AR>
int dispatch(const CmParams &p, size_t smem, cudaStream_t st)
{
    if (p.o.method == HB_SYMPLECTIC) return launch_cm<AR, void>(p, smem, st);
    if (p.o.method == HB_RK4) return launch_cm<AR, TabRK4>(p, smem, st);
    if (p.o.method == HB_RK6) return launch_cm<AR, TabRK6>(p, smem, st);
    if (p.o.method == HB_RK8) return launch_cm<AR, TabRK8>(p, smem, st);
    return HB_ERR_UNSUPPORTED;
}

// _recursive_update_poly's schedule flattened (symplectic.py:543-560): order-2 kernels with their time steps
void tao_schedule(double ts, int order, hb_cm_opts *o)
{
    if (order == 2) {
        if (o->n_sub < HB_MAX_TAO_SUBSTEPS) o->sub_ts[o->n_sub] = ts;
        o->n_sub++;
    } else {
        const double gamma = 1.0 / (2.0 - std::pow(2.0, 1.0 / ((double)order + 1.0)));
        tao_schedule(gamma * ts, order - 2, o);
        tao_schedule((1.0 - 2.0 * gamma) * ts, order - 2, o);
        tao_schedule(gamma * ts, order - 2, o);
    }
}


// ---- _ExtendedSymplectic.integrate: the Tao integrator over a time grid --------------------------------------------
// _integrate_symplectic (symplectic.py:564-653): the extended state [Q,P,X,Y] is built ONCE and carried across the grid
// (the centre-manifold map above rebuilds it every dt); the step of grid interval i is np.diff(t_values)[i], so omega,
// the triple-jump sub-steps and their cos/sin are per-interval values (hb_tao_grid_prepare evaluates them with the host
// libm).  _integrate_symplectic_until_event (:657-782): the same loop with a terminal plane event, refined by bisection
// on the cubic Hermite interpolant of the step (_hermite_refine_event_symplectic :282-367).
struct SympParams {
    long long n;
    const double *y0;      // [n][6] = [Q, P]
    int m, n_sub;
    const double *tab;     // [m-1][3][n_sub]: sub_ts, cos, sin of every grid interval
    const double *t_vals;  // [m] signed grid (event mode)
    double *traj;          // [n][m][6] or nullptr
    double *derivs;        // [n][m][6] or nullptr (fixed-step RK grid form: _Solution.derivatives)
    int ev_idx, ev_dir;
    double ev_off, xtol, gtol;
    int *hit;              // event mode outputs
    double *t_hit, *y_hit; // [n], [n][6]
    int *n_rows;           // trajectory rows written (i + 1 at a hit, m otherwise)
    HbWorkspace *ws;
    const TermMeta *terms;
    int ptr[7];
    int n_terms;
    int D;
};

HB_DEV double pick6c(const double (&v)[6], int i)
{
    return (i == 0) ? v[0] : (i == 1) ? v[1] : (i == 2) ? v[2] : (i == 3) ? v[3] : (i == 4) ? v[4] : v[5];
}

// _hermite_eval_dense_symplectic (symplectic.py:229-279)
template <class AR>
HB_DEV void hermite_eval6(const double (&y0)[6], const double (&f0)[6], const double (&y1)[6], const double (&f1)[6],
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

template <class AR, bool EVENT>
__global__ void __launch_bounds__(CM_BLOCK) k_symp_grid(const SympParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    TermMeta *terms = reinterpret_cast<TermMeta *>(smem);
    double *pw_all = reinterpret_cast<double *>(smem + (((size_t)p.n_terms * sizeof(TermMeta) + 15) & ~(size_t)15));
    for (int i = threadIdx.x; i < p.n_terms; i += CM_BLOCK) terms[i] = p.terms[i];
    __syncthreads();
    double *pw = pw_all + threadIdx.x;

    for (;;) {
        const long long idx = hb_fetch_index(p.ws);
        if (idx >= p.n) break;
        const double *s0 = p.y0 + idx * 6;
        double Q[3], P[3], X[3], Y[3], yo[6], fo[6], g_old = 0.0;
#pragma unroll
        for (int d = 0; d < 6; ++d) yo[d] = s0[d];
#pragma unroll
        for (int i = 0; i < 3; ++i) { Q[i] = X[i] = yo[i]; P[i] = Y[i] = yo[3 + i]; }
        double *rows = p.traj ? p.traj + idx * (long long)p.m * 6 : nullptr;
        if (rows) {
#pragma unroll
            for (int d = 0; d < 6; ++d) rows[d] = yo[d];
        }
        if (EVENT) {
            cm_rhs<AR>(p, pw, terms, yo, fo);            // _eval_hamiltonian_derivative (symplectic.py:181-226)
            g_old = AR::sub(pick6c(yo, p.ev_idx), p.ev_off);
        }
        int hit = 0, n_rows = p.m;
        double th = 0.0, yh[6];
        for (int i = 0; i < p.m - 1; ++i) {
            const double *tb = p.tab + (size_t)i * 3 * p.n_sub;
            tao_update<AR>(p, pw, terms, tb, tb + p.n_sub, tb + 2 * p.n_sub, p.n_sub, Q, P, X, Y);
            double yn[6] = {Q[0], Q[1], Q[2], P[0], P[1], P[2]};
            if (EVENT) {
                double fn[6];
                cm_rhs<AR>(p, pw, terms, yn, fn);
                const double g_new = AR::sub(pick6c(yn, p.ev_idx), p.ev_off);
                if (hb_event_crossed(g_old, g_new, p.ev_dir)) {
                    const double t0 = p.t_vals[i], h = AR::sub(p.t_vals[i + 1], t0);
                    double a = 0.0, b = 1.0, g_left = g_old, xh = 1.0;
                    bool done = false;
                    for (int it = 0; it < 128; ++it) {
                        const double mid = AR::mul(0.5, AR::add(a, b));
                        hermite_eval6<AR>(yo, fo, yn, fn, mid, h, yh);
                        const double g_mid = AR::sub(pick6c(yh, p.ev_idx), p.ev_off);
                        if (fabs(g_mid) <= p.gtol) { xh = mid; done = true; break; }
                        if (hb_crossed_direction(g_left, g_mid, p.ev_dir)) b = mid;
                        else { a = mid; g_left = g_mid; }
                        if (AR::mul(AR::sub(b, a), fabs(h)) <= p.xtol) break;
                    }
                    if (!done) { xh = b; hermite_eval6<AR>(yo, fo, yn, fn, b, h, yh); }
                    th = AR::add(t0, AR::mul(xh, h));
                    hit = 1;
                    n_rows = i + 1;
                    break;
                }
                g_old = g_new;
#pragma unroll
                for (int d = 0; d < 6; ++d) { yo[d] = yn[d]; fo[d] = fn[d]; }
            }
            if (rows) {
                double *o = rows + (long long)(i + 1) * 6;
#pragma unroll
                for (int d = 0; d < 6; ++d) o[d] = yn[d];
            }
        }
        if (EVENT) {
            if (!hit) {
                th = p.t_vals[p.m - 1];
#pragma unroll
                for (int d = 0; d < 6; ++d) yh[d] = yo[d];
            }
            p.hit[idx] = hit;
            p.t_hit[idx] = th;
            p.n_rows[idx] = n_rows;
#pragma unroll
            for (int d = 0; d < 6; ++d) p.y_hit[idx * 6 + d] = yh[d];
        }
    }
}

// ---- _FixedStepRK.integrate on a polynomial Hamiltonian system: the `_ham` kernels of the RK classes ---------------
// _integrate_fixed_rk_ham (rk.py:592-656: states AND derivatives on the grid, step h = t[i+1] - t[i]) and
// _integrate_fixed_rk_until_event_ham (:722-757) with _hermite_refine_in_step (:331-391); stage loop
// rk_embedded_step_ham_jit_kernel (:216-270).  The `_ham` kernels take system.rhs_params and never see a direction
// wrapper (golden vectors: tests/golden/ham_rk.npz), so there is no direction argument here.
template <class AR>
struct SympRhs {
    const SympParams &p;
    double *pw;
    const TermMeta *terms;
    HB_DEV void operator()(const double (&y)[6], double (&dy)[6]) const { cm_rhs<AR>(p, pw, terms, y, dy); }
};

template <class AR, class TAB, bool EVENT>
__global__ void __launch_bounds__(CM_BLOCK) k_ham_rk_grid(const SympParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    TermMeta *terms = reinterpret_cast<TermMeta *>(smem);
    double *pw_all = reinterpret_cast<double *>(smem + (((size_t)p.n_terms * sizeof(TermMeta) + 15) & ~(size_t)15));
    for (int i = threadIdx.x; i < p.n_terms; i += CM_BLOCK) terms[i] = p.terms[i];
    __syncthreads();
    double *pw = pw_all + threadIdx.x;
    const SympRhs<AR> rhs{p, pw, terms};

    for (;;) {
        const long long idx = hb_fetch_index(p.ws);
        if (idx >= p.n) break;
        const double *s0 = p.y0 + idx * 6;
        double y[6], fp[6], yn[6], fn[6], g_prev = 0.0;
#pragma unroll
        for (int d = 0; d < 6; ++d) y[d] = s0[d];
        rhs(y, fp);
        double *rows = p.traj ? p.traj + idx * (long long)p.m * 6 : nullptr;
        double *drows = p.derivs ? p.derivs + idx * (long long)p.m * 6 : nullptr;
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            if (rows) rows[d] = y[d];
            if (drows) drows[d] = fp[d];
        }
        if (EVENT) g_prev = AR::sub(pick6c(y, p.ev_idx), p.ev_off);
        int hit = 0, n_rows = p.m;
        double th = 0.0, yh[6];
        for (int i = 0; i < p.m - 1; ++i) {
            const double tn = p.t_vals[i], h = AR::sub(p.t_vals[i + 1], tn);
            double k[TAB::S][6];
#pragma unroll
            for (int d = 0; d < 6; ++d) k[0][d] = fp[d];
            g_run_stages<AR, TAB, SympRhs<AR>, 1>(rhs, y, k, h);
#pragma unroll
            for (int d = 0; d < 6; ++d) yn[d] = y[d];
            g_high_acc<AR, TAB, 0>(yn, k, h);
            rhs(yn, fn);
            if (EVENT) {
                const double g_new = AR::sub(pick6c(yn, p.ev_idx), p.ev_off);
                if (hb_event_crossed(g_prev, g_new, p.ev_dir)) {
                    double a = 0.0, b = 1.0, g_left = g_prev, xh = 1.0;
                    bool done = false;
                    for (int it = 0; it < 128; ++it) {
                        const double mid = AR::mul(0.5, AR::add(a, b));
                        hermite_eval6<AR>(y, fp, yn, fn, mid, h, yh);
                        const double g_mid = AR::sub(pick6c(yh, p.ev_idx), p.ev_off);
                        if (fabs(g_mid) <= p.gtol) { xh = mid; done = true; break; }
                        if (hb_crossed_direction(g_left, g_mid, p.ev_dir)) b = mid;
                        else { a = mid; g_left = g_mid; }
                        if (AR::mul(AR::sub(b, a), fabs(h)) <= p.xtol) break;
                    }
                    if (!done) { xh = b; hermite_eval6<AR>(y, fp, yn, fn, b, h, yh); }
                    th = AR::add(tn, AR::mul(xh, h));
                    hit = 1;
                    n_rows = i + 1;
                    break;
                }
                g_prev = g_new;
            }
#pragma unroll
            for (int d = 0; d < 6; ++d) { y[d] = yn[d]; fp[d] = fn[d]; }
            if (rows) {
                double *o = rows + (long long)(i + 1) * 6;
#pragma unroll
                for (int d = 0; d < 6; ++d) o[d] = y[d];
            }
            if (drows) {
                double *o = drows + (long long)(i + 1) * 6;
#pragma unroll
                for (int d = 0; d < 6; ++d) o[d] = fp[d];
            }
        }
        if (EVENT) {
            if (!hit) {
                th = p.t_vals[p.m - 1];
#pragma unroll
                for (int d = 0; d < 6; ++d) yh[d] = y[d];
            }
            p.hit[idx] = hit;
            p.t_hit[idx] = th;
            p.n_rows[idx] = n_rows;
#pragma unroll
            for (int d = 0; d < 6; ++d) p.y_hit[idx * 6 + d] = yh[d];
        }
    }
}

template <class KERN>
int launch_grid_kernel(KERN kern, const SympParams &p, size_t smem, cudaStream_t st)
{
    HB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CM_BLOCK, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (p.n + CM_BLOCK - 1) / CM_BLOCK;
    const long long cap = (long long)sms * per_sm;       // persistent: resident CTAs pull trajectories from the queue
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, CM_BLOCK, smem, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

template <class AR, bool EVENT>
int launch_ham_rk(const SympParams &p, int method, size_t smem, cudaStream_t st)
{
    if (method == HB_RK4) return launch_grid_kernel(k_ham_rk_grid<AR, TabRK4, EVENT>, p, smem, st);
    if (method == HB_RK6) return launch_grid_kernel(k_ham_rk_grid<AR, TabRK6, EVENT>, p, smem, st);
    if (method == HB_RK8) return launch_grid_kernel(k_ham_rk_grid<AR, TabRK8, EVENT>, p, smem, st);
    return HB_ERR_UNSUPPORTED;
}

// ---- AdaptiveRK on a polynomial Hamiltonian system: the `_ham` kernels of _DOP853 / _RK45 ---------------------------
// _integrate_dop853_ham (rk.py:2553-2676) / _integrate_rk45_ham (:1403-1456): the adaptive loops of hb_cr3bp.cu /
// hb_cr3bp_rk.cu with the polynomial vector field, dense output at every t_eval AND the derivative re-evaluated there
// (_Solution.derivatives); _integrate_dop853_until_event_ham (:2807-2868) / _integrate_rk45_until_event_ham (:1589-1633):
// terminal plane event refined on the method's own dense interpolant.  Same controller (utils.py), same first step.
struct HamAdParams {
    SympParams s;          // batch, grid (t_vals = t_eval), outputs traj / derivs, event fields, hit / t_hit / y_hit
    double rtol, atol, max_step, min_step, t0, tmax;
    long long max_attempts;
    int *nacc, *nrej, *status;
};

template <class AR>
HB_DEV double ham_initial_step(const double (&y)[6], const double (&f)[6], const HamAdParams &q)   // utils.py:127-157
{
    double a[6], b[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        const double sc = AR::madd(q.rtol, fabs(y[d]), q.atol);
        a[d] = AR::div(y[d], sc);
        b[d] = AR::div(f[d], sc);
    }
    const double sq = AR::sqrt(6.0);
    const double d0 = AR::div(hbc::norm2_ext6(a), sq), d1 = AR::div(hbc::norm2_ext6(b), sq);
    double h = (d0 < 1.0e-5 || d1 < 1.0e-5) ? 1.0e-6 : AR::div(AR::mul(0.01, d0), d1);
    if (h > q.max_step) h = q.max_step;
    if (h < q.min_step) h = q.min_step;
    return h;
}

// METHOD: HB_DOP853 or HB_RK45; EVENT: terminal plane event instead of the dense grid
template <class AR, int METHOD, bool EVENT>
__global__ void __launch_bounds__(CM_BLOCK) k_ham_adaptive(const HamAdParams q)
{
    const SympParams &p = q.s;
    extern __shared__ __align__(16) unsigned char smem[];
    TermMeta *terms = reinterpret_cast<TermMeta *>(smem);
    double *pw_all = reinterpret_cast<double *>(smem + (((size_t)p.n_terms * sizeof(TermMeta) + 15) & ~(size_t)15));
    for (int i = threadIdx.x; i < p.n_terms; i += CM_BLOCK) terms[i] = p.terms[i];
    __syncthreads();
    double *pw = pw_all + threadIdx.x;
    const SympRhs<AR> rhs{p, pw, terms};
    constexpr bool D8 = METHOD == HB_DOP853;
    constexpr int NK = D8 ? 13 : 6;

    for (;;) {
        const long long idx = hb_fetch_index(p.ws);
        if (idx >= p.n) break;
        const double *s0 = p.y0 + idx * 6;
        double y[6], yh[6], k[NK][6], k6[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) y[d] = s0[d];
        rhs(y, k[0]);
        double t = EVENT ? q.t0 : p.t_vals[0];
        const double tf = EVENT ? q.tmax : p.t_vals[p.m - 1];
        double h = ham_initial_step<AR>(y, k[0], q);
        double err_prev = -1.0, g_prev = 0.0, t_end = 0.0;
        if (EVENT) g_prev = AR::sub(pick6c(y, p.ev_idx), p.ev_off);
        int nacc = 0, nrej = 0, cursor = 0, fin = -1;
        long long attempts = 0;
        double *rows = p.traj ? p.traj + idx * (long long)p.m * 6 : nullptr;
        double *drows = p.derivs ? p.derivs + idx * (long long)p.m * 6 : nullptr;
        double yo[6], fo[6];
        while ((t - tf) < 0.0 && fin < 0) {
            h = hb_clamp_step(h, q.max_step, q.min_step);
            if (AR::add(t, h) > tf) h = fabs(AR::sub(tf, t));
            double err;
            if constexpr (D8) {
                dop853_stages<AR>(y, k, h, yh, rhs);
                double n5 = 0.0, n3 = 0.0;
                dop853_err_sums<AR>(y, yh, k, h, q.rtol, q.atol, n5, n3);
                err = dop853_err_norm<AR>(n5, n3, h, 6.0);
            } else {
                g_run_stages<AR, Tab45, SympRhs<AR>, 1>(rhs, y, k, h);
#pragma unroll
                for (int d = 0; d < 6; ++d) yh[d] = y[d];
                g_high_acc<AR, Tab45, 0>(yh, k, h);
                rhs(yh, k6);
                double ev[6];
#pragma unroll
                for (int d = 0; d < 6; ++d) ev[d] = 0.0;
                rk45_err_acc<AR, 0>(ev, k, k6, h);
#pragma unroll
                for (int d = 0; d < 6; ++d)
                    ev[d] = AR::div(ev[d], AR::madd(q.rtol, fmax(fabs(y[d]), fabs(yh[d])), q.atol));
                err = AR::div(hbc::norm2_ext6(ev), AR::sqrt(6.0));
            }
            ++attempts;
            if (err <= 1.0) {
                const double t_new = AR::add(t, h);
                ++nacc;
                const bool last = !((t_new - tf) < 0.0);
                const double hseg = EVENT ? h : AR::sub(t_new, t);
                bool crossed = false;
                double g_new = 0.0;
                if (EVENT) {
                    g_new = AR::sub(pick6c(yh, p.ev_idx), p.ev_off);
                    crossed = hb_event_crossed(g_prev, g_new, p.ev_dir);
                }
                if (crossed || (!EVENT && cursor < p.m && (last || p.t_vals[cursor] < t_new))) {
                    // the method's dense interpolant of this step
                    double F[D8 ? 7 : 1][6], Q[D8 ? 1 : 6][4];
                    if constexpr (D8) {
                        if (hseg != 0.0) dense_cache<AR>(y, yh, hseg, k, F, rhs);
                    } else {
#pragma unroll
                        for (int d = 0; d < 6; ++d)
#pragma unroll
                            for (int c = 0; c < 4; ++c) Q[d][c] = 0.0;
                        rk45_q_acc<AR, 0, 0>(Q, k, k6);
                    }
                    auto eval = [&](double x, double (&o)[6]) {
                        if constexpr (D8) dense_eval<AR>(y, F, x, o);
                        else rk45_eval<AR>(y, Q, x, hseg, o);
                    };
                    if (EVENT) {
                        // _dop853_refine_in_step_ham (rk.py:2106-2170) / _rk45_refine_in_step (:1072-1092)
                        double a = 0.0, b = 1.0, g_left = g_prev, xh = 1.0;
                        bool found = false;
                        for (int it = 0; it < 128; ++it) {
                            const double mid = AR::mul(0.5, AR::add(a, b));
                            eval(mid, yo);
                            const double g_mid = AR::sub(pick6c(yo, p.ev_idx), p.ev_off);
                            if (fabs(g_mid) <= p.gtol) { xh = mid; found = true; break; }
                            if (hb_crossed_direction(g_left, g_mid, p.ev_dir)) b = mid;
                            else { a = mid; g_left = g_mid; }
                            if (AR::mul(AR::sub(b, a), fabs(h)) <= p.xtol) break;
                        }
                        if (!found) xh = b;
                        eval(xh, yo);
                        t_end = AR::madd(xh, h, t);
                        fin = HB_TRAJ_HIT;
                    } else {
                        while (cursor < p.m) {
                            const double tq = p.t_vals[cursor];
                            if (!(last || tq < t_new)) break;
                            if (hseg == 0.0) {
#pragma unroll
                                for (int d = 0; d < 6; ++d) yo[d] = y[d];
                            } else {
                                eval(AR::div(AR::sub(tq, t), hseg), yo);
                            }
                            double *o = rows + (long long)cursor * 6;
#pragma unroll
                            for (int d = 0; d < 6; ++d) o[d] = yo[d];
                            if (drows) {
                                rhs(yo, fo);
                                double *od = drows + (long long)cursor * 6;
#pragma unroll
                                for (int d = 0; d < 6; ++d) od[d] = fo[d];
                            }
                            ++cursor;
                        }
                    }
                }
                if (EVENT && !crossed) g_prev = g_new;
                if (fin < 0) {
                    t = t_new;
#pragma unroll
                    for (int d = 0; d < 6; ++d) { y[d] = yh[d]; k[0][d] = D8 ? k[NK - 1][d] : k6[d]; }
                    if constexpr (D8) h = AR::mul(h, hb_pi_accept_factor<AR>(err, err_prev, 8.0));
                    else h = AR::mul(h, hb_pi_accept_factor<AR>(err, err_prev, 5.0));
                    err_prev = err;
                }
            } else {
                ++nrej;
                h = AR::mul(h, hb_pi_reject_factor<AR>(err, D8 ? 8.0 : 5.0));
                h = hb_clamp_step(h, q.max_step, q.min_step);
            }
            if (fin < 0) {
                if (!(h == h) || !(err == err)) fin = HB_TRAJ_NONFINITE;
                else if (attempts >= q.max_attempts) fin = HB_TRAJ_MAXSTEPS;
            }
        }
        if (fin < 0) fin = HB_TRAJ_OK;
        if (EVENT) {
            if (fin != HB_TRAJ_HIT) {
                t_end = t;
#pragma unroll
                for (int d = 0; d < 6; ++d) yo[d] = y[d];
            }
            p.t_hit[idx] = t_end;
#pragma unroll
            for (int d = 0; d < 6; ++d) p.y_hit[idx * 6 + d] = yo[d];
        } else {
            for (; cursor < p.m; ++cursor) {       // zero-length span or early termination: hold the last state
                double *o = rows + (long long)cursor * 6;
#pragma unroll
                for (int d = 0; d < 6; ++d) o[d] = y[d];
                if (drows) {
                    double *od = drows + (long long)cursor * 6;
#pragma unroll
                    for (int d = 0; d < 6; ++d) od[d] = k[0][d];
                }
            }
        }
        q.nacc[idx] = nacc; q.nrej[idx] = nrej; q.status[idx] = fin;
    }
}

template <class KERN>
int launch_ham_adaptive(KERN kern, const HamAdParams &q, size_t smem, cudaStream_t st)
{
    HB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    HB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CM_BLOCK, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (q.s.n + CM_BLOCK - 1) / CM_BLOCK;
    const long long cap = (long long)sms * per_sm;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, CM_BLOCK, smem, st>>>(q);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

template <bool EVENT>
int dispatch_ham_adaptive(const HamAdParams &q, int method, int arith, size_t smem, cudaStream_t st)
{
    const bool par = arith == HB_ARITH_PARITY;
    if (method == HB_DOP853)
        return par ? launch_ham_adaptive(k_ham_adaptive<ArParity, HB_DOP853, EVENT>, q, smem, st)
                   : launch_ham_adaptive(k_ham_adaptive<ArFast, HB_DOP853, EVENT>, q, smem, st);
    if (method == HB_RK45)
        return par ? launch_ham_adaptive(k_ham_adaptive<ArParity, HB_RK45, EVENT>, q, smem, st)
                   : launch_ham_adaptive(k_ham_adaptive<ArFast, HB_RK45, EVENT>, q, smem, st);
    return HB_ERR_UNSUPPORTED;
}

template <class AR, bool EVENT>
int launch_symp(const SympParams &p, size_t smem, cudaStream_t st)
{
    return launch_grid_kernel(k_symp_grid<AR, EVENT>, p, smem, st);
}

int ham_table(const hb_polyham *ham, int64_t n, const double *y0, void *workspace, SympParams &p, size_t &smem);

int symp_common(const hb_polyham *ham, const hb_symp_opts *o, int64_t n, const double *y0, const double *tao_tab,
                void *workspace, SympParams &p, size_t &smem)
{
    if (!ham || !o || !workspace || n < 0) return HB_ERR_BADARG;
    if (o->order < 2 || (o->order % 2) != 0 || o->order > 8) return HB_ERR_UNSUPPORTED;
    if (o->m < 2 || o->n_sub <= 0 || o->n_sub > HB_MAX_TAO_SUBSTEPS) return HB_ERR_BADARG;   // prepare not called
    if (o->arith != HB_ARITH_PARITY && o->arith != HB_ARITH_FAST) return HB_ERR_BADARG;
    if (n > 0 && !tao_tab) return HB_ERR_BADARG;
    p.m = o->m; p.n_sub = o->n_sub; p.tab = tao_tab;
    return ham_table(ham, n, y0, workspace, p, smem);
}

int ham_table(const hb_polyham *ham, int64_t n, const double *y0, void *workspace, SympParams &p, size_t &smem)
{
    if (!ham || !workspace || n < 0) return HB_ERR_BADARG;
    if (ham->n_dof != 3 || ham->max_deg < 0 || ham->max_deg > 30) return HB_ERR_UNSUPPORTED;
    if (n > 0 && (!y0 || !ham->terms)) return HB_ERR_BADARG;
    p.n = n; p.y0 = y0;
    p.ws = (HbWorkspace *)workspace;
    p.terms = (const TermMeta *)ham->terms;
    for (int i = 0; i < 7; ++i) p.ptr[i] = (int)ham->ptr[i];
    p.n_terms = (int)ham->ptr[6];
    p.D = ham->max_deg;
    smem = (((size_t)p.n_terms * sizeof(TermMeta) + 15) & ~(size_t)15) + (size_t)6 * (p.D + 1) * CM_BLOCK * sizeof(double);
    if (smem > 227 * 1024) return HB_ERR_UNSUPPORTED;
    return HB_OK;
}

}  // namespace

extern "C" int hb_cm_prepare(hb_cm_opts *o, double c_omega)
{
    if (!o) return HB_ERR_BADARG;
    o->n_sub = 0;
    if (o->method != HB_SYMPLECTIC) return (o->method == HB_RK4 || o->method == HB_RK6 || o->method == HB_RK8) ? HB_OK : HB_ERR_UNSUPPORTED;
    if (o->order < 2 || (o->order % 2) != 0 || o->order > 8) return HB_ERR_UNSUPPORTED;
    const double step = o->dt - 0.0;                                   // np.diff([0, dt])
    const double omega = std::pow(c_omega * step, -(double)o->order);  // _get_tao_omega (symplectic.py:38-60)
    tao_schedule(step, o->order, o);
    if (o->n_sub > HB_MAX_TAO_SUBSTEPS) return HB_ERR_UNSUPPORTED;
    for (int j = 0; j < o->n_sub; ++j) {
        o->sub_cos[j] = std::cos(2 * omega * o->sub_ts[j]);
        o->sub_sin[j] = std::sin(2 * omega * o->sub_ts[j]);
    }
    return HB_OK;
}

extern "C" int hb_cm_poincare_map(const hb_polyham *ham, const hb_cm_opts *opts, int64_t n, const double *seeds,
                                  int32_t *flags, double *out, double *t_out, void *workspace, void *stream)
{
    if (!ham || !opts || !workspace || n < 0) return HB_ERR_BADARG;
    if (ham->n_dof != 3 || ham->max_deg < 0 || ham->max_deg > 30) return HB_ERR_UNSUPPORTED;
    if (opts->section < 0 || opts->section > 3 || opts->max_steps < 0) return HB_ERR_BADARG;
    if (opts->arith != HB_ARITH_PARITY && opts->arith != HB_ARITH_FAST) return HB_ERR_BADARG;
    if (opts->method == HB_SYMPLECTIC && opts->n_sub <= 0) return HB_ERR_BADARG;   // hb_cm_prepare not called
    if (n > 0 && (!seeds || !flags || !out || !t_out || !ham->terms)) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    CmParams p{};
    p.o = *opts; p.n = n; p.seeds = seeds; p.flags = flags; p.out = out; p.t_out = t_out;
    p.ws = (HbWorkspace *)workspace;
    p.terms = (const TermMeta *)ham->terms;
    for (int i = 0; i < 7; ++i) p.ptr[i] = (int)ham->ptr[i];
    p.n_terms = (int)ham->ptr[6];
    p.D = ham->max_deg;
    const size_t smem = (((size_t)p.n_terms * sizeof(TermMeta) + 15) & ~(size_t)15) +
                        (size_t)6 * (p.D + 1) * CM_BLOCK * sizeof(double);
    if (smem > 227 * 1024) return HB_ERR_UNSUPPORTED;
    return (opts->arith == HB_ARITH_PARITY) ? dispatch<ArParity>(p, smem, st) : dispatch<ArFast>(p, smem, st);
}

extern "C" int hb_tao_grid_prepare(const double *t_vals_signed, int32_t m, int32_t order, double c_omega,
                                   int32_t *n_sub_out, double *tab, int64_t tab_capacity)
{
    if (!t_vals_signed || !n_sub_out || m < 2) return HB_ERR_BADARG;
    if (order < 2 || (order % 2) != 0 || order > 8) return HB_ERR_UNSUPPORTED;
    int n_sub = 1;
    for (int k = order; k > 2; k -= 2) n_sub *= 3;
    *n_sub_out = n_sub;
    if (!tab) return HB_OK;                                            // size query
    if (tab_capacity < (int64_t)(m - 1) * 3 * n_sub) return HB_ERR_BADARG;
    hb_cm_opts o;
    for (int i = 0; i < m - 1; ++i) {
        const double dt = t_vals_signed[i + 1] - t_vals_signed[i];     // np.diff(t_values) (symplectic.py:642)
        const double omega = std::pow(c_omega * dt, -(double)order);   // _get_tao_omega (symplectic.py:38-60)
        o.n_sub = 0;
        tao_schedule(dt, order, &o);
        double *row = tab + (size_t)i * 3 * n_sub;
        for (int j = 0; j < n_sub; ++j) {
            row[j] = o.sub_ts[j];
            row[n_sub + j] = std::cos(2 * omega * o.sub_ts[j]);
            row[2 * n_sub + j] = std::sin(2 * omega * o.sub_ts[j]);
        }
    }
    return HB_OK;
}

extern "C" int hb_ham_symplectic_dense(const hb_polyham *ham, const hb_symp_opts *opts, int64_t n, const double *y0,
                                       const double *tao_tab, double *traj, void *workspace, void *stream)
{
    SympParams p{};
    size_t smem = 0;
    const int rc = symp_common(ham, opts, n, y0, tao_tab, workspace, p, smem);
    if (rc != HB_OK) return rc;
    if (n > 0 && !traj) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    p.traj = traj;
    return (opts->arith == HB_ARITH_PARITY) ? launch_symp<ArParity, false>(p, smem, st) : launch_symp<ArFast, false>(p, smem, st);
}

extern "C" int hb_ham_symplectic_event(const hb_polyham *ham, const hb_symp_opts *opts, const hb_event *ev, int64_t n,
                                       const double *y0, const double *t_vals_signed, const double *tao_tab,
                                       double *traj, int32_t *hit, double *t_hit, double *y_hit, int32_t *n_rows,
                                       void *workspace, void *stream)
{
    SympParams p{};
    size_t smem = 0;
    const int rc = symp_common(ham, opts, n, y0, tao_tab, workspace, p, smem);
    if (rc != HB_OK) return rc;
    if (!ev || ev->idx < 0 || ev->idx > 5) return HB_ERR_BADARG;
    if (n > 0 && (!t_vals_signed || !hit || !t_hit || !y_hit || !n_rows)) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    p.traj = traj; p.t_vals = t_vals_signed;
    p.ev_idx = ev->idx; p.ev_dir = ev->direction; p.ev_off = ev->offset; p.xtol = ev->xtol; p.gtol = ev->gtol;
    p.hit = hit; p.t_hit = t_hit; p.y_hit = y_hit; p.n_rows = n_rows;
    return (opts->arith == HB_ARITH_PARITY) ? launch_symp<ArParity, true>(p, smem, st) : launch_symp<ArFast, true>(p, smem, st);
}

extern "C" int hb_ham_rk_dense(const hb_polyham *ham, int32_t method, int32_t arith, int64_t n, const double *y0,
                               const double *t_vals, int32_t m, double *traj, double *derivs, void *workspace,
                               void *stream)
{
    if (method != HB_RK4 && method != HB_RK6 && method != HB_RK8) return HB_ERR_UNSUPPORTED;
    if (arith != HB_ARITH_PARITY && arith != HB_ARITH_FAST) return HB_ERR_BADARG;
    if (m < 2) return HB_ERR_BADARG;
    SympParams p{};
    size_t smem = 0;
    const int rc = ham_table(ham, n, y0, workspace, p, smem);
    if (rc != HB_OK) return rc;
    if (n > 0 && (!traj || !t_vals)) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    p.m = m; p.t_vals = t_vals; p.traj = traj; p.derivs = derivs;
    return (arith == HB_ARITH_PARITY) ? launch_ham_rk<ArParity, false>(p, method, smem, st)
                                      : launch_ham_rk<ArFast, false>(p, method, smem, st);
}

extern "C" int hb_ham_rk_event(const hb_polyham *ham, int32_t method, int32_t arith, const hb_event *ev, int64_t n,
                               const double *y0, const double *t_vals, int32_t m, double *traj, int32_t *hit,
                               double *t_hit, double *y_hit, int32_t *n_rows, void *workspace, void *stream)
{
    if (method != HB_RK4 && method != HB_RK6 && method != HB_RK8) return HB_ERR_UNSUPPORTED;
    if (arith != HB_ARITH_PARITY && arith != HB_ARITH_FAST) return HB_ERR_BADARG;
    if (m < 2 || !ev || ev->idx < 0 || ev->idx > 5) return HB_ERR_BADARG;
    SympParams p{};
    size_t smem = 0;
    const int rc = ham_table(ham, n, y0, workspace, p, smem);
    if (rc != HB_OK) return rc;
    if (n > 0 && (!t_vals || !hit || !t_hit || !y_hit || !n_rows)) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    p.m = m; p.t_vals = t_vals; p.traj = traj;
    p.ev_idx = ev->idx; p.ev_dir = ev->direction; p.ev_off = ev->offset; p.xtol = ev->xtol; p.gtol = ev->gtol;
    p.hit = hit; p.t_hit = t_hit; p.y_hit = y_hit; p.n_rows = n_rows;
    return (arith == HB_ARITH_PARITY) ? launch_ham_rk<ArParity, true>(p, method, smem, st)
                                      : launch_ham_rk<ArFast, true>(p, method, smem, st);
}

static int ham_adaptive_fill(const hb_polyham *ham, const hb_integ *integ, int64_t n, const double *y0, void *workspace,
                             int32_t *n_acc, int32_t *n_rej, int32_t *status, HamAdParams &q, size_t &smem)
{
    if (!integ) return HB_ERR_BADARG;
    if (integ->method != HB_DOP853 && integ->method != HB_RK45) return HB_ERR_UNSUPPORTED;
    if (integ->arith != HB_ARITH_PARITY && integ->arith != HB_ARITH_FAST) return HB_ERR_BADARG;
    const int rc = ham_table(ham, n, y0, workspace, q.s, smem);
    if (rc != HB_OK) return rc;
    if (n > 0 && (!n_acc || !n_rej || !status)) return HB_ERR_BADARG;
    q.rtol = integ->rtol; q.atol = integ->atol; q.max_step = integ->max_step; q.min_step = integ->min_step;
    q.max_attempts = integ->max_attempts > 0 ? integ->max_attempts : 2147483647LL;
    q.nacc = n_acc; q.nrej = n_rej; q.status = status;
    return HB_OK;
}

extern "C" int hb_ham_adaptive_dense(const hb_polyham *ham, const hb_integ *integ, int64_t n, const double *y0,
                                     const double *t_eval, int32_t m, double *states, double *derivs, int32_t *n_acc,
                                     int32_t *n_rej, int32_t *status, void *workspace, void *stream)
{
    HamAdParams q{};
    size_t smem = 0;
    const int rc = ham_adaptive_fill(ham, integ, n, y0, workspace, n_acc, n_rej, status, q, smem);
    if (rc != HB_OK) return rc;
    if (m < 2 || (n > 0 && (!t_eval || !states))) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    q.s.m = m; q.s.t_vals = t_eval; q.s.traj = states; q.s.derivs = derivs;
    return dispatch_ham_adaptive<false>(q, integ->method, integ->arith, smem, st);
}

extern "C" int hb_ham_adaptive_event(const hb_polyham *ham, const hb_integ *integ, const hb_event *ev, int64_t n,
                                     const double *y0, double t0, double tmax, double *t_hit, double *y_hit,
                                     int32_t *n_acc, int32_t *n_rej, int32_t *status, void *workspace, void *stream)
{
    HamAdParams q{};
    size_t smem = 0;
    const int rc = ham_adaptive_fill(ham, integ, n, y0, workspace, n_acc, n_rej, status, q, smem);
    if (rc != HB_OK) return rc;
    if (!ev || ev->idx < 0 || ev->idx > 5 || (n > 0 && (!t_hit || !y_hit))) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    q.t0 = t0; q.tmax = tmax; q.s.t_hit = t_hit; q.s.y_hit = y_hit;
    q.s.ev_idx = ev->idx; q.s.ev_dir = ev->direction; q.s.ev_off = ev->offset; q.s.xtol = ev->xtol; q.s.gtol = ev->gtol;
    return dispatch_ham_adaptive<true>(q, integ->method, integ->arith, smem, st);
}
