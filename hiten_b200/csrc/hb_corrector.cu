// hb_corrector.cu -- batched single-shooting differential correction of periodic orbits.
//
// Replaces, for MANY independent orbits advanced in lock-step (paths relative to hiten/):
//   _NewtonBackend.run                        algorithms/corrector/backends/newton.py:20-150
//   _CorrectorBackend._compute_jacobian / _solve_delta_dense   algorithms/corrector/backends/base.py
//   _ArmijoLineSearch.__call__                algorithms/corrector/stepping/armijo.py:60-170  (plain.py for the capped step)
//   _SingleShootingOrbitOperators             algorithms/corrector/operators.py:319-452
//   _SingleHitBackend._cross / _cross_event_driven   algorithms/poincare/singlehit/backend.py:164-282
//   _halo_quadratic_term                      algorithms/types/services/orbits.py:917-946
//
// The reference corrects one orbit at a time: every Newton iteration is an event propagation to the symmetry plane
// (bit-exact here: hb_cr3bp_event, after hb_cr3bp_propagate to the window start when that is not ~0), a 42-state STM propagation to
// the event time (hb_cr3bp_stm, per-orbit tf) or four more event propagations (central differences), a 2x2 solve,
// and an Armijo back-tracking search whose every trial is another event propagation.  Here all orbits of a batch
// share those launches: per-orbit Newton / line-search state lives in HBM (SoA), small bookkeeping kernels decide
// per orbit (converged / accept / shrink / fail), a compaction kernel builds the list of orbits that take part in
// the next launch, and the host only reads that list's length.  No per-orbit host work, no CPU arithmetic.
#include <math_constants.h>

#include "hb_common.cuh"

namespace {

enum Phase : int {
    PH_ACTIVE = 0,      // has a current residual, waits for the next Newton step
    PH_SEARCH = 1,      // inside the line search (alpha, delta valid)
    PH_DONE = 10,       // converged
    PH_MAXATT = 11,     // max_attempts exhausted
    PH_STEPFAIL = 12,   // line search found no productive step
    PH_NOEVENT = 13,    // no plane crossing for the current iterate
    PH_SINGULAR = 14,   // singular Jacobian
};

enum Source : int { SRC_CURRENT = 0, SRC_TRIAL = 1, SRC_FD = 2 };

struct Corr {           // device views into the scratch block (n entries per row)
    long long n;
    const double *x0;   // [6][n] initial guesses (SoA)
    double *p;          // [2][n] current controls
    double *r;          // [2][n] current residual
    double *rnorm, *tev;
    double *xev;        // [6][n]
    double *delta;      // [2][n]
    double *alpha, *bestnorm, *bestalpha, *besttev;
    double *bestp;      // [2][n]
    double *bestr;      // [2][n]
    double *bestxev;    // [6][n]
    double *tr_t;       // trial event time
    double *tr_x;       // [6][n] trial event state
    double *fd;         // [8][n]: r(+h e_0), r(-h e_0), r(+h e_1), r(-h e_1), 2 residual components each
    int *tr_ok, *fd_ok;
    int *phase, *iters;
    int *list, *miss;   // compacted orbit ids
    int *counters;      // [0] list length, [1] miss length
    unsigned long long *steps;   // [0] 6-state attempted steps, [1] 42-state attempted steps
    // staging for the propagation launches (stride = launch size)
    double *stage, *align, *yhit, *thit, *tfs, *phi;
    int *nacc, *nrej, *st, *nacc2, *nrej2, *st2;
};

__global__ void k_init(Corr c, hb_correct_opts o)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += (long long)gridDim.x * blockDim.x) {
        c.p[i] = c.x0[(long long)o.ctrl[0] * c.n + i];
        c.p[c.n + i] = c.x0[(long long)o.ctrl[1] * c.n + i];
        c.phase[i] = PH_ACTIVE;
        c.iters[i] = 0;
        c.rnorm[i] = CUDART_NAN;
        c.tev[i] = CUDART_NAN;
    }
}

__global__ void k_compact(Corr c, int phase, int *list, int *counter)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += (long long)gridDim.x * blockDim.x) {
        const bool sel = c.phase[i] == phase;
        const unsigned m = __ballot_sync(__activemask(), sel);
        if (sel) {
            const int lane = threadIdx.x & 31;
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(counter, __popc(m));
            base = __shfl_sync(m, base, leader);
            list[base + __popc(m & ((1u << lane) - 1))] = (int)i;
        }
    }
}

// stage[c][j] = full state of orbit list[j] with its controls taken from `src`
__global__ void k_gather(Corr c, hb_correct_opts o, const int *list, int cnt, int src, int fd_col, double fd_sign)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        const long long i = list[j];
        double x[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) x[d] = c.x0[(long long)d * c.n + i];
        double p0 = c.p[i], p1 = c.p[c.n + i];
        if (src == SRC_TRIAL) {                                   // x_trial = x0 + alpha * delta (armijo.py:113)
            const double a = c.alpha[i];
            p0 = __dadd_rn(p0, __dmul_rn(a, c.delta[i]));
            p1 = __dadd_rn(p1, __dmul_rn(a, c.delta[c.n + i]));
        } else if (src == SRC_FD) {                               // x_p[i] += h_i / x_m[i] -= h_i (base.py)
            const double pv = fd_col == 0 ? p0 : p1;
            const double h = __dmul_rn(o.fd_step, fmax(1.0, fabs(pv)));
            if (fd_col == 0) p0 = fd_sign > 0 ? __dadd_rn(p0, h) : __dsub_rn(p0, h);
            else p1 = fd_sign > 0 ? __dadd_rn(p1, h) : __dsub_rn(p1, h);
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            double v = x[d];
            if (d == o.ctrl[0]) v = p0;
            if (d == o.ctrl[1]) v = p1;
            c.stage[(long long)d * cnt + j] = v;
        }
    }
}

// results of one event launch -> per-orbit trial buffers; orbits without a hit go to the miss list (first window)
// or are marked as failed trials (fallback window)
__global__ void k_scatter_event(Corr c, const int *list, int cnt, double t_start, double span, int final_window)
{
    unsigned long long steps = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        const long long i = list[j];
        steps += (unsigned long long)(c.nacc[j] + c.nrej[j] + c.nacc2[j] + c.nrej2[j]);
        const double trel = c.thit[j];
        const bool hit = c.st[j] == HB_TRAJ_OK && c.st2[j] == HB_TRAJ_HIT && trel < span && trel >= 0.0;
        if (hit) {
            c.tr_ok[i] = 1;
            c.tr_t[i] = __dadd_rn(t_start, trel);
#pragma unroll
            for (int d = 0; d < 6; ++d) c.tr_x[(long long)d * c.n + i] = c.yhit[(long long)d * cnt + j];
        } else if (final_window) {
            c.tr_ok[i] = 0;
        } else {
            c.miss[atomicAdd(&c.counters[1], 1)] = (int)i;
        }
    }
    if (steps) atomicAdd(&c.steps[0], steps);
}

HB_DEV void load_trial_residual(const Corr &c, const hb_correct_opts &o, long long i, double &r0, double &r1)
{
    r0 = __dsub_rn(c.tr_x[(long long)o.res[0] * c.n + i], o.target[0]);
    r1 = __dsub_rn(c.tr_x[(long long)o.res[1] * c.n + i], o.target[1]);
}

HB_DEV void take_trial_as_current(const Corr &c, long long i, double r0, double r1, double nrm)
{
    c.r[i] = r0;
    c.r[c.n + i] = r1;
    c.rnorm[i] = nrm;
    c.tev[i] = c.tr_t[i];
#pragma unroll
    for (int d = 0; d < 6; ++d) c.xev[(long long)d * c.n + i] = c.tr_x[(long long)d * c.n + i];
}

// trial buffers -> current residual (first evaluation, and after a plain step)
__global__ void k_set_current(Corr c, hb_correct_opts o, const int *list, int cnt, int fail_phase)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        const long long i = list[j];
        if (!c.tr_ok[i]) { c.phase[i] = fail_phase; continue; }
        double r0, r1;
        load_trial_residual(c, o, i, r0, r1);
        take_trial_as_current(c, i, r0, r1, fmax(fabs(r0), fabs(r1)));
        c.phase[i] = PH_ACTIVE;
    }
}

// newton.py:96-118 (top of iteration k) and :137-150 (after the loop)
__global__ void k_converged(Corr c, hb_correct_opts o, int k, int last)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += (long long)gridDim.x * blockDim.x) {
        if (c.phase[i] != PH_ACTIVE) continue;
        c.iters[i] = k;
        if (c.rnorm[i] < o.tol) c.phase[i] = PH_DONE;
        else if (last) c.phase[i] = PH_MAXATT;
    }
}

__global__ void k_fd_store(Corr c, hb_correct_opts o, const int *list, int cnt, int slot)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        const long long i = list[j];
        if (slot == 0) c.fd_ok[i] = 1;
        if (!c.tr_ok[i]) { c.fd_ok[i] = 0; continue; }
        double r0, r1;
        load_trial_residual(c, o, i, r0, r1);
        c.fd[(long long)(2 * slot) * c.n + i] = r0;
        c.fd[(long long)(2 * slot + 1) * c.n + i] = r1;
    }
}

// Jacobian (operators.py:437-450 / base.py central differences), _solve_delta_dense, step cap; then either the
// start of the line search (alpha = 1) or the plain step p += delta
__global__ void k_newton_delta(Corr c, hb_correct_opts o, double mu, const int *list, int cnt)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        const long long i = list[j];
        double J[4];
        if (o.finite_difference) {
            if (!c.fd_ok[i]) { c.phase[i] = PH_NOEVENT; continue; }
#pragma unroll
            for (int col = 0; col < 2; ++col) {
                const double pv = c.p[(long long)col * c.n + i];
                const double h = __dmul_rn(o.fd_step, fmax(1.0, fabs(pv)));
                const double den = __dmul_rn(2.0, h);
#pragma unroll
                for (int a = 0; a < 2; ++a)
                    J[2 * a + col] = __ddiv_rn(__dsub_rn(c.fd[(long long)(4 * col + a) * c.n + i],
                                                         c.fd[(long long)(4 * col + 2 + a) * c.n + i]), den);
            }
        } else {
            if (c.st[j] != HB_TRAJ_OK) { c.phase[i] = PH_SINGULAR; continue; }
            const double *phi = c.phi + 42ll * j;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) J[2 * a + b] = phi[6 * o.res[a] + o.ctrl[b]];
            if (o.halo_quadratic) {
                const double X = c.xev[i], Y = c.xev[c.n + i], Z = c.xev[2 * c.n + i];
                double vy = c.xev[4 * c.n + i];
                const double mu2 = __dsub_rn(1.0, mu);
                const double xa = __dadd_rn(X, mu), xb = __dsub_rn(X, mu2);
                const double yz = __dadd_rn(__dmul_rn(Y, Y), __dmul_rn(Z, Z));
                const double s1 = __dadd_rn(__dmul_rn(xa, xa), yz), s2 = __dadd_rn(__dmul_rn(xb, xb), yz);
                const double rho1 = __ddiv_rn(1.0, hb_pow_libm(s1, 1.5)), rho2 = __ddiv_rn(1.0, hb_pow_libm(s2, 1.5));
                const double omega_x = __dadd_rn(__dsub_rn(-__dmul_rn(__dmul_rn(mu2, xa), rho1),
                                                           __dmul_rn(__dmul_rn(mu, xb), rho2)), X);
                const double DD[2] = {__dadd_rn(__dmul_rn(2.0, vy), omega_x),
                                      __dsub_rn(-__dmul_rn(__dmul_rn(mu2, Z), rho1), __dmul_rn(__dmul_rn(mu, Z), rho2))};
                if (fabs(vy) < 1e-9) vy = vy != 0.0 ? copysign(1e-9, vy) : 1e-9;
                const int cols[2] = {0, 4};
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b)
                        J[2 * a + b] = __dsub_rn(J[2 * a + b], __ddiv_rn(__dmul_rn(DD[a], phi[6 + cols[b]]), vy));
            }
        }
        // _solve_delta_dense: 2-norm condition number (closed form for 2x2), ridge 1e-12 when > 1e8, solve(J, -r)
        double a = J[0], b = J[1], cc = J[2], d = J[3];
        const double s = a * a + b * b + cc * cc + d * d, det = a * d - b * cc;
        const double smax2 = 0.5 * (s + sqrt(fmax(s * s - 4.0 * det * det, 0.0)));
        const double cond = smax2 / fabs(det);
        if (cond != cond || cond > 1e8) { a += 1e-12; d += 1e-12; }
        double b0 = -c.r[i], b1 = -c.r[c.n + i];
        if (fabs(cc) > fabs(a)) {
            double t;
            t = a; a = cc; cc = t;
            t = b; b = d; d = t;
            t = b0; b0 = b1; b1 = t;
        }
        // np.linalg.solve = LAPACK dgesv on OpenBLAS kernels; roundings identified against numpy (oracle/ho_correct.c):
        // multiplier c * (1/a), Schur update mul + sub, both substitutions fused, final quotients true divisions
        const double l = __dmul_rn(cc, __ddiv_rn(1.0, a));
        const double u22 = __dsub_rn(d, __dmul_rn(l, b));
        if (a == 0.0 || u22 == 0.0 || !(fabs(u22) <= CUDART_INF)) { c.phase[i] = PH_SINGULAR; continue; }
        double d1 = __ddiv_rn(__fma_rn(-l, b0, b1), u22);
        double d0 = __ddiv_rn(__fma_rn(-b, d1, b0), a);
        if (o.max_delta < CUDART_INF) {                             // armijo.py:98-107 / plain.py
            const double dn = fmax(fabs(d0), fabs(d1));
            if (dn > o.max_delta) {
                const double sc = __ddiv_rn(o.max_delta, dn);
                d0 = __dmul_rn(d0, sc);
                d1 = __dmul_rn(d1, sc);
            }
        }
        if (o.line_search) {
            c.delta[i] = d0;
            c.delta[c.n + i] = d1;
            c.alpha[i] = 1.0;
            c.bestnorm[i] = c.rnorm[i];
            c.bestalpha[i] = 0.0;
            c.phase[i] = PH_SEARCH;
        } else {
            c.p[i] = __dadd_rn(c.p[i], d0);
            c.p[c.n + i] = __dadd_rn(c.p[c.n + i], d1);
        }
    }
}

// one round of the back-tracking loop (armijo.py:110-170) for the orbits that just evaluated a trial
__global__ void k_armijo_update(Corr c, hb_correct_opts o, const int *list, int cnt)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) {
        const long long i = list[j];
        const double alpha = c.alpha[i];
        if (c.tr_ok[i]) {
            double r0, r1;
            load_trial_residual(c, o, i, r0, r1);
            const double nt = fmax(fabs(r0), fabs(r1));
            const double p0 = __dadd_rn(c.p[i], __dmul_rn(alpha, c.delta[i]));
            const double p1 = __dadd_rn(c.p[c.n + i], __dmul_rn(alpha, c.delta[c.n + i]));
            if (nt <= __dmul_rn(__dsub_rn(1.0, __dmul_rn(o.armijo_c, alpha)), c.rnorm[i])) {
                c.p[i] = p0;
                c.p[c.n + i] = p1;
                take_trial_as_current(c, i, r0, r1, nt);
                c.phase[i] = PH_ACTIVE;
                continue;
            }
            if (nt < c.bestnorm[i]) {
                c.bestnorm[i] = nt;
                c.bestalpha[i] = alpha;
                c.bestp[i] = p0;
                c.bestp[c.n + i] = p1;
                c.bestr[i] = r0;
                c.bestr[c.n + i] = r1;
                c.besttev[i] = c.tr_t[i];
#pragma unroll
                for (int d = 0; d < 6; ++d) c.bestxev[(long long)d * c.n + i] = c.tr_x[(long long)d * c.n + i];
            }
        }
        const double an = __dmul_rn(alpha, o.alpha_reduction);
        c.alpha[i] = an;
        if (an >= o.min_alpha) continue;                            // next round
        if (c.bestalpha[i] > 0.0) {                                 // fallback: best point seen
            c.p[i] = c.bestp[i];
            c.p[c.n + i] = c.bestp[c.n + i];
            c.r[i] = c.bestr[i];
            c.r[c.n + i] = c.bestr[c.n + i];
            c.rnorm[i] = c.bestnorm[i];
            c.tev[i] = c.besttev[i];
#pragma unroll
            for (int d = 0; d < 6; ++d) c.xev[(long long)d * c.n + i] = c.bestxev[(long long)d * c.n + i];
            c.phase[i] = PH_ACTIVE;
        } else {
            c.phase[i] = PH_STEPFAIL;
        }
    }
}

__global__ void k_stm_steps(Corr c, int cnt)
{
    unsigned long long steps = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x)
        steps += (unsigned long long)(c.nacc[j] + c.nrej[j]);
    if (steps) atomicAdd(&c.steps[1], steps);
}

__global__ void k_copy_tf(Corr c, const int *list, int cnt)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += gridDim.x * blockDim.x) c.tfs[j] = c.tev[list[j]];
}

__global__ void k_finish(Corr c, hb_correct_opts o, double *xc, double *half, int *iters, double *rnorm, int *status)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            double v = c.x0[(long long)d * c.n + i];
            if (d == o.ctrl[0]) v = c.p[i];
            if (d == o.ctrl[1]) v = c.p[c.n + i];
            xc[(long long)d * c.n + i] = v;
        }
        const int ph = c.phase[i];
        status[i] = ph == PH_DONE ? HB_CORR_CONVERGED : ph == PH_MAXATT ? HB_CORR_MAX_ATTEMPTS
                  : ph == PH_STEPFAIL ? HB_CORR_STEP_FAILED : ph == PH_NOEVENT ? HB_CORR_NO_EVENT : HB_CORR_SINGULAR;
        half[i] = ph == PH_DONE ? c.tev[i] : CUDART_NAN;           // _half_period: event time of the corrected state
        iters[i] = c.iters[i];
        rnorm[i] = c.rnorm[i];
    }
}

struct Layout {
    size_t off = 0;
    template <class T> T *take(char *base, size_t count)
    {
        off = (off + 255) & ~(size_t)255;
        T *p = base ? (T *)(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

size_t carve(Corr &c, char *base, long long n)
{
    Layout L;
    const size_t N = (size_t)n;
    c.p = L.take<double>(base, 2 * N);       c.r = L.take<double>(base, 2 * N);
    c.rnorm = L.take<double>(base, N);       c.tev = L.take<double>(base, N);
    c.xev = L.take<double>(base, 6 * N);     c.delta = L.take<double>(base, 2 * N);
    c.alpha = L.take<double>(base, N);       c.bestnorm = L.take<double>(base, N);
    c.bestalpha = L.take<double>(base, N);   c.besttev = L.take<double>(base, N);
    c.bestp = L.take<double>(base, 2 * N);   c.bestr = L.take<double>(base, 2 * N);
    c.bestxev = L.take<double>(base, 6 * N); c.tr_t = L.take<double>(base, N);
    c.tr_x = L.take<double>(base, 6 * N);    c.fd = L.take<double>(base, 8 * N);
    c.tr_ok = L.take<int>(base, N);          c.fd_ok = L.take<int>(base, N);
    c.phase = L.take<int>(base, N);          c.iters = L.take<int>(base, N);
    c.list = L.take<int>(base, N);           c.miss = L.take<int>(base, N);
    c.counters = L.take<int>(base, 64);      c.steps = L.take<unsigned long long>(base, 8);
    c.stage = L.take<double>(base, 6 * N);   c.align = L.take<double>(base, 6 * N);
    c.yhit = L.take<double>(base, 6 * N);    c.thit = L.take<double>(base, N);
    c.tfs = L.take<double>(base, N);         c.phi = L.take<double>(base, 42 * N);
    c.nacc = L.take<int>(base, N);           c.nrej = L.take<int>(base, N);
    c.st = L.take<int>(base, N);             c.nacc2 = L.take<int>(base, N);
    c.nrej2 = L.take<int>(base, N);          c.st2 = L.take<int>(base, N);
    return (L.off + 255) & ~(size_t)255;
}

inline unsigned blocks_for(long long n, int threads = 256)
{
    long long b = (n + threads - 1) / threads;
    return (unsigned)(b < 1 ? 1 : b > 4096 ? 4096 : b);
}

struct Driver {
    Corr c;
    hb_cr3bp sys;
    hb_integ integ, integ_ev;
    hb_correct_opts o;
    hb_event ev;
    void *ws;
    cudaStream_t st;
    int rc = HB_OK;

    bool ok(int r) { if (r != HB_OK && rc == HB_OK) rc = r; return rc == HB_OK; }
    bool cuda_ok(cudaError_t e) { return ok(e == cudaSuccess ? HB_OK : (int)e); }

    // number of orbits in `phase`, compacted into list (counter slot 0) -- one 4-byte read-back
    int compact(int phase, int *list, int slot)
    {
        if (!cuda_ok(cudaMemsetAsync(c.counters + slot, 0, sizeof(int), st))) return 0;
        k_compact<<<blocks_for(c.n), 256, 0, st>>>(c, phase, list, c.counters + slot);
        return read_counter(slot);
    }
    int read_counter(int slot)
    {
        int h = 0;
        if (!cuda_ok(cudaMemcpyAsync(&h, c.counters + slot, sizeof(int), cudaMemcpyDeviceToHost, st))) return 0;
        if (!cuda_ok(cudaStreamSynchronize(st))) return 0;
        return h;
    }
    // one window of _cross_event_driven for the staged states (backend.py:196-233)
    void window(const int *list, int cnt, double t0, double tmax, int final_window)
    {
        double t_start = t0;
        if (t_start <= 0.0) t_start = 1e-12;
        double span = tmax - t_start;
        if (span < 0.0) span = 0.0;
        if (fabs(0.0 - t_start) <= 1e-8 + 1e-5 * fabs(t_start)) {
            // np.isclose(t_eval[0], t_eval[-1]) -> _propagate_dynsys returns the initial state (base.py:420-424):
            // the 1e-12 "alignment" of the first window is the identity
            if (!cuda_ok(cudaMemcpyAsync(c.align, c.stage, sizeof(double) * 6 * (size_t)cnt, cudaMemcpyDeviceToDevice, st)))
                return;
            if (!cuda_ok(cudaMemsetAsync(c.nacc, 0, sizeof(int) * (size_t)cnt, st))) return;
            if (!cuda_ok(cudaMemsetAsync(c.nrej, 0, sizeof(int) * (size_t)cnt, st))) return;
            if (!cuda_ok(cudaMemsetAsync(c.st, 0, sizeof(int) * (size_t)cnt, st))) return;
        } else if (!ok(hb_cr3bp_propagate(&sys, &integ, cnt, c.stage, 0.0, t_start, nullptr, 0, c.align, c.nacc, c.nrej,
                                          c.st, ws, st))) return;
        if (!ok(hb_cr3bp_event(&sys, &integ_ev, &ev, cnt, c.align, 0.0, span, nullptr, c.thit, c.yhit, c.nacc2, c.nrej2,
                               c.st2, ws, st))) return;
        k_scatter_event<<<blocks_for(cnt), 256, 0, st>>>(c, list, cnt, t_start, span, final_window);
    }
    // _cross with t_guess = None for the listed orbits: window [0, pi], then [pi/2 - 0.15, + pi] for the misses
    void event_eval(const int *list, int cnt, int src, int fd_col = 0, double fd_sign = 0.0)
    {
        if (cnt <= 0 || rc != HB_OK) return;
        const double pi = 3.141592653589793;
        const double t_start = pi / 2.0 - 0.15, half_span = pi * 0.5;
        double t0 = t_start - half_span;
        if (t0 < 0.0) t0 = 0.0;
        if (!cuda_ok(cudaMemsetAsync(c.counters + 1, 0, sizeof(int), st))) return;
        k_gather<<<blocks_for(cnt), 256, 0, st>>>(c, o, list, cnt, src, fd_col, fd_sign);
        window(list, cnt, t0, t0 + 2.0 * half_span, 0);
        const int n_miss = read_counter(1);
        if (n_miss > 0 && rc == HB_OK) {
            k_gather<<<blocks_for(n_miss), 256, 0, st>>>(c, o, c.miss, n_miss, src, fd_col, fd_sign);
            window(c.miss, n_miss, t_start, t_start + pi, 1);
        }
    }

    void run()
    {
        k_init<<<blocks_for(c.n), 256, 0, st>>>(c, o);
        int cnt = compact(PH_ACTIVE, c.list, 0);
        event_eval(c.list, cnt, SRC_CURRENT);
        k_set_current<<<blocks_for(cnt), 256, 0, st>>>(c, o, c.list, cnt, PH_NOEVENT);
        for (int k = 0; rc == HB_OK; ++k) {
            const int last = k >= o.max_attempts;
            k_converged<<<blocks_for(c.n), 256, 0, st>>>(c, o, k, last);
            if (last) break;
            cnt = compact(PH_ACTIVE, c.list, 0);
            if (cnt == 0) break;
            if (o.finite_difference) {
                for (int col = 0; col < 2; ++col)
                    for (int sg = 0; sg < 2; ++sg) {
                        event_eval(c.list, cnt, SRC_FD, col, sg == 0 ? 1.0 : -1.0);
                        k_fd_store<<<blocks_for(cnt), 256, 0, st>>>(c, o, c.list, cnt, 2 * col + sg);
                    }
            } else {
                k_gather<<<blocks_for(cnt), 256, 0, st>>>(c, o, c.list, cnt, SRC_CURRENT, 0, 0.0);
                k_copy_tf<<<blocks_for(cnt), 256, 0, st>>>(c, c.list, cnt);
                if (!ok(hb_cr3bp_stm(&sys, &integ, cnt, c.stage, 0.0, 0.0, c.tfs, c.phi, c.nacc, c.nrej, c.st, ws, st)))
                    break;
                k_stm_steps<<<blocks_for(cnt), 256, 0, st>>>(c, cnt);
            }
            k_newton_delta<<<blocks_for(cnt), 256, 0, st>>>(c, o, sys.mu, c.list, cnt);
            if (o.line_search) {
                while (rc == HB_OK) {
                    const int ns = compact(PH_SEARCH, c.list, 0);
                    if (ns == 0) break;
                    event_eval(c.list, ns, SRC_TRIAL);
                    k_armijo_update<<<blocks_for(ns), 256, 0, st>>>(c, o, c.list, ns);
                }
            } else {
                cnt = compact(PH_ACTIVE, c.list, 0);            // orbits that just stepped (singular ones dropped out)
                event_eval(c.list, cnt, SRC_CURRENT);
                k_set_current<<<blocks_for(cnt), 256, 0, st>>>(c, o, c.list, cnt, PH_STEPFAIL);
            }
        }
    }
};

}  // namespace

extern "C" int64_t hb_correct_scratch_bytes(int64_t n)
{
    if (n < 0) return HB_ERR_BADARG;
    Corr c{};
    return (int64_t)carve(c, nullptr, n > 0 ? n : 1);
}

extern "C" int hb_correct_orbits(const hb_cr3bp *sys, const hb_integ *integ, const hb_correct_opts *opts, int64_t n,
                                 const double *x0_soa, double *xc_soa, double *half_period, int32_t *iterations,
                                 double *residual_norm, int32_t *status, int64_t *rk_steps6, int64_t *rk_steps42,
                                 void *scratch, int64_t scratch_bytes, void *workspace, void *stream)
{
    if (!sys || !integ || !opts || n < 0 || !workspace) return HB_ERR_BADARG;
    if (sys->fwd != 1) return HB_ERR_UNSUPPORTED;                 // the shipped correction configs integrate forward
    if (integ->method != HB_DOP853) return HB_ERR_UNSUPPORTED;    // IntegrationConfig(method="adaptive"), order 8
    for (int k = 0; k < 2; ++k)
        if (opts->ctrl[k] < 0 || opts->ctrl[k] > 5 || opts->res[k] < 0 || opts->res[k] > 5) return HB_ERR_BADARG;
    if (opts->ctrl[0] == opts->ctrl[1] || opts->event_idx < 0 || opts->event_idx > 5) return HB_ERR_BADARG;
    if (opts->max_attempts < 0 || !(opts->tol > 0.0) || !(opts->max_delta > 0.0)) return HB_ERR_BADARG;
    if (opts->line_search && (!(opts->alpha_reduction > 0.0 && opts->alpha_reduction < 1.0) || !(opts->min_alpha > 0.0)))
        return HB_ERR_BADARG;
    if (opts->finite_difference && !(opts->fd_step > 0.0)) return HB_ERR_BADARG;
    if (rk_steps6) *rk_steps6 = 0;
    if (rk_steps42) *rk_steps42 = 0;
    if (n == 0) return HB_OK;
    if (n > 0x7fffffff) return HB_ERR_UNSUPPORTED;
    if (!x0_soa || !xc_soa || !half_period || !iterations || !residual_norm || !status || !scratch) return HB_ERR_BADARG;
    Driver d{};
    if (scratch_bytes < (int64_t)carve(d.c, (char *)scratch, n)) return HB_ERR_BADARG;
    d.c.n = n;
    d.c.x0 = x0_soa;
    d.sys = *sys;
    d.integ = *integ;
    d.integ_ev = *integ;
    d.integ_ev.max_step = 1e300;                                   // RungeKutta(order=853, rtol, atol): max_step = inf
    d.o = *opts;
    d.ev = hb_event{opts->event_idx, 0, opts->event_offset, 1e-12, 1e-12};
    d.ws = workspace;
    d.st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(d.c.steps, 0, 8 * sizeof(unsigned long long), d.st));
    d.run();
    if (d.rc != HB_OK) return d.rc;
    k_finish<<<blocks_for(n), 256, 0, d.st>>>(d.c, d.o, xc_soa, half_period, iterations, residual_norm, status);
    unsigned long long hs[2] = {0, 0};
    HB_CUDA_TRY(cudaMemcpyAsync(hs, d.c.steps, sizeof hs, cudaMemcpyDeviceToHost, d.st));
    HB_CUDA_TRY(cudaStreamSynchronize(d.st));
    HB_CUDA_TRY(cudaGetLastError());
    if (rk_steps6) *rk_steps6 = (int64_t)hs[0];
    if (rk_steps42) *rk_steps42 = (int64_t)hs[1];
    return HB_OK;
}
