/* hiten_oracle.c -- CPU restatement of HITEN's vector fields and Runge-Kutta integrators.
 * TEST INFRASTRUCTURE (see hiten_oracle.h).  Every function cites the reference lines it
 * follows (paths relative to /root/reference/src/hiten/).
 *
 * Arithmetic rules kept from the reference's Numba code generation (fastmath=False):
 *   - no FMA contraction (compile with -ffp-contract=off), Python operator precedence and
 *     left-to-right association, x**2 -> x*x, r**3 -> r*(r*r) (Numba static-power lowering);
 *   - np.dot on short float64 vectors = OpenBLAS ddot scalar tail = sequential FMA
 *     accumulation (measured in the build container, n=6: 2000/2000 bit matches);
 *   - np.linalg.norm = OpenBLAS dnrm2 x87 kernel = extended-precision accumulate + sqrt
 *     (measured, n=6: 2000/2000 bit matches with long double);
 *   - float ** float = libm pow().
 */
#include "hiten_oracle.h"
#include "ho_coeffs.h"
#include "ho_poly.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define HO_MAXDIM 42
#define HO_MAXSTAGE 16

/* ===================================================================================== */
/* Vector fields                                                                         */
/* ===================================================================================== */

/* dynamics/rtbp.py:65-74 */
void ho_crtbp_accel(const double *s, double mu, double *out)
{
    const double x = s[0], y = s[1], z = s[2], vx = s[3], vy = s[4], vz = s[5];
    const double xm = x + mu;
    const double om = 1.0 - mu;
    const double xo = x - om;
    const double yy = y * y, zz = z * z;
    const double r1 = sqrt(xm * xm + yy + zz);
    const double r2 = sqrt(xo * xo + yy + zz);
    const double r1c = r1 * (r1 * r1);
    const double r2c = r2 * (r2 * r2);
    const double ax = 2.0 * vy + x - om * xm / r1c - mu * (x - 1.0 + mu) / r2c;
    const double ay = -2.0 * vx + y - om * y / r1c - mu * y / r2c;
    const double az = -om * z / r1c - mu * z / r2c;
    out[0] = vx; out[1] = vy; out[2] = vz; out[3] = ax; out[4] = ay; out[5] = az;
}

/* dynamics/rtbp.py:115-165 (float ** 1.5 / 2.5 go through libm pow) */
void ho_jacobian_crtbp(double x, double y, double z, double mu, double *F)
{
    const double mu2 = 1.0 - mu;
    const double xm = x + mu, xo = x - mu2;
    const double r2 = xm * xm + y * y + z * z;
    const double R2 = xo * xo + y * y + z * z;
    const double r3 = pow(r2, 1.5), r5 = pow(r2, 2.5);
    const double R3 = pow(R2, 1.5), R5 = pow(R2, 2.5);
    const double common = mu2 / r3 + mu / R3;
    const double omgxx = 1.0 + mu2 / r5 * 3.0 * (xm * xm) + mu / R5 * 3.0 * (xo * xo) - common;
    const double omgyy = 1.0 + mu2 / r5 * 3.0 * (y * y) + mu / R5 * 3.0 * (y * y) - common;
    const double omgzz = 0.0 + mu2 / r5 * 3.0 * (z * z) + mu / R5 * 3.0 * (z * z) - common;
    const double omgxy = 3.0 * y * (mu2 * xm / r5 + mu * xo / R5);
    const double omgxz = 3.0 * z * (mu2 * xm / r5 + mu * xo / R5);
    const double omgyz = 3.0 * y * z * (mu2 / r5 + mu / R5);
    memset(F, 0, 36 * sizeof(double));
    F[0 * 6 + 3] = 1.0; F[1 * 6 + 4] = 1.0; F[2 * 6 + 5] = 1.0;
    F[3 * 6 + 0] = omgxx; F[3 * 6 + 1] = omgxy; F[3 * 6 + 2] = omgxz;
    F[4 * 6 + 0] = omgxy; F[4 * 6 + 1] = omgyy; F[4 * 6 + 2] = omgyz;
    F[5 * 6 + 0] = omgxz; F[5 * 6 + 1] = omgyz; F[5 * 6 + 2] = omgzz;
    F[3 * 6 + 4] = 2.0; F[4 * 6 + 3] = -2.0;
}

/* dynamics/rtbp.py:210-255 */
void ho_var_equations(const double *PHI, double mu, double *out)
{
    const double *Phi = PHI;
    const double x = PHI[36], y = PHI[37], z = PHI[38], vx = PHI[39], vy = PHI[40], vz = PHI[41];
    double F[36];
    ho_jacobian_crtbp(x, y, z, mu, F);
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += F[i * 6 + k] * Phi[k * 6 + j];
            out[i * 6 + j] = s;
        }
    const double mu2 = 1.0 - mu;
    const double xm = x + mu, xo = x - mu2;
    const double r2 = xm * xm + y * y + z * z;
    const double R2 = xo * xo + y * y + z * z;
    const double r3 = pow(r2, 1.5), R3 = pow(R2, 1.5);
    out[36] = vx; out[37] = vy; out[38] = vz;
    out[39] = x - mu2 * (xm / r3) - mu * (xo / R3) + 2.0 * vy;
    out[40] = y - mu2 * (y / r3) - mu * (y / R3) - 2.0 * vx;
    out[41] = -mu2 * (z / r3) - mu * (z / R3);
}

/* base vector field + _DirectedSystem wrapper (dynamics/base.py:296-305) */
void ho_rhs(const ho_system *sys, double t, const double *y, double *dy)
{
    (void)t;
    switch (sys->kind) {
    case HO_SYS_CR3BP6: ho_crtbp_accel(y, sys->mu, dy); break;
    case HO_SYS_VAR42: ho_var_equations(y, sys->mu, dy); break;
    default: ho_polyham_rhs(sys->ham, y, dy); break;
    }
    if (sys->fwd == -1) {
        int lo = sys->flip_lo, hi = sys->flip_hi;
        if (lo < 0) { lo = 0; hi = sys->dim; }
        for (int d = lo; d < hi; ++d) dy[d] = -dy[d];
    }
}

static inline double ev_g(const ho_event *ev, const double *y) { return y[ev->idx] - ev->offset; }

/* ===================================================================================== */
/* Controller helpers  (algorithms/integrators/utils.py)                                  */
/* ===================================================================================== */
static int event_crossed(double gp, double gn, int dir)                      /* utils.py:14-39 */
{
    if (dir == 0) return (gp < 0.0 && gn > 0.0) || (gp > 0.0 && gn < 0.0) || (gn == 0.0);
    if (dir > 0) return (gp < 0.0 && gn > 0.0) || (gn == 0.0);
    return (gp > 0.0 && gn < 0.0) || (gn == 0.0);
}
static int crossed_direction(double gl, double gm, int dir)                  /* utils.py:43-69 */
{
    if (dir == 0) return (gl < 0.0 && gm > 0.0) || (gl > 0.0 && gm < 0.0);
    if (dir > 0) return (gl < 0.0 && gm > 0.0);
    return (gl > 0.0 && gm < 0.0);
}
static double select_initial_step(double d0, double d1, double mn, double mx) /* utils.py:127-157 */
{
    double h = (d0 < 1.0e-5 || d1 < 1.0e-5) ? 1.0e-6 : 0.01 * d0 / d1;
    if (h > mx) h = mx;
    if (h < mn) h = mn;
    return h;
}
static double clamp_step(double h, double mx, double mn)                     /* utils.py:161-182 */
{
    if (h > mx) h = mx;
    if (h < mn) h = mn;
    return h;
}
static double adjust_to_endpoint(double t, double h, double tend)            /* utils.py:186-207 */
{
    if (t + h > tend) return fabs(tend - t);
    return h;
}
static double pi_accept_factor(double err, double err_prev, double order)    /* utils.py:216-255 */
{
    const double beta = 1.0 / (order + 1.0);
    const double alpha = 0.4 * beta;
    double f;
    if (err_prev < 0.0) f = (err == 0.0) ? 10.0 : 0.9 * pow(err, -beta);
    else f = (err == 0.0) ? 10.0 : 0.9 * pow(err, -beta) * pow(err_prev, alpha);
    if (!(f == f)) f = 10.0;
    if (f < 0.2) f = 0.2;
    if (f > 10.0) f = 10.0;
    return f;
}
static double pi_reject_factor(double err, double order)                     /* utils.py:259-287 */
{
    const double e = 1.0 / order;
    double f = (err <= 0.0) ? 0.2 : 0.9 * pow(err, -e);
    if (!(f == f)) f = 0.2;
    if (f < 0.2) f = 0.2;
    if (f > 10.0) f = 10.0;
    return f;
}

/* np.linalg.norm on a float64 vector -> OpenBLAS dnrm2 (x87 extended accumulate). */
static double np_norm(const double *v, int n)
{
    long double s = 0.0L;
    for (int i = 0; i < n; ++i) s += (long double)v[i] * (long double)v[i];
    return (double)sqrtl(s);
}
/* np.dot(a, a) on a short float64 vector -> OpenBLAS ddot scalar tail, FMA-contracted. */
static double np_dot(const double *a, const double *b, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = fma(b[i], a[i], s);
    return s;
}

/* scale0 / d0 / d1 / h0  (rk.py:2445-2448) */
static double initial_step(const double *y, const double *f0, int n, const ho_tol *tol)
{
    double a[HO_MAXDIM], b[HO_MAXDIM];
    for (int d = 0; d < n; ++d) {
        const double sc = tol->atol + tol->rtol * fabs(y[d]);
        a[d] = y[d] / sc;
        b[d] = f0[d] / sc;
    }
    const double sq = sqrt((double)n);
    const double d0 = np_norm(a, n) / sq;
    const double d1 = np_norm(b, n) / sq;
    return select_initial_step(d0, d1, tol->min_step, tol->max_step);
}

/* ===================================================================================== */
/* Generic explicit RK stage sweep (rk.py:1670-1685, 873-888, 186-199)                     */
/* ===================================================================================== */
typedef struct {
    int s;                /* number of stages summed by B                                */
    const double *A;      /* row-major, lda columns                                       */
    int lda;
    const double *B;
    const double *C;
    int extra_stage;      /* 1: also k[s] = f(t+h, y_high)                               */
} rk_tab;

static void rk_stages(const ho_system *sys, const rk_tab *tb, double t, const double *y, double h,
                      double *k /* [s+1][n] */, double *y_high)
{
    const int n = sys->dim, s = tb->s;
    double ys[HO_MAXDIM];
    ho_rhs(sys, t, y, k);
    for (int i = 1; i < s; ++i) {
        memcpy(ys, y, n * sizeof(double));
        for (int j = 0; j < i; ++j) {
            const double a = tb->A[i * tb->lda + j];
            if (a != 0.0) {
                const double ha = h * a;
                const double *kj = k + j * n;
                for (int d = 0; d < n; ++d) ys[d] += ha * kj[d];
            }
        }
        ho_rhs(sys, t + tb->C[i] * h, ys, k + i * n);
    }
    memcpy(y_high, y, n * sizeof(double));
    for (int j = 0; j < s; ++j) {
        const double b = tb->B[j];
        if (b != 0.0) {
            const double hb = h * b;
            const double *kj = k + j * n;
            for (int d = 0; d < n; ++d) y_high[d] += hb * kj[d];
        }
    }
    if (tb->extra_stage) ho_rhs(sys, t + h, y_high, k + s * n);
}

static void get_tab(int method, rk_tab *tb)
{
    switch (method) {
    case HO_DOP853: tb->s = 12; tb->A = &HO_DOP853_A[0][0]; tb->lda = 16; tb->B = HO_DOP853_B; tb->C = HO_DOP853_C; tb->extra_stage = 1; break;
    case HO_RK45: tb->s = 6; tb->A = &HO_RK45_A[0][0]; tb->lda = 5; tb->B = HO_RK45_B; tb->C = HO_RK45_C; tb->extra_stage = 1; break;
    case HO_RK4: tb->s = 4; tb->A = &HO_RK4_A[0][0]; tb->lda = 4; tb->B = HO_RK4_B; tb->C = HO_RK4_C; tb->extra_stage = 0; break;
    case HO_RK6: tb->s = 7; tb->A = &HO_RK6_A[0][0]; tb->lda = 7; tb->B = HO_RK6_B; tb->C = HO_RK6_C; tb->extra_stage = 0; break;
    default: tb->s = 13; tb->A = &HO_RK8_A[0][0]; tb->lda = 13; tb->B = HO_RK8_B; tb->C = HO_RK8_C; tb->extra_stage = 0; break;
    }
}

/* One adaptive attempt: stages + error norm.  DOP853: rk.py:1686-1699, 2457-2467.
 * RK45: rk.py:889-896, 1332-1333. */
static double adaptive_attempt(const ho_system *sys, int method, const rk_tab *tb, const ho_tol *tol,
                               double t, const double *y, double h, double *k, double *y_high)
{
    const int n = sys->dim;
    rk_stages(sys, tb, t, y, h, k, y_high);
    double scale[HO_MAXDIM];
    for (int d = 0; d < n; ++d)
        scale[d] = tol->atol + tol->rtol * fmax(fabs(y[d]), fabs(y_high[d]));
    if (method == HO_DOP853) {
        double e5[HO_MAXDIM], e3[HO_MAXDIM];
        for (int d = 0; d < n; ++d) { e5[d] = 0.0; e3[d] = 0.0; }
        for (int j = 0; j < 13; ++j) {
            const double c5 = HO_DOP853_E5[j], c3 = HO_DOP853_E3[j];
            const double *kj = k + j * n;
            if (c5 != 0.0) for (int d = 0; d < n; ++d) e5[d] += c5 * kj[d];
            if (c3 != 0.0) for (int d = 0; d < n; ++d) e3[d] += c3 * kj[d];
        }
        for (int d = 0; d < n; ++d) { e5[d] *= h; e3[d] *= h; }
        for (int d = 0; d < n; ++d) { e5[d] = e5[d] / scale[d]; e3[d] = e3[d] / scale[d]; }
        const double n5 = np_dot(e5, e5, n), n3 = np_dot(e3, e3, n);
        if (n5 == 0.0 && n3 == 0.0) return 0.0;
        const double denom = n5 + 0.01 * n3;
        return fabs(h) * n5 / sqrt(denom * (double)n);
    } else {
        double ev[HO_MAXDIM];
        for (int d = 0; d < n; ++d) ev[d] = 0.0;
        for (int j = 0; j < 7; ++j) {
            const double c = HO_RK45_E[j];
            if (c != 0.0) {
                const double hc = h * c;
                const double *kj = k + j * n;
                for (int d = 0; d < n; ++d) ev[d] += hc * kj[d];
            }
        }
        for (int d = 0; d < n; ++d) ev[d] = ev[d] / scale[d];
        return np_norm(ev, n) / sqrt((double)n);
    }
}

/* ===================================================================================== */
/* Dense output                                                                          */
/* ===================================================================================== */
/* _dop853_build_dense_cache  rk.py:1836-1875 ; Kext must hold K[0..12] on entry */
static void dop853_dense_cache(const ho_system *sys, double t_old, const double *y_old, const double *f_old,
                               const double *y_new, const double *f_new, double h, double *Kext /*[16][n]*/,
                               double *F /*[7][n]*/)
{
    const int n = sys->dim;
    double ys[HO_MAXDIM];
    for (int srow = 13; srow < 16; ++srow) {
        for (int d = 0; d < n; ++d) {
            double acc = 0.0;
            for (int r = 0; r < srow; ++r) {
                const double a = HO_DOP853_A[srow][r];
                if (a != 0.0) acc += a * Kext[r * n + d];
            }
            ys[d] = y_old[d] + h * acc;
        }
        ho_rhs(sys, t_old + HO_DOP853_C[srow] * h, ys, Kext + srow * n);
    }
    for (int d = 0; d < n; ++d) {
        const double dy = y_new[d] - y_old[d];
        F[0 * n + d] = dy;
        F[1 * n + d] = h * f_old[d] - dy;
        F[2 * n + d] = 2.0 * dy - h * (f_new[d] + f_old[d]);
    }
    for (int i = 0; i < 4; ++i)
        for (int d = 0; d < n; ++d) {
            double acc = 0.0;
            for (int r = 0; r < 16; ++r) {
                const double c = HO_DOP853_D[i][r];
                if (c != 0.0) acc += c * Kext[r * n + d];
            }
            F[(3 + i) * n + d] = h * acc;
        }
}
/* _dop853_eval_dense  rk.py:1989-2003 */
static void dop853_eval(const double *y_old, const double *F, int n, double x, double *out)
{
    for (int d = 0; d < n; ++d) out[d] = 0.0;
    for (int i = 6; i >= 0; --i) {
        for (int d = 0; d < n; ++d) out[d] += F[i * n + d];
        if ((6 - i) % 2 == 0) { for (int d = 0; d < n; ++d) out[d] *= x; }
        else { const double omx = 1.0 - x; for (int d = 0; d < n; ++d) out[d] *= omx; }
    }
    for (int d = 0; d < n; ++d) out[d] += y_old[d];
}
/* _rk45_build_Q_cache rk.py:988-997 ; Q[d][c] */
static void rk45_q_cache(const double *K, int n, double *Q)
{
    for (int c = 0; c < 4; ++c) {
        for (int d = 0; d < n; ++d) Q[d * 4 + c] = 0.0;
        for (int r = 0; r < 7; ++r) {
            const double p = HO_RK45_P[r][c];
            if (p != 0.0) for (int d = 0; d < n; ++d) Q[d * 4 + c] += p * K[r * n + d];
        }
    }
}
/* _rk45_eval_dense rk.py:1021-1034 */
static void rk45_eval(const double *y_old, const double *Q, int n, double x, double h, double *out)
{
    double p[4], val = x;
    for (int c = 0; c < 4; ++c) { p[c] = val; val *= x; }
    for (int d = 0; d < n; ++d) {
        double acc = 0.0;
        for (int c = 0; c < 4; ++c) acc += Q[d * 4 + c] * p[c];
        out[d] = y_old[d] + h * acc;
    }
}
/* _hermite_eval_dense rk.py:296-312 */
static void hermite_eval(const double *y0, const double *f0, const double *y1, const double *f1, int n,
                         double x, double h, double *out)
{
    const double x2 = x * x, x3 = x2 * x;
    const double H00 = 2.0 * x3 - 3.0 * x2 + 1.0;
    const double H10 = x3 - 2.0 * x2 + x;
    const double H01 = -2.0 * x3 + 3.0 * x2;
    const double H11 = x3 - x2;
    for (int d = 0; d < n; ++d)
        out[d] = H00 * y0[d] + H10 * (h * f0[d]) + H01 * y1[d] + H11 * (h * f1[d]);
}

/* In-step bisection shared by _dop853_refine_in_step (rk.py:2079-2102),
 * _rk45_refine_in_step (rk.py:1072-1092) and _hermite_refine_in_step (rk.py:373-391). */
typedef struct {
    int kind;  /* 0 dop853, 1 rk45, 2 hermite */
    int n;
    const double *y0, *f0, *y1, *f1, *F, *Q;
    double h;
} dense_ctx;
static void dense_eval(const dense_ctx *c, double x, double *out)
{
    if (c->kind == 0) dop853_eval(c->y0, c->F, c->n, x, out);
    else if (c->kind == 1) rk45_eval(c->y0, c->Q, c->n, x, c->h, out);
    else hermite_eval(c->y0, c->f0, c->y1, c->f1, c->n, x, c->h, out);
}
static void refine_in_step(const dense_ctx *c, const ho_event *ev, double t0, double *t_hit, double *y_hit)
{
    double a = 0.0, b = 1.0;
    double g_left = ev_g(ev, c->y0);
    double ym[HO_MAXDIM];
    for (int it = 0; it < 128; ++it) {
        const double mid = 0.5 * (a + b);
        dense_eval(c, mid, ym);
        const double g_mid = ev_g(ev, ym);
        if (fabs(g_mid) <= ev->gtol) {
            *t_hit = t0 + mid * c->h;
            dense_eval(c, mid, y_hit);
            return;
        }
        if (crossed_direction(g_left, g_mid, ev->direction)) b = mid;
        else { a = mid; g_left = g_mid; }
        if ((b - a) * fabs(c->h) <= ev->xtol) break;
    }
    *t_hit = t0 + b * c->h;
    dense_eval(c, b, y_hit);
}

/* ===================================================================================== */
/* Adaptive drivers                                                                      */
/* ===================================================================================== */
typedef struct {
    double *ts, *ys, *dys, *Ks;
    int cap, n_nodes, n, kstride;
} node_store;

static void store_init(node_store *st, int n, int kstride)
{
    st->cap = 256; st->n_nodes = 0; st->n = n; st->kstride = kstride;
    st->ts = (double *)malloc(sizeof(double) * st->cap);
    st->ys = (double *)malloc(sizeof(double) * st->cap * n);
    st->dys = (double *)malloc(sizeof(double) * st->cap * n);
    st->Ks = (double *)malloc(sizeof(double) * st->cap * kstride);
}
static void store_grow(node_store *st)
{
    st->cap *= 2;
    st->ts = (double *)realloc(st->ts, sizeof(double) * st->cap);
    st->ys = (double *)realloc(st->ys, sizeof(double) * st->cap * st->n);
    st->dys = (double *)realloc(st->dys, sizeof(double) * st->cap * st->n);
    st->Ks = (double *)realloc(st->Ks, sizeof(double) * st->cap * st->kstride);
}
static void store_free(node_store *st) { free(st->ts); free(st->ys); free(st->dys); free(st->Ks); }

/* The shared adaptive loop (rk.py:2431-2484 / 1306-1353).  If st != NULL every accepted node and
 * its stage block are recorded for dense output. */
static void adaptive_loop(const ho_system *sys, int method, const ho_tol *tol, const double *y0, double t0,
                          double tf, double *y_out, node_store *st, int64_t *counts)
{
    const int n = sys->dim;
    const double order = (method == HO_DOP853) ? 8.0 : 5.0;
    rk_tab tb; get_tab(method, &tb);
    double k[HO_MAXSTAGE * HO_MAXDIM], y[HO_MAXDIM], yh[HO_MAXDIM], f0[HO_MAXDIM];
    double t = t0;
    memcpy(y, y0, n * sizeof(double));
    ho_rhs(sys, t, y, f0);
    if (st) {
        st->ts[0] = t;
        memcpy(st->ys, y, n * sizeof(double));
        memcpy(st->dys, f0, n * sizeof(double));
        st->n_nodes = 1;
    }
    double h = initial_step(y, f0, n, tol);
    double err_prev = -1.0;
    int64_t acc = 0, rej = 0;
    while ((t - tf) * 1.0 < 0.0) {
        h = clamp_step(h, tol->max_step, tol->min_step);
        h = adjust_to_endpoint(t, h, tf);
        const double err = adaptive_attempt(sys, method, &tb, tol, t, y, h, k, yh);
        if (err <= 1.0) {
            const double t_new = t + h;
            ++acc;
            if (st) {
                if (st->n_nodes == st->cap) store_grow(st);
                const int i = st->n_nodes;
                st->ts[i] = t_new;
                memcpy(st->ys + (size_t)i * n, yh, n * sizeof(double));
                ho_rhs(sys, t_new, yh, st->dys + (size_t)i * n);
                memcpy(st->Ks + (size_t)(i - 1) * st->kstride, k, sizeof(double) * (tb.s + 1) * n);
                st->n_nodes = i + 1;
            }
            t = t_new;
            memcpy(y, yh, n * sizeof(double));
            h = h * pi_accept_factor(err, err_prev, order);
            err_prev = err;
        } else {
            ++rej;
            h = h * pi_reject_factor(err, order);
            h = clamp_step(h, tol->max_step, tol->min_step);
        }
    }
    if (y_out) memcpy(y_out, y, n * sizeof(double));
    if (counts) { counts[0] = acc; counts[1] = rej; }
}

/* End state with the reference's API semantics: _propagate_dynsys(..., steps=2) returns the DENSE
 * interpolant evaluated at t_eval[-1] on the last accepted segment (rk.py:2503-2543), which differs
 * from the last node by an ulp (y_old + (y_new - y_old)) or, if t_new overshoots tf by an ulp, by a
 * genuine interpolation.  Only the last segment's node data is kept. */
int ho_adaptive_final(const ho_system *sys, int method, const ho_tol *tol, const double *y0, double t0,
                      double tf, double *yf, int64_t *counts)
{
    const int n = sys->dim;
    const double order = (method == HO_DOP853) ? 8.0 : 5.0;
    rk_tab tb; get_tab(method, &tb);
    double k[HO_MAXSTAGE * HO_MAXDIM], Kseg[HO_MAXSTAGE * HO_MAXDIM];
    double y[HO_MAXDIM], yh[HO_MAXDIM], fcur[HO_MAXDIM], y_old[HO_MAXDIM], f_old[HO_MAXDIM];
    double t = t0, t_old = t0, t_prev2 = t0;
    int have_seg = 0;
    memcpy(y, y0, n * sizeof(double));
    ho_rhs(sys, t, y, fcur);
    double h = initial_step(y, fcur, n, tol);
    double err_prev = -1.0;
    int64_t acc = 0, rej = 0;
    while ((t - tf) * 1.0 < 0.0) {
        h = clamp_step(h, tol->max_step, tol->min_step);
        h = adjust_to_endpoint(t, h, tf);
        const double err = adaptive_attempt(sys, method, &tb, tol, t, y, h, k, yh);
        if (err <= 1.0) {
            ++acc;
            t_prev2 = t_old;
            t_old = t;
            memcpy(y_old, y, n * sizeof(double));
            memcpy(f_old, fcur, n * sizeof(double));
            memcpy(Kseg, k, sizeof(double) * (tb.s + 1) * n);
            have_seg = 1;
            t = t + h;
            memcpy(y, yh, n * sizeof(double));
            ho_rhs(sys, t, y, fcur);
            h = h * pi_accept_factor(err, err_prev, order);
            err_prev = err;
        } else {
            ++rej;
            h = h * pi_reject_factor(err, order);
            h = clamp_step(h, tol->max_step, tol->min_step);
        }
    }
    (void)t_prev2;
    if (counts) { counts[0] = acc; counts[1] = rej; }
    if (!have_seg) { memcpy(yf, y, n * sizeof(double)); return 0; }
    /* searchsorted(ts, tf, 'right')-1 clipped to n_nodes-2 -> the last segment [t_old, t] */
    const double hseg = t - t_old;
    if (method == HO_DOP853 && hseg == 0.0) { memcpy(yf, y_old, n * sizeof(double)); return 0; }
    const double x = (tf - t_old) / hseg;
    if (method == HO_DOP853) {
        double F[7 * HO_MAXDIM];
        dop853_dense_cache(sys, t_old, y_old, f_old, y, fcur, hseg, Kseg, F);
        dop853_eval(y_old, F, n, x, yf);
    } else {
        double Q[4 * HO_MAXDIM];
        rk45_q_cache(Kseg, n, Q);
        rk45_eval(y_old, Q, n, x, hseg, yf);
    }
    return 0;
}

/* np.searchsorted(ts, tq, side='right') - 1, clipped (rk.py:2505-2509) */
static int seg_index(const double *ts, int n_nodes, double tq)
{
    int lo = 0, hi = n_nodes;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (tq < ts[mid]) hi = mid; else lo = mid + 1;
    }
    int j = lo - 1;
    if (j < 0) j = 0;
    if (j > n_nodes - 2) j = n_nodes - 2;
    return j;
}

int ho_adaptive_dense(const ho_system *sys, int method, const ho_tol *tol, const double *y0,
                      const double *t_eval, int m, double *y_out, int64_t *counts)
{
    const int n = sys->dim;
    const int kstride = 13 * n;
    node_store st; store_init(&st, n, kstride);
    adaptive_loop(sys, method, tol, y0, t_eval[0], t_eval[m - 1], NULL, &st, counts);
    double Kext[16 * HO_MAXDIM], F[7 * HO_MAXDIM], Q[4 * HO_MAXDIM];
    int last_j = -1;
    for (int idx = 0; idx < m; ++idx) {
        const double tq = t_eval[idx];
        const int j = seg_index(st.ts, st.n_nodes, tq);
        const double t0s = st.ts[j], t1s = st.ts[j + 1];
        const double hseg = t1s - t0s;
        const double *y_old = st.ys + (size_t)j * n;
        if (method == HO_DOP853 && hseg == 0.0) {
            memcpy(y_out + (size_t)idx * n, y_old, n * sizeof(double));
            continue;
        }
        const double x = (tq - t0s) / hseg;
        if (j != last_j) {
            if (method == HO_DOP853) {
                memcpy(Kext, st.Ks + (size_t)j * kstride, sizeof(double) * 13 * n);
                dop853_dense_cache(sys, t0s, y_old, st.dys + (size_t)j * n, st.ys + (size_t)(j + 1) * n,
                                   st.dys + (size_t)(j + 1) * n, hseg, Kext, F);
            } else {
                rk45_q_cache(st.Ks + (size_t)j * kstride, n, Q);
            }
            last_j = j;
        }
        if (method == HO_DOP853) dop853_eval(y_old, F, n, x, y_out + (size_t)idx * n);
        else rk45_eval(y_old, Q, n, x, hseg, y_out + (size_t)idx * n);
    }
    store_free(&st);
    return 0;
}

/* _integrate_dop853_until_event rk.py:2747-2803 ; _integrate_rk45_until_event rk.py:1542-1590 */
int ho_adaptive_event(const ho_system *sys, int method, const ho_tol *tol, const ho_event *ev,
                      const double *y0, double t0, double tmax, double *t_hit, double *y_hit,
                      double *y_last, int64_t *counts)
{
    const int n = sys->dim;
    const double order = (method == HO_DOP853) ? 8.0 : 5.0;
    rk_tab tb; get_tab(method, &tb);
    double Kext[HO_MAXSTAGE * HO_MAXDIM], y[HO_MAXDIM], yh[HO_MAXDIM], fc[HO_MAXDIM], fn[HO_MAXDIM];
    double F[7 * HO_MAXDIM], Q[4 * HO_MAXDIM];
    double t = t0;
    memcpy(y, y0, n * sizeof(double));
    ho_rhs(sys, t, y, fc);
    double g_prev = ev_g(ev, y);
    double h = initial_step(y, fc, n, tol);
    double err_prev = -1.0;
    int64_t acc = 0, rej = 0;
    int hit = 0;
    while ((t - tmax) * 1.0 < 0.0) {
        h = clamp_step(h, tol->max_step, tol->min_step);
        h = adjust_to_endpoint(t, h, tmax);
        const double err = adaptive_attempt(sys, method, &tb, tol, t, y, h, Kext, yh);
        if (err <= 1.0) {
            const double t_new = t + h;
            ++acc;
            if (method == HO_DOP853) ho_rhs(sys, t_new, yh, fn);
            const double g_new = ev_g(ev, yh);
            if (event_crossed(g_prev, g_new, ev->direction)) {
                dense_ctx c; memset(&c, 0, sizeof c);
                c.n = n; c.y0 = y; c.h = h;
                if (method == HO_DOP853) {
                    dop853_dense_cache(sys, t, y, fc, yh, fn, h, Kext, F);
                    c.kind = 0; c.F = F;
                } else {
                    rk45_q_cache(Kext, n, Q);
                    c.kind = 1; c.Q = Q;
                }
                refine_in_step(&c, ev, t, t_hit, y_hit);
                memcpy(y_last, yh, n * sizeof(double));
                hit = 1;
                break;
            }
            t = t_new;
            memcpy(y, yh, n * sizeof(double));
            if (method == HO_DOP853) memcpy(fc, fn, n * sizeof(double));
            else ho_rhs(sys, t, y, fc);
            g_prev = g_new;
            h = h * pi_accept_factor(err, err_prev, order);
            err_prev = err;
        } else {
            ++rej;
            h = h * pi_reject_factor(err, order);
            h = clamp_step(h, tol->max_step, tol->min_step);
        }
    }
    if (!hit) {
        *t_hit = t;
        memcpy(y_hit, y, n * sizeof(double));
        memcpy(y_last, y, n * sizeof(double));
    }
    if (counts) { counts[0] = acc; counts[1] = rej; }
    return hit;
}

/* ===================================================================================== */
/* Fixed-step drivers                                                                    */
/* ===================================================================================== */
/* _integrate_fixed_rk rk.py:572-588 */
int ho_fixed_dense(const ho_system *sys, int method, const double *y0, const double *t_vals, int m,
                   double *y_out)
{
    const int n = sys->dim;
    rk_tab tb; get_tab(method, &tb);
    double k[HO_MAXSTAGE * HO_MAXDIM], yh[HO_MAXDIM];
    memcpy(y_out, y0, n * sizeof(double));
    for (int idx = 0; idx < m - 1; ++idx) {
        const double tn = t_vals[idx];
        const double h = t_vals[idx + 1] - tn;
        rk_stages(sys, &tb, tn, y_out + (size_t)idx * n, h, k, yh);
        memcpy(y_out + (size_t)(idx + 1) * n, yh, n * sizeof(double));
    }
    return 0;
}

/* _integrate_fixed_rk_until_event rk.py:693-718 */
int ho_fixed_event(const ho_system *sys, int method, const ho_event *ev, const double *y0,
                   const double *t_vals, int m, double *t_hit, double *y_hit)
{
    const int n = sys->dim;
    rk_tab tb; get_tab(method, &tb);
    double k[HO_MAXSTAGE * HO_MAXDIM], y[HO_MAXDIM], yh[HO_MAXDIM], fp[HO_MAXDIM], fn[HO_MAXDIM];
    memcpy(y, y0, n * sizeof(double));
    ho_rhs(sys, t_vals[0], y, fp);
    double g_prev = ev_g(ev, y);
    for (int idx = 0; idx < m - 1; ++idx) {
        const double tn = t_vals[idx];
        const double h = t_vals[idx + 1] - tn;
        rk_stages(sys, &tb, tn, y, h, k, yh);
        ho_rhs(sys, tn + h, yh, fn);
        const double g_new = ev_g(ev, yh);
        if (event_crossed(g_prev, g_new, ev->direction)) {
            dense_ctx c; memset(&c, 0, sizeof c);
            c.kind = 2; c.n = n; c.y0 = y; c.f0 = fp; c.y1 = yh; c.f1 = fn; c.h = h;
            refine_in_step(&c, ev, tn, t_hit, y_hit);
            return 1;
        }
        memcpy(y, yh, n * sizeof(double));
        memcpy(fp, fn, n * sizeof(double));
        g_prev = g_new;
    }
    *t_hit = t_vals[m - 1];
    memcpy(y_hit, y, n * sizeof(double));
    return 0;
}

/* ---- _ExtendedSymplectic.integrate (symplectic.py:877-1004) --------------------------------------------------- */
/* _integrate_symplectic (symplectic.py:564-653) */
int ho_symplectic_dense(const ho_polyham *ham, const double *y0, const double *t_vals, int m, int order,
                        double c_omega, double *traj)
{
    if (m < 1 || order < 2 || (order % 2) != 0) return -1;
    double q[12];
    memcpy(traj, y0, 6 * sizeof(double));
    memcpy(q, y0, 6 * sizeof(double));
    memcpy(q + 6, y0, 6 * sizeof(double));
    for (int i = 0; i < m - 1; ++i) {
        const double dt = t_vals[i + 1] - t_vals[i];                  /* np.diff(t_values) */
        ho_tao_update(ham, q, dt, order, c_omega);
        memcpy(traj + (size_t)(i + 1) * 6, q, 6 * sizeof(double));
    }
    return 0;
}

/* _integrate_symplectic_until_event (symplectic.py:657-782) */
int ho_symplectic_event(const ho_polyham *ham, const ho_event *ev, const double *y0, const double *t_vals, int m,
                        int order, double c_omega, double *t_hit, double *y_hit, double *traj, int *n_rows)
{
    if (m < 1 || order < 2 || (order % 2) != 0) return -1;
    double q[12], y_old[6], y_new[6], f_old[6], f_new[6];
    memcpy(y_old, y0, sizeof y_old);
    if (traj) memcpy(traj, y0, 6 * sizeof(double));
    ho_polyham_rhs(ham, y_old, f_old);                                 /* _eval_hamiltonian_derivative :181-226 */
    double g_old = ev_g(ev, y_old);
    memcpy(q, y0, 6 * sizeof(double));
    memcpy(q + 6, y0, 6 * sizeof(double));
    for (int i = 0; i < m - 1; ++i) {
        const double dt = t_vals[i + 1] - t_vals[i];
        ho_tao_update(ham, q, dt, order, c_omega);
        memcpy(y_new, q, sizeof y_new);
        ho_polyham_rhs(ham, y_new, f_new);
        const double g_new = ev_g(ev, y_new);
        if (event_crossed(g_old, g_new, ev->direction)) {
            dense_ctx c; memset(&c, 0, sizeof c);
            c.kind = 2; c.n = 6; c.y0 = y_old; c.f0 = f_old; c.y1 = y_new; c.f1 = f_new; c.h = dt;
            refine_in_step(&c, ev, t_vals[i], t_hit, y_hit);
            *n_rows = i + 1;
            return 1;
        }
        if (traj) memcpy(traj + (size_t)(i + 1) * 6, y_new, sizeof y_new);
        memcpy(y_old, y_new, sizeof y_old);
        memcpy(f_old, f_new, sizeof f_old);
        g_old = g_new;
    }
    *t_hit = t_vals[m - 1];
    memcpy(y_hit, y_old, sizeof y_old);
    *n_rows = m;
    return 0;
}

/* ===================================================================================== */
/* Batch drivers (pthreads; a shared atomic cursor hands out trajectories)                */
/* ===================================================================================== */
typedef void (*ho_item_fn)(int64_t i, void *ctx);
typedef struct { ho_item_fn fn; void *ctx; int64_t n; atomic_llong *cursor; int64_t chunk; } pf_arg;

static void *pf_worker(void *p)
{
    pf_arg *a = (pf_arg *)p;
    for (;;) {
        const int64_t b = atomic_fetch_add(a->cursor, a->chunk);
        if (b >= a->n) break;
        const int64_t e = (b + a->chunk < a->n) ? b + a->chunk : a->n;
        for (int64_t i = b; i < e; ++i) a->fn(i, a->ctx);
    }
    return NULL;
}

void ho_parallel_for(int64_t n, int n_threads, int64_t chunk, ho_item_fn fn, void *ctx)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    atomic_llong cursor = 0;
    pf_arg a = { fn, ctx, n, &cursor, chunk < 1 ? 1 : chunk };
    if (n_threads == 1) { pf_worker(&a); return; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < n_threads - 1; ++i)
        if (pthread_create(&th[started], NULL, pf_worker, &a) == 0) ++started;
    pf_worker(&a);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

int ho_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

typedef struct {
    const ho_system *sys; int method; const ho_tol *tol; const double *y0; double t0, tf;
    const double *t_eval; int m; double *out; int64_t *counts;
} batch_ctx;

static void batch_final_item(int64_t i, void *p)
{
    batch_ctx *c = (batch_ctx *)p;
    const int dim = c->sys->dim;
    ho_adaptive_final(c->sys, c->method, c->tol, c->y0 + i * dim, c->t0, c->tf, c->out + i * dim,
                      c->counts ? c->counts + 2 * i : NULL);
}
static void batch_dense_item(int64_t i, void *p)
{
    batch_ctx *c = (batch_ctx *)p;
    const int dim = c->sys->dim;
    ho_adaptive_dense(c->sys, c->method, c->tol, c->y0 + i * dim, c->t_eval, c->m,
                      c->out + (size_t)i * c->m * dim, c->counts ? c->counts + 2 * i : NULL);
}

int ho_batch_final(const ho_system *sys, int method, const ho_tol *tol, const double *y0, int64_t n,
                   double t0, double tf, double *yf, int64_t *counts, int n_threads)
{
    batch_ctx c = { sys, method, tol, y0, t0, tf, NULL, 0, yf, counts };
    ho_parallel_for(n, n_threads, 4, batch_final_item, &c);
    return 0;
}

int ho_batch_dense(const ho_system *sys, int method, const ho_tol *tol, const double *y0, int64_t n,
                   const double *t_eval, int m, double *y_out, int64_t *counts, int n_threads)
{
    batch_ctx c = { sys, method, tol, y0, 0.0, 0.0, t_eval, m, y_out, counts };
    ho_parallel_for(n, n_threads, 1, batch_dense_item, &c);
    return 0;
}
