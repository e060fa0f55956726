/* ho_poly.c -- CPU restatement of the polynomial-Hamiltonian hot path (TEST INFRASTRUCTURE).
 * Reference (paths relative to hiten/):
 *   _poly_evaluate / _polynomial_evaluate   algorithms/polynomial/algebra.py:403-461, operations.py:551-583
 *   _hamiltonian_rhs                        algorithms/dynamics/hamiltonian.py:35-90
 *   _eval_dH_dQ / _eval_dH_dP               algorithms/integrators/symplectic.py:105-178
 *   Tao maps + recursion                    algorithms/integrators/symplectic.py:38-60, 370-560, 636-652
 *   _integrate_rk_ham / _poincare_step / _detect_crossing / _poincare_map
 *                                           algorithms/poincare/centermanifold/backend.py:35-382
 *   _hermite_scalar                         algorithms/poincare/utils.py:54-97
 * The reference evaluates in complex128 with |imag| <= 1e-17 coefficients and keeps the real part; with
 * real evaluation points every imaginary contribution to the real part is an exact zero, so real
 * arithmetic on the real parts reproduces it.
 */
#include "ho_poly.h"
#include "ho_coeffs.h"

#include <math.h>
#include <string.h>

#define MAXDEG 32

double ho_poly_eval_partial(const ho_polyham *ham, int p, const double *pt6)
{
    double pw[6][MAXDEG + 1];
    const int D = ham->max_deg;
    for (int v = 0; v < 6; ++v) {
        pw[v][0] = 1.0;
        for (int e = 1; e <= D; ++e) pw[v][e] = pw[v][e - 1] * pt6[v];
    }
    double total = 0.0, acc = 0.0;
    int dcur = -1;
    for (int64_t i = ham->ptr[p]; i < ham->ptr[p + 1]; ++i) {
        if (ham->deg[i] != dcur) {
            if (dcur >= 0) total += acc;
            acc = 0.0;
            dcur = ham->deg[i];
        }
        double term = 1.0;
        const int32_t *e = ham->exp + 6 * i;
        for (int v = 0; v < 6; ++v) term *= pw[v][e[v]];
        acc += ham->coef[i] * term;
    }
    if (dcur >= 0) total += acc;
    return total;
}

static void grad(const ho_polyham *ham, const double *Q, const double *P, double *dHdQ, double *dHdP)
{
    double pt[6] = { Q[0], Q[1], Q[2], P[0], P[1], P[2] };
    for (int i = 0; i < 3; ++i) {
        if (dHdQ) dHdQ[i] = ho_poly_eval_partial(ham, i, pt);
        if (dHdP) dHdP[i] = ho_poly_eval_partial(ham, 3 + i, pt);
    }
}

void ho_polyham_rhs(const ho_polyham *ham, const double *y, double *dy)
{
    double dq[3], dp[3];
    grad(ham, y, y + 3, dq, dp);
    for (int i = 0; i < 3; ++i) { dy[i] = dp[i]; dy[3 + i] = -dq[i]; }
}

/* ---- Tao extended-phase-space maps: q_ext = [Q, P, X, Y] ------------------------------------- */
static void phi_a(const ho_polyham *ham, double *q, double delta)          /* symplectic.py:399-411 */
{
    double dq[3], dp[3];
    grad(ham, q + 0, q + 9, dq, dp);                 /* (Q, Y) */
    for (int i = 0; i < 3; ++i) { q[3 + i] -= delta * dq[i]; q[6 + i] += delta * dp[i]; }
}
static void phi_b(const ho_polyham *ham, double *q, double delta)          /* symplectic.py:443-455 */
{
    double dq[3], dp[3];
    grad(ham, q + 6, q + 3, dq, dp);                 /* (X, P) */
    for (int i = 0; i < 3; ++i) { q[0 + i] += delta * dp[i]; q[9 + i] -= delta * dq[i]; }
}
static void phi_c(double *q, double delta, double omega)                   /* symplectic.py:481-506 */
{
    const double c = cos(2 * omega * delta), s = sin(2 * omega * delta);
    for (int i = 0; i < 3; ++i) {
        const double Q = q[i], P = q[3 + i], X = q[6 + i], Y = q[9 + i];
        const double qpx = Q + X, qmx = Q - X, ppy = P + Y, pmy = P - Y;
        q[i] = 0.5 * (qpx + c * qmx + s * pmy);
        q[3 + i] = 0.5 * (ppy - s * qmx + c * pmy);
        q[6 + i] = 0.5 * (qpx - c * qmx - s * pmy);
        q[9 + i] = 0.5 * (ppy + s * qmx - c * pmy);
    }
}
static void tao_recursive(const ho_polyham *ham, double *q, double ts, int order, double omega) /* :543-560 */
{
    if (order == 2) {
        phi_a(ham, q, 0.5 * ts);
        phi_b(ham, q, 0.5 * ts);
        phi_c(q, ts, omega);
        phi_b(ham, q, 0.5 * ts);
        phi_a(ham, q, 0.5 * ts);
    } else {
        const double gamma = 1.0 / (2.0 - pow(2.0, 1.0 / ((double)order + 1.0)));
        tao_recursive(ham, q, gamma * ts, order - 2, omega);
        tao_recursive(ham, q, (1.0 - 2.0 * gamma) * ts, order - 2, omega);
        tao_recursive(ham, q, gamma * ts, order - 2, omega);
    }
}
/* one _recursive_update_poly call on an extended state q_ext[12] = [Q,P,X,Y] (symplectic.py:509-560); exported for the
 * grid / event drivers of _ExtendedSymplectic.integrate in hiten_oracle.c */
void ho_tao_update(const ho_polyham *ham, double *q_ext, double dt, int order, double c_omega)
{
    const double omega = pow(c_omega * dt, -(double)order);     /* _get_tao_omega (symplectic.py:38-60) */
    tao_recursive(ham, q_ext, dt, order, omega);
}
/* _integrate_symplectic over t_vals = [0, dt]  (symplectic.py:636-652) */
static void tao_step(const ho_polyham *ham, const double *y, double dt, int order, double c_omega, double *y_new)
{
    double q[12];
    memcpy(q, y, 6 * sizeof(double));
    memcpy(q + 6, y, 6 * sizeof(double));
    const double step = dt - 0.0;
    const double omega = pow(c_omega * step, -(double)order);
    tao_recursive(ham, q, step, order, omega);
    memcpy(y_new, q, 6 * sizeof(double));
}

/* _integrate_rk_ham over t_vals = [0, dt]  (centermanifold/backend.py:144-184) */
static void rk_ham_step(const ho_polyham *ham, const double *y, double dt, int order, double *y_new)
{
    const double *A, *B;
    int S;
    if (order == 4) { A = &HO_RK4_A[0][0]; B = HO_RK4_B; S = 4; }
    else if (order == 6) { A = &HO_RK6_A[0][0]; B = HO_RK6_B; S = 7; }
    else { A = &HO_RK8_A[0][0]; B = HO_RK8_B; S = 13; }
    double k[13][6], ys[6];
    const double h = dt - 0.0;
    for (int s = 0; s < S; ++s) {
        memcpy(ys, y, sizeof ys);
        for (int j = 0; j < s; ++j) {
            const double a = A[s * S + j];
            if (a != 0.0) { const double ha = h * a; for (int d = 0; d < 6; ++d) ys[d] += ha * k[j][d]; }
        }
        double dq[3], dp[3];
        grad(ham, ys, ys + 3, dq, dp);
        for (int i = 0; i < 3; ++i) { k[s][i] = dp[i]; k[s][3 + i] = -dq[i]; }
    }
    memcpy(y_new, y, 6 * sizeof(double));
    for (int s = 0; s < S; ++s) {
        const double b = B[s];
        if (b != 0.0) { const double hb = h * b; for (int d = 0; d < 6; ++d) y_new[d] += hb * k[s][d]; }
    }
}

static double hermite_scalar(double s, double y0, double y1, double dy0, double dy1, double dt)  /* poincare/utils.py:93-97 */
{
    const double oms = 1.0 - s;
    const double h00 = (1.0 + 2.0 * s) * (oms * oms);
    const double h10 = s * (oms * oms);
    const double h01 = (s * s) * (3.0 - 2.0 * s);
    const double h11 = (s * s) * (s - 1.0);
    return h00 * y0 + h10 * dy0 * dt + h01 * y1 + h11 * dy1 * dt;
}

/* _poincare_step (backend.py:280-311); section: 0 q2, 1 p2, 2 q3, 3 p3 */
static int poincare_step(const ho_polyham *ham, const double *seed, double dt, int order, int max_steps,
                         int use_symplectic, int section, double c_omega, double *out, double *t_cross)
{
    static const int fidx[4] = { 1, 4, 2, 5 };      /* position of q2, p2, q3, p3 in [q1,q2,q3,p1,p2,p3] */
    double so[6] = { 0, seed[0], seed[2], 0, seed[1], seed[3] }, sn[6], rn[6], ro[6];
    double elapsed = 0.0;
    for (int it = 0; it < max_steps; ++it) {
        if (use_symplectic) tao_step(ham, so, dt, order, c_omega, sn);
        else rk_ham_step(ham, so, dt, order, sn);
        ho_polyham_rhs(ham, sn, rn);
        const double f_old = so[fidx[section]], f_new = sn[fidx[section]];
        int crossed = 0;
        if (!(f_old * f_new >= 0.0)) {
            int good;
            if (section == 2) good = sn[5] > 0.0;         /* q3: p3_new > 0 */
            else if (section == 0) good = sn[4] > 0.0;    /* q2: p2_new > 0 */
            else if (section == 3) good = rn[2] > 0.0;    /* p3: dq3/dt > 0 */
            else good = rn[1] > 0.0;                      /* p2: dq2/dt > 0 */
            crossed = good;
        }
        if (crossed) {
            const double alpha = f_old / (f_old - f_new);
            ho_polyham_rhs(ham, so, ro);
            out[0] = hermite_scalar(alpha, so[1], sn[1], ro[1], rn[1], dt);
            out[1] = hermite_scalar(alpha, so[4], sn[4], ro[4], rn[4], dt);
            out[2] = hermite_scalar(alpha, so[2], sn[2], ro[2], rn[2], dt);
            out[3] = hermite_scalar(alpha, so[5], sn[5], ro[5], rn[5], dt);
            *t_cross = elapsed + alpha * dt;
            return 1;
        }
        memcpy(so, sn, sizeof so);
        elapsed += dt;
    }
    out[0] = out[1] = out[2] = out[3] = 0.0;
    *t_cross = 0.0;
    return 0;
}

typedef struct {
    const ho_polyham *ham; const double *seeds; double dt; int order, max_steps, symp, section; double c_omega;
    int64_t *flags; double *out, *t_out;
} cm_ctx;

void ho_parallel_for(int64_t n, int n_threads, int64_t chunk, void (*fn)(int64_t, void *), void *ctx);

static void cm_item(int64_t i, void *p)
{
    cm_ctx *c = (cm_ctx *)p;
    c->flags[i] = poincare_step(c->ham, c->seeds + 4 * i, c->dt, c->order, c->max_steps, c->symp, c->section,
                                c->c_omega, c->out + 4 * i, c->t_out + i);
}

int ho_cm_poincare_map(const ho_polyham *ham, const double *seeds, int64_t n, double dt, int order, int max_steps,
                       int use_symplectic, int section, double c_omega, int64_t *flags, double *out, double *t_out,
                       int n_threads)
{
    if (ham->max_deg > MAXDEG) return -1;
    cm_ctx c = { ham, seeds, dt, order, max_steps, use_symplectic, section, c_omega, flags, out, t_out };
    ho_parallel_for(n, n_threads, 1, cm_item, &c);
    return 0;
}


/* ---- seed lifting ------------------------------------------------------------------------------- */
typedef struct { const ho_polyham *H; double st[6]; int idx; double h0; } resid_ctx;

static double residual(resid_ctx *c, double x)                     /* interfaces.py:232-238 */
{
    c->st[c->idx] = x;
    return ho_poly_eval_partial(c->H, 0, c->st) - c->h0;
}

/* solve_bracketed_brent, rootfinding.py:92-190; returns 1 and *root, or 0 (None) */
static int brent(resid_ctx *c, double a, double b, double xtol, int max_iter, double *root)
{
    double fa = residual(c, a), fb = residual(c, b);
    if (fa == 0.0) { *root = a; return 1; }
    if (fb == 0.0) { *root = b; return 1; }
    if (fa * fb > 0.0) return 0;
    double cc = a, fc = fa, d = b - a, e = d;
    const double eps = 2.220446049250313e-16;
    double tol, m;
    for (int it = 0; it < max_iter; ++it) {
        if (fb == 0.0) { *root = b; return 1; }
        if (fb * fc > 0.0) { cc = a; fc = fa; d = b - a; e = d; }
        if (fabs(fc) < fabs(fb)) {
            a = b; b = cc; cc = a;                               /* a, b, c = b, c, b */
            fa = fb; fb = fc; fc = fa;
        }
        tol = 2.0 * eps * fabs(b) + 0.5 * xtol;
        m = 0.5 * (cc - b);
        if (fabs(m) <= tol) { *root = b; return 1; }
        if (fabs(e) >= tol && fabs(fa) > fabs(fb)) {
            const double s = fb / fa;
            double p, q;
            if (a == cc) {
                p = 2.0 * m * s;
                q = 1.0 - s;
            } else {
                const double q_ = fa / fc, r = fb / fc;
                p = s * (2.0 * m * q_ * (q_ - r) - (b - a) * (r - 1.0));
                q = (q_ - 1.0) * (r - 1.0) * (s - 1.0);
            }
            if (p > 0.0) q = -q; else p = -p;
            const double lim1 = 3.0 * m * q - fabs(tol * q), lim2 = fabs(e * q);
            if ((2.0 * p) < (lim2 < lim1 ? lim2 : lim1)) { e = d; d = p / q; }   /* Python min(): first arg unless second is smaller */
            else { d = m; e = m; }
        } else { d = m; e = m; }
        a = b; fa = fb;
        if (fabs(d) > tol) b = b + d;
        else b = b + (m > 0.0 ? tol : -tol);
        fb = residual(c, b);
    }
    tol = 2.0 * eps * fabs(b) + 0.5 * xtol;
    m = 0.5 * (cc - b);
    if (fabs(m) <= tol || fb == 0.0) { *root = b; return 1; }
    return 0;
}

int ho_cm_solve_missing(const ho_polyham *H, const double *fixed6, int solve_idx, double h0, double initial_guess,
                        double expand_factor, int max_expand, int symmetric, double xtol, double *root)
{
    resid_ctx c;
    c.H = H; c.idx = solve_idx; c.h0 = h0;
    memcpy(c.st, fixed6, sizeof c.st);
    if (residual(&c, 0.0) > 0.0) return 0;
    double b = initial_guess, r_b = residual(&c, b);
    int n_expand = 0;
    while (r_b <= 0.0 && n_expand < max_expand) { b *= expand_factor; r_b = residual(&c, b); ++n_expand; }
    if (r_b > 0.0) return brent(&c, 0.0, b, xtol, 200, root);
    if (symmetric) {
        double a_neg = -initial_guess, r_a = residual(&c, a_neg);
        n_expand = 0;
        while (r_a <= 0.0 && n_expand < max_expand) { a_neg *= expand_factor; r_a = residual(&c, a_neg); ++n_expand; }
        if (r_a > 0.0) return brent(&c, a_neg, 0.0, xtol, 200, root);
    }
    return 0;
}

int ho_cm_lift(const ho_polyham *H, int section, const double *pts, int64_t n, double h0, double initial_guess,
               double expand_factor, int max_expand, int symmetric, double xtol, int64_t *ok, double *states)
{
    if (section < 0 || section > 3) return -1;
    /* variable slots: q1 0, q2 1, q3 2, p1 3, p2 4, p3 5; output order (q2, p2, q3, p3) */
    static const int plane_a[4] = {2, 2, 1, 1}, plane_b[4] = {5, 5, 4, 4};      /* plane coordinates per section */
    static const int missing[4] = {4, 1, 5, 2};                                  /* q2->p2, p2->q2, q3->p3, p3->q3 */
    static const int out_slot[6] = {-1, 0, 2, -1, 1, 3};
    for (int64_t i = 0; i < n; ++i) {
        double fixed[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, root = 0.0;
        fixed[plane_a[section]] = pts[2 * i];
        fixed[plane_b[section]] = pts[2 * i + 1];
        ok[i] = ho_cm_solve_missing(H, fixed, missing[section], h0, initial_guess, expand_factor, max_expand, symmetric,
                                    xtol, &root);
        double *o = states + 4 * i;
        o[0] = o[1] = o[2] = o[3] = 0.0;
        if (ok[i]) {
            o[out_slot[plane_a[section]]] = pts[2 * i];
            o[out_slot[plane_b[section]]] = pts[2 * i + 1];
            o[out_slot[missing[section]]] = root;
        }
    }
    return 0;
}
