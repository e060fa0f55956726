/* ho_correct.c -- CPU restatement of single-shooting differential correction of periodic orbits
 * (TEST INFRASTRUCTURE, see hiten_oracle.h).
 *
 * Reference (paths relative to src/hiten/):
 *   _NewtonBackend.run                      algorithms/corrector/backends/newton.py:20-150
 *   _CorrectorBackend._compute_jacobian     algorithms/corrector/backends/base.py (central differences)
 *   _CorrectorBackend._solve_delta_dense    algorithms/corrector/backends/base.py (cond test, 1e-12 ridge, solve)
 *   _ArmijoLineSearch.__call__              algorithms/corrector/stepping/armijo.py:60-170
 *   _CorrectorPlainStep                     algorithms/corrector/stepping/plain.py
 *   _SingleShootingOrbitOperators           algorithms/corrector/operators.py:319-452 (residual / Jacobian)
 *   _SingleHitBackend._cross / _cross_event_driven   algorithms/poincare/singlehit/backend.py:164-282
 *   _HaloOrbitCorrectionService._halo_quadratic_term algorithms/types/services/orbits.py:917-946
 *
 * Parity: the event propagation is the bit-exact 6-state DOP853 path; the STM comes from the 42-state path
 * (5e-13 relative vs the reference, see hiten_oracle.c), so for the analytic-Jacobian families corrected states agree with the reference to Newton-convergence level (checked against
 * tests/golden/correction.npz: |dx| <= 1e-10, iteration counts within one -- the reference's |R| < 1e-12 test
 * sits on the 1e-12 noise floor of its own event solver), not bit for bit.  The finite-difference family (vertical
 * orbits) has no STM in the loop and the 2x2 solve reproduces LAPACK's roundings: bit-exact (25 iterations).
 */
#include <math.h>
#include <string.h>

#include "hiten_oracle.h"

#define HO_PI 3.141592653589793

/* _plane_crossing_factory(...)(dynsys, x0, forward=1) with t_guess = None: returns 1 and (t_hit, x_hit) or 0. */
static int cross_window(double mu, const double *x0, int idx, double offset, double t0, double tmax, double *t_hit,
                        double *x_hit)
{
    ho_system sys = {HO_SYS_CR3BP6, 6, mu, 1, -1, -1, 0};
    ho_tol tol = {1e-12, 1e-12, 1e4, 10.0 * 2.220446049250313e-16};
    double t_start = t0;
    if (t_start <= 0.0) t_start = 1e-12;                       /* backend.py:205-207 */
    double te[2] = {0.0, t_start}, ya[12];
    int64_t counts[2];
    if (fabs(0.0 - t_start) <= 1e-8 + 1e-5 * fabs(t_start))           /* np.isclose(t_eval[0], t_eval[-1]): */
        memcpy(ya + 6, x0, 6 * sizeof(double));                        /* zero-length shortcut, base.py:420-424 */
    else
        ho_adaptive_dense(&sys, HO_DOP853, &tol, x0, te, 2, ya, counts);   /* _propagate_dynsys(..., steps=2) */
    double span = tmax - t_start;
    if (span < 0.0) span = 0.0;
    ho_tol etol = {1e-12, 1e-12, 1e300, 10.0 * 2.220446049250313e-16}; /* RungeKutta(order=853, rtol, atol) */
    ho_event ev = {idx, offset, 0, 1e-12, 1e-12};
    double t_rel, y_last[6];
    int hit = ho_adaptive_event(&sys, HO_DOP853, &etol, &ev, ya + 6, 0.0, span, &t_rel, x_hit, y_last, counts);
    if (!hit || !(t_rel < span && t_rel >= 0.0)) return 0;
    *t_hit = t_start + t_rel;
    return 1;
}

int ho_plane_crossing(double mu, const double *x0, int idx, double offset, double *t_hit, double *x_hit)
{
    /* _cross, t_guess None: first window [0, pi], fallback [pi/2 - 0.15, + pi] (backend.py:236-282) */
    double t_start = HO_PI / 2.0 - 0.15;
    double half_span = HO_PI * 0.5;
    double t0 = t_start - half_span;
    if (t0 < 0.0) t0 = 0.0;
    double tmax = t0 + 2.0 * half_span;
    if (cross_window(mu, x0, idx, offset, t0, tmax, t_hit, x_hit)) return 1;
    return cross_window(mu, x0, idx, offset, t_start, t_start + HO_PI, t_hit, x_hit);
}

static int residual(double mu, const ho_correct_opts *o, const double *base, const double *p, double *r, double *t_ev,
                    double *x_ev)
{
    double x[6];
    memcpy(x, base, sizeof x);
    x[o->ctrl[0]] = p[0];
    x[o->ctrl[1]] = p[1];
    if (!ho_plane_crossing(mu, x, o->event_idx, o->event_offset, t_ev, x_ev)) return 0;
    r[0] = x_ev[o->res[0]] - o->target[0];
    r[1] = x_ev[o->res[1]] - o->target[1];
    return 1;
}

static double norm_inf2(const double *r) { return fmax(fabs(r[0]), fabs(r[1])); }

/* _solve_delta_dense for a 2x2 system: delta = solve(J (+1e-12 I), -r). Returns 0 if singular. */
int ho_solve_delta2(const double *J, const double *r, double *delta)
{
    double a = J[0], b = J[1], c = J[2], d = J[3];
    double s = a * a + b * b + c * c + d * d, det = a * d - b * c;
    double smax2 = 0.5 * (s + sqrt(fmax(s * s - 4.0 * det * det, 0.0)));
    double cond = smax2 / fabs(det);                          /* sigma_max / sigma_min */
    if (isnan(cond) || cond > 1e8) { a += 1e-12; d += 1e-12; }
    double b0 = -r[0], b1 = -r[1];
    if (fabs(c) > fabs(a)) {                                  /* partial pivoting */
        double t;
        t = a; a = c; c = t;
        t = b; b = d; d = t;
        t = b0; b0 = b1; b1 = t;
    }
    if (a == 0.0) return 0;
    /* np.linalg.solve = LAPACK dgesv (dgetrf2 + dgetrs, OpenBLAS kernels); the roundings below were identified against
     * numpy on 4000 random systems (0 mismatches): the multiplier is c * (1/a), the Schur update a separately rounded
     * mul + sub, both substitutions fused multiply-adds, the two final quotients true divisions */
    double l = c * (1.0 / a);
    double u22 = d - l * b;
    if (u22 == 0.0) return 0;
    double y1 = fma(-l, b0, b1);
    delta[1] = y1 / u22;
    delta[0] = fma(-b, delta[1], b0) / a;
    return 1;
}

static void jacobian(double mu, const ho_correct_opts *o, const double *base, const double *p, double t_ev,
                     const double *x_ev, double *J, int *ok)
{
    *ok = 1;
    if (o->finite_difference) {                               /* base.py _compute_jacobian */
        for (int i = 0; i < 2; ++i) {
            double pp[2] = {p[0], p[1]}, pm[2] = {p[0], p[1]}, rp[2], rm[2], t, xe[6];
            double h = o->fd_step * fmax(1.0, fabs(p[i]));
            pp[i] += h;
            pm[i] -= h;
            if (!residual(mu, o, base, pp, rp, &t, xe) || !residual(mu, o, base, pm, rm, &t, xe)) { *ok = 0; return; }
            J[0 * 2 + i] = (rp[0] - rm[0]) / (2.0 * h);
            J[1 * 2 + i] = (rp[1] - rm[1]) / (2.0 * h);
        }
        return;
    }
    /* operators.py:437-450: Phi(t_event) from _compute_stm(var_dynsys, x_full, t_event, steps) */
    double x[6], phi0[42], phi[42];
    memcpy(x, base, sizeof x);
    x[o->ctrl[0]] = p[0];
    x[o->ctrl[1]] = p[1];
    memset(phi0, 0, sizeof phi0);
    for (int i = 0; i < 6; ++i) { phi0[7 * i] = 1.0; phi0[36 + i] = x[i]; }
    ho_system sys = {HO_SYS_VAR42, 42, mu, 1, -1, -1, 0};
    ho_tol tol = {1e-12, 1e-12, 1e4, 10.0 * 2.220446049250313e-16};
    int64_t counts[2];
    ho_adaptive_final(&sys, HO_DOP853, &tol, phi0, 0.0, t_ev, phi, counts);
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) J[2 * a + b] = phi[6 * o->res[a] + o->ctrl[b]];
    if (o->halo_quadratic) {                                  /* orbits.py:917-946 */
        double X = x_ev[0], Y = x_ev[1], Z = x_ev[2], vy = x_ev[4];
        double mu2 = 1 - mu;
        double rho_1 = 1 / pow((X + mu) * (X + mu) + Y * Y + Z * Z, 1.5);
        double rho_2 = 1 / pow((X - mu2) * (X - mu2) + Y * Y + Z * Z, 1.5);
        double omega_x = -(mu2 * (X + mu) * rho_1) - (mu * (X - mu2) * rho_2) + X;
        double DD[2] = {2 * vy + omega_x, -(mu2 * Z * rho_1) - (mu * Z * rho_2)};
        if (fabs(vy) < 1e-9) vy = vy != 0 ? copysign(1e-9, vy) : 1e-9;
        const int cols[2] = {0, 4};                           /* hard-wired (X, VY) in the reference */
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) J[2 * a + b] -= DD[a] * phi[6 * 1 + cols[b]] / vy;
    }
}

/* status: 0 converged, 1 max_attempts exhausted, 2 step strategy failed (ConvergenceError), 3 no event for the
 * current iterate (the reference raises), 4 singular Jacobian */
int ho_correct_orbit(double mu, const ho_correct_opts *o, const double *x0, double *x_out, double *half_period,
                     int *iterations, double *residual_norm)
{
    double p[2] = {x0[o->ctrl[0]], x0[o->ctrl[1]]}, r[2], t_ev = 0.0, x_ev[6];
    int status = 1, k;
    memcpy(x_out, x0, 6 * sizeof(double));
    *half_period = NAN;
    double r_norm = NAN;
    for (k = 0; k <= o->max_attempts; ++k) {
        if (!residual(mu, o, x0, p, r, &t_ev, x_ev)) { status = 3; break; }
        r_norm = norm_inf2(r);
        if (r_norm < o->tol) { status = 0; break; }
        if (k == o->max_attempts) break;
        double J[4], delta[2];
        int ok;
        jacobian(mu, o, x0, p, t_ev, x_ev, J, &ok);
        if (!ok) { status = 3; break; }
        if (!ho_solve_delta2(J, r, delta)) { status = 4; break; }
        /* step cap (armijo.py:98-107, plain.py) */
        double scale = 1.0;
        if (!isinf(o->max_delta)) {
            double dn = norm_inf2(delta);
            if (dn > o->max_delta) {
                scale = o->max_delta / dn;
                delta[0] = delta[0] * scale;
                delta[1] = delta[1] * scale;
            }
        }
        if (!o->line_search) {
            p[0] = p[0] + delta[0];
            p[1] = p[1] + delta[1];
            continue;
        }
        double alpha = 1.0, best_p[2] = {p[0], p[1]}, best_norm = r_norm, best_alpha = 0.0;
        int accepted = 0;
        while (alpha >= o->min_alpha) {
            double pt[2] = {p[0] + alpha * delta[0], p[1] + alpha * delta[1]}, rt[2], tt, xe[6];
            if (residual(mu, o, x0, pt, rt, &tt, xe)) {
                double nt = norm_inf2(rt);
                if (nt <= (1.0 - o->armijo_c * alpha) * r_norm) {
                    p[0] = pt[0];
                    p[1] = pt[1];
                    accepted = 1;
                    break;
                }
                if (nt < best_norm) { best_p[0] = pt[0]; best_p[1] = pt[1]; best_norm = nt; best_alpha = alpha; }
            }
            alpha *= o->alpha_reduction;
        }
        if (!accepted) {
            if (best_alpha > 0) { p[0] = best_p[0]; p[1] = best_p[1]; }
            else { status = 2; break; }
        }
    }
    x_out[o->ctrl[0]] = p[0];
    x_out[o->ctrl[1]] = p[1];
    *iterations = k > o->max_attempts ? o->max_attempts : k;
    *residual_norm = r_norm;
    if (status == 0) {
        /* _half_period (interfaces.py): event time of the corrected state = the last residual evaluation's */
        *half_period = t_ev;
    }
    return status;
}
