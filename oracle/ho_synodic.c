/* ho_synodic.c -- CPU restatement of the synodic-section detector (TEST INFRASTRUCTURE).
 * Reference: hiten/algorithms/poincare/synodic/backend.py
 *   detect_on_trajectory :687-821, _detect_with_segment_refine :458-659 (linear branch),
 *   _on_surface_indices :134-183, _crossing_indices_and_alpha :184-233, _refine_hits_linear :234-273,
 *   _order_and_dedup_hits :382-455.
 * The shipped defaults always select the LINEAR branch: ho_synodic_detect; the cubic branch
 * (reached by calling the backend with interp_kind="cubic") is ho_synodic_detect_cubic below.
 * direction: 0 = None, +1, -1.  Returns the number of hits written (<= cap).
 */
#include "hiten_oracle.h"

#include <math.h>
#include <string.h>

typedef struct {
    double *t, *x;
    int cap, n, dim, pi, pj, max_hits;
    double dtt, dpt;
    double last_t, last_u, last_v;
} hit_sink;

/* _order_and_dedup_hits: candidates arrive already ordered by segment */
static int sink_push(hit_sink *s, double th, const double *xh)
{
    if (s->max_hits > 0 && s->n >= s->max_hits) return 0;
    if (s->n > 0) {
        if (fabs(th - s->last_t) <= s->dtt) return 1;
        const double du = xh[s->pi] - s->last_u, dv = xh[s->pj] - s->last_v;
        if ((du * du + dv * dv) <= (s->dpt * s->dpt)) return 1;
    }
    if (s->n < s->cap) {
        s->t[s->n] = th;
        memcpy(s->x + (size_t)s->n * s->dim, xh, sizeof(double) * s->dim);
    }
    s->last_t = th; s->last_u = xh[s->pi]; s->last_v = xh[s->pj];
    s->n++;
    return 1;
}

int ho_synodic_detect(const double *times, const double *states, int m, int dim, int idx, double offset,
                      int direction, int proj_i, int proj_j, int segment_refine, double tol_on_surface,
                      double dedup_time_tol, double dedup_point_tol, int max_hits, double *hit_times,
                      double *hit_states, int cap)
{
    hit_sink s = { hit_times, hit_states, cap, 0, dim, proj_i, proj_j, max_hits, dedup_time_tol, dedup_point_tol, 0, 0, 0 };
    if (m < 2) return 0;
    const int r = segment_refine;
    double xh[64];
#define G(k) (states[(size_t)(k) * dim + idx] - offset)
    for (int k = 0; k < m - 1; ++k) {
        const double t0 = times[k], t1 = times[k + 1];
        const double gk = G(k), gk1 = G(k + 1);
        const double *x0 = states + (size_t)k * dim, *x1 = states + (size_t)(k + 1) * dim;
        int accept_left = 0;
        if (fabs(gk) < tol_on_surface) {
            if (direction == 0) accept_left = 1;
            else if (direction == 1) accept_left = (gk1 >= 0.0) || ((k - 1 >= 0) && (G(k - 1) <= 0.0));
            else accept_left = (gk1 <= 0.0) || ((k - 1 >= 0) && (G(k - 1) >= 0.0));
        }
        if (r > 0) {
            if (accept_left && !sink_push(&s, t0, x0)) return s.n < cap ? s.n : cap;
            const double step = 1.0 / (double)(r + 1);
            for (int mm = 0; mm <= r; ++mm) {
                const double s_lo = (double)mm * step, s_hi = (double)(mm + 1) * step;
                if (s_hi > 1.0 + 1e-15) break;
                if (accept_left && mm == 0) continue;
                const double g_lo = (1.0 - s_lo) * gk + s_lo * gk1;
                const double g_hi = (1.0 - s_hi) * gk + s_hi * gk1;
                int crosses;
                if (direction == 0) crosses = (g_lo * g_hi <= 0.0) && (g_lo != g_hi);
                else if (direction == 1) crosses = (g_lo < 0.0) && (g_hi >= 0.0);
                else crosses = (g_lo > 0.0) && (g_hi <= 0.0);
                if (!crosses) continue;
                double s_star;
                if (g_lo == g_hi) s_star = 0.5 * (s_lo + s_hi);
                else {
                    double al = g_lo / (g_lo - g_hi);
                    al = fmin(1.0, fmax(0.0, al));
                    s_star = s_lo + al * (s_hi - s_lo);
                }
                const double th = (1.0 - s_star) * t0 + s_star * t1;
                for (int d = 0; d < dim; ++d) xh[d] = x0[d] + s_star * (x1[d] - x0[d]);
                if (!sink_push(&s, th, xh)) return s.n < cap ? s.n : cap;
            }
        } else {
            if (accept_left) {
                if (!sink_push(&s, t0, x0)) break;
                continue;
            }
            int crosses;
            if (direction == 0) crosses = (gk * gk1 <= 0.0) && (gk != gk1);
            else if (direction == 1) crosses = (gk < 0.0) && (gk1 >= 0.0);
            else crosses = (gk > 0.0) && (gk1 <= 0.0);
            if (!crosses) continue;
            double al = gk / (gk - gk1);
            al = fmin(1.0, fmax(0.0, al));
            const double th = (1.0 - al) * t0 + al * t1;
            for (int d = 0; d < dim; ++d) xh[d] = x0[d] + al * (x1[d] - x0[d]);
            if (!sink_push(&s, th, xh)) break;
        }
    }
#undef G
    return s.n < cap ? s.n : cap;
}

/* ---- the CUBIC branch (interp_kind == "cubic") -------------------------------------------------------------------
 * Reference: backend.py _detect_with_segment_refine :541-553 (slopes of g), :584-631 (Hermite sub-interval values,
 * Newton on the cubic clamped to the sub-interval, cubic Hermite hit state), _refine_hits_cubic :274-379
 * (segment_refine == 0: Newton clamped to [0, 1]), utils.py _hermite_scalar :54-98 / _hermite_der :101-148.
 * _hermite_* are Numba functions: `x ** 2` with a literal exponent is x * x there; the hit-state weights are plain
 * Python floats, where `x ** 2` is libm pow(x, 2.0) -- kept apart on purpose.  Every cubic formula is guarded by
 * `dt > 0.0` in the reference: a trajectory with decreasing times (a backward tube) silently gets the linear ones. */
static double sq_mul(double x) { return x * x; }
static double hermite_scalar(double s, double y0, double y1, double dy0, double dy1, double dt)
{
    const double h00 = (1.0 + 2.0 * s) * sq_mul(1.0 - s);
    const double h10 = s * sq_mul(1.0 - s);
    const double h01 = sq_mul(s) * (3.0 - 2.0 * s);
    const double h11 = sq_mul(s) * (s - 1.0);
    return h00 * y0 + h10 * dy0 * dt + h01 * y1 + h11 * dy1 * dt;
}
static double hermite_der(double s, double y0, double y1, double dy0, double dy1, double dt)
{
    const double dh00 = 6.0 * s * (s - 1.0) + sq_mul(1.0 - s) * 2.0 - 2.0 * (1.0 - s) * (1.0 + 2.0 * s);
    const double dh10 = sq_mul(1.0 - s) + s * (2.0 * (s - 1.0));
    const double dh01 = 6.0 * s * (1.0 - s) - 2.0 * s * (3.0 - 2.0 * s);
    const double dh11 = 2.0 * s * (s - 1.0) + sq_mul(s);
    return dh00 * y0 + dh10 * dy0 * dt + dh01 * y1 + dh11 * dy1 * dt;
}
/* CPython float ** 2 (floatobject.c float_pow: shortcuts for a base of 1 and 0, otherwise libm pow).  glibc's pow(x, 2.0)
 * differs from x * x in ~0.09 % of arguments; the call goes through a volatile pointer because gcc folds a literal
 * pow(x, 2.0) into x * x. */
static double (*volatile libm_pow)(double, double) = pow;
static double py_sq(double x)
{
    if (x == 1.0) return 1.0;
    if (x == 0.0) return 0.0;
    return libm_pow(x, 2.0);
}
/* hit state at s on segment k: cubic Hermite through the four neighbouring samples when they exist (backend.py:628-645) */
static void cubic_state(const double *times, const double *states, int m, int dim, int k, double s, double dt,
                        int use_cubic, double *xh)
{
    const double *x0 = states + (size_t)k * dim, *x1 = states + (size_t)(k + 1) * dim;
    if (use_cubic && dt > 0.0 && (k - 1) >= 0 && (k + 2) < m) {
        const double *xm = states + (size_t)(k - 1) * dim, *xp = states + (size_t)(k + 2) * dim;
        const double dtm = times[k + 1] - times[k - 1], dtp = times[k + 2] - times[k];
        const double h00 = (1.0 + 2.0 * s) * py_sq(1.0 - s);
        const double h10 = s * py_sq(1.0 - s);
        const double h01 = py_sq(s) * (3.0 - 2.0 * s);
        const double h11 = py_sq(s) * (s - 1.0);
        for (int d = 0; d < dim; ++d) {
            const double dx0 = (x1[d] - xm[d]) / dtm, dx1 = (xp[d] - x0[d]) / dtp;
            xh[d] = h00 * x0[d] + h10 * dx0 * dt + h01 * x1[d] + h11 * dx1 * dt;
        }
    } else {
        for (int d = 0; d < dim; ++d) xh[d] = x0[d] + s * (x1[d] - x0[d]);
    }
}

int ho_synodic_detect_cubic(const double *times, const double *states, int m, int dim, int idx, double offset,
                            int direction, int proj_i, int proj_j, int segment_refine, double tol_on_surface,
                            double dedup_time_tol, double dedup_point_tol, int max_hits, int newton_max_iter,
                            double *hit_times, double *hit_states, int cap)
{
    hit_sink s = { hit_times, hit_states, cap, 0, dim, proj_i, proj_j, max_hits, dedup_time_tol, dedup_point_tol, 0, 0, 0 };
    if (m < 2) return 0;
    const int r = segment_refine;
    double xh[64];
#define G(k) (states[(size_t)(k) * dim + idx] - offset)
    for (int k = 0; k < m - 1; ++k) {
        const double t0 = times[k], t1 = times[k + 1];
        const double dt = t1 - t0;
        const double gk = G(k), gk1 = G(k + 1);
        const double *x0 = states + (size_t)k * dim;
        const int cubic = dt > 0.0;
        double d0 = 0.0, d1 = 0.0;
        if (cubic) {
            d0 = (k - 1 >= 0) ? (G(k + 1) - G(k - 1)) / (times[k + 1] - times[k - 1]) : (G(k + 1) - G(k)) / (times[k + 1] - times[k]);
            d1 = (k + 2 < m) ? (G(k + 2) - G(k)) / (times[k + 2] - times[k]) : (G(k + 1) - G(k)) / (times[k + 1] - times[k]);
        }
        int accept_left = 0;
        if (fabs(gk) < tol_on_surface) {
            if (direction == 0) accept_left = 1;
            else if (direction == 1) accept_left = (gk1 >= 0.0) || ((k - 1 >= 0) && (G(k - 1) <= 0.0));
            else accept_left = (gk1 <= 0.0) || ((k - 1 >= 0) && (G(k - 1) >= 0.0));
        }
        if (r > 0) {
            if (accept_left && !sink_push(&s, t0, x0)) return s.n < cap ? s.n : cap;
            const double step = 1.0 / (double)(r + 1);
            for (int mm = 0; mm <= r; ++mm) {
                const double s_lo = (double)mm * step, s_hi = (double)(mm + 1) * step;
                if (s_hi > 1.0 + 1e-15) break;
                if (accept_left && mm == 0) continue;
                double g_lo, g_hi;
                if (cubic) {
                    g_lo = hermite_scalar(s_lo, gk, gk1, d0, d1, dt);
                    g_hi = hermite_scalar(s_hi, gk, gk1, d0, d1, dt);
                } else {
                    g_lo = (1.0 - s_lo) * gk + s_lo * gk1;
                    g_hi = (1.0 - s_hi) * gk + s_hi * gk1;
                }
                int crosses;
                if (direction == 0) crosses = (g_lo * g_hi <= 0.0) && (g_lo != g_hi);
                else if (direction == 1) crosses = (g_lo < 0.0) && (g_hi >= 0.0);
                else crosses = (g_lo > 0.0) && (g_hi <= 0.0);
                if (!crosses) continue;
                double s_star;
                if (g_lo == g_hi) s_star = 0.5 * (s_lo + s_hi);
                else {
                    double al = g_lo / (g_lo - g_hi);
                    al = fmin(1.0, fmax(0.0, al));
                    s_star = s_lo + al * (s_hi - s_lo);
                }
                if (cubic) {
                    for (int it = 0; it < newton_max_iter; ++it) {
                        const double f = hermite_scalar(s_star, gk, gk1, d0, d1, dt);
                        const double df = hermite_der(s_star, gk, gk1, d0, d1, dt);
                        if (df == 0.0) break;
                        s_star -= f / df;
                        if (s_star < s_lo) { s_star = s_lo; break; }
                        if (s_star > s_hi) { s_star = s_hi; break; }
                    }
                }
                const double th = (1.0 - s_star) * t0 + s_star * t1;
                cubic_state(times, states, m, dim, k, s_star, dt, 1, xh);
                if (!sink_push(&s, th, xh)) return s.n < cap ? s.n : cap;
            }
        } else {
            if (accept_left) {
                if (!sink_push(&s, t0, x0)) break;
                continue;
            }
            int crosses;
            if (direction == 0) crosses = (gk * gk1 <= 0.0) && (gk != gk1);
            else if (direction == 1) crosses = (gk < 0.0) && (gk1 >= 0.0);
            else crosses = (gk > 0.0) && (gk1 <= 0.0);
            if (!crosses) continue;
            double al = gk / (gk - gk1);
            al = fmin(1.0, fmax(0.0, al));
            double s_star = al;
            if (cubic) {
                for (int it = 0; it < newton_max_iter; ++it) {
                    const double f = hermite_scalar(s_star, gk, gk1, d0, d1, dt);
                    const double df = hermite_der(s_star, gk, gk1, d0, d1, dt);
                    if (df == 0.0) break;
                    s_star -= f / df;
                    if (s_star < 0.0) { s_star = 0.0; break; }
                    if (s_star > 1.0) { s_star = 1.0; break; }
                }
            }
            const double th = (1.0 - s_star) * t0 + s_star * t1;
            cubic_state(times, states, m, dim, k, s_star, dt, 1, xh);
            if (!sink_push(&s, th, xh)) break;
        }
    }
#undef G
    return s.n < cap ? s.n : cap;
}

/* Batch form for bench.py's CPU legs: hit count over n uniformly sampled trajectories (pthreads). */
void ho_parallel_for(int64_t n, int n_threads, int64_t chunk, void (*fn)(int64_t, void *), void *ctx);

typedef struct {
    const double *times, *states; int m, dim, idx; double offset; int direction, pi, pj, refine;
    double tol, dtt, dpt; int64_t *counts;
} syn_ctx;

static void syn_item(int64_t i, void *p)
{
    syn_ctx *c = (syn_ctx *)p;
    double ht[64], hs[64 * 6];
    c->counts[i] = ho_synodic_detect(c->times, c->states + (size_t)i * c->m * c->dim, c->m, c->dim, c->idx, c->offset,
                                     c->direction, c->pi, c->pj, c->refine, c->tol, c->dtt, c->dpt, 0, ht, hs, 64);
}

int64_t ho_batch_synodic_count(const double *times, const double *states, int64_t n, int m, int dim, int idx,
                               double offset, int direction, int proj_i, int proj_j, int segment_refine,
                               double tol_on_surface, double dedup_time_tol, double dedup_point_tol, int n_threads)
{
    int64_t *counts = (int64_t *)__builtin_malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    syn_ctx c = { times, states, m, dim, idx, offset, direction, proj_i, proj_j, segment_refine, tol_on_surface,
                  dedup_time_tol, dedup_point_tol, counts };
    ho_parallel_for(n, n_threads, 8, syn_item, &c);
    int64_t tot = 0;
    for (int64_t i = 0; i < n; ++i) tot += counts[i];
    __builtin_free(counts);
    return tot;
}
