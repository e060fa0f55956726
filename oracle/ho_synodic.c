/* ho_synodic.c -- CPU restatement of the synodic-section detector (TEST INFRASTRUCTURE).
 * Reference: hiten/algorithms/poincare/synodic/backend.py
 *   detect_on_trajectory :687-821, _detect_with_segment_refine :458-659 (linear branch),
 *   _on_surface_indices :134-183, _crossing_indices_and_alpha :184-233, _refine_hits_linear :234-273,
 *   _order_and_dedup_hits :382-455.
 * The shipped defaults always select the LINEAR branch; the cubic branch is
 * not restated.  direction: 0 = None, +1, -1.  Returns the number of hits written (<= cap).
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
