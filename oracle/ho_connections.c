/* ho_connections.c -- CPU restatement of the connection search (TEST INFRASTRUCTURE).
 * Reference: hiten/algorithms/connections/backends.py
 *   _pair_counts / _radpair2d / _radius_pairs_2d      :31-171   all (i, j) with |pu_i - ps_j|^2 <= eps^2, i-major
 *   mutual-nearest filter (Python dicts)               :468-489  first strict minimum in pair order on both sides
 *   _nearest_neighbor_2d(_numba)                       :174-233  first strict minimum over j != i
 *   _closest_points_on_segments_2d                     :237-320
 *   _refine_pairs_on_section                           :323-423
 *   Delta-V, classification, stable sort by delta_v    :507-533  (np.linalg.norm of a 3-vector = sqrt of an
 *                                                                FMA-accumulated dot product, OpenBLAS ddot)
 * Brute force like the reference (O(N*M)); the GPU path bins the points instead.                     */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static double d2(const double *a, const double *b)
{
    const double dx = a[0] - b[0], dy = a[1] - b[1];
    return dx * dx + dy * dy;
}

static int64_t nearest(const double *p, int64_t n, int64_t i)          /* :174-211 */
{
    double best = 1e300;
    int64_t bj = -1;
    for (int64_t j = 0; j < n; ++j) {
        if (j == i) continue;
        const double v = d2(p + 2 * i, p + 2 * j);
        if (v < best) { best = v; bj = j; }
    }
    return bj;
}

static void closest_on_segments(double a0x, double a0y, double a1x, double a1y, double b0x, double b0y, double b1x,
                                double b1y, double *s_out, double *t_out, double *px, double *py, double *qx, double *qy)
{
    const double ux = a1x - a0x, uy = a1y - a0y, vx = b1x - b0x, vy = b1y - b0y, wx = a0x - b0x, wy = a0y - b0y;
    const double A = ux * ux + uy * uy, B = ux * vx + uy * vy, C = vx * vx + vy * vy, D = ux * wx + uy * wy,
                 E = vx * wx + vy * wy;
    const double den = A * C - B * B;
    double s = 0.0, t = 0.0;
    if (den > 0.0) { s = (B * E - C * D) / den; t = (A * E - B * D) / den; }
    if (s < 0.0) { s = 0.0; if (C > 0.0) t = E / C; }
    else if (s > 1.0) { s = 1.0; if (C > 0.0) t = (E + B) / C; }
    if (t < 0.0) {
        t = 0.0;
        if (A > 0.0) { s = -D / A; if (s < 0.0) s = 0.0; else if (s > 1.0) s = 1.0; }
    } else if (t > 1.0) {
        t = 1.0;
        if (A > 0.0) { s = (B - D) / A; if (s < 0.0) s = 0.0; else if (s > 1.0) s = 1.0; }
    }
    *s_out = s; *t_out = t;
    *px = a0x + s * ux; *py = a0y + s * uy; *qx = b0x + t * vx; *qy = b0y + t * vy;
}

static double norm3(const double *v) { return sqrt(fma(v[2], v[2], fma(v[1], v[1], fma(v[0], v[0], 0.0)))); }

typedef struct { double dv; int64_t k; } sort_key;
static int cmp_key(const void *a, const void *b)
{
    const sort_key *x = a, *y = b;
    if (x->dv < y->dv) return -1;
    if (x->dv > y->dv) return 1;
    return (x->k > y->k) - (x->k < y->k);                        /* stable: ties keep pair order */
}

/* returns the number of accepted connections (<= cap written), sorted by delta_v; -1 on allocation failure */
int64_t ho_connections(const double *pu, int64_t n, const double *ps, int64_t m, const double *Xu, const double *Xs,
                       double eps, double dv_tol, double bal_tol, int64_t cap, int64_t *kind, double *dv_out,
                       double *pt, double *su, double *ss, int64_t *iu, int64_t *is, int64_t *pairs_considered)
{
    *pairs_considered = 0;
    if (n == 0 || m == 0) return 0;
    const double r2 = eps * eps;
    double *bi_v = malloc(sizeof(double) * n), *bj_v = malloc(sizeof(double) * m);
    int64_t *bi = malloc(sizeof(int64_t) * n), *bj = malloc(sizeof(int64_t) * m);
    if (!bi_v || !bj_v || !bi || !bj) return -1;
    for (int64_t i = 0; i < n; ++i) bi[i] = -1;
    for (int64_t j = 0; j < m; ++j) bj[j] = -1;
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < m; ++j) {
            const double v = d2(pu + 2 * i, ps + 2 * j);
            if (v <= r2) {
                ++total;
                if (bi[i] < 0 || v < bi_v[i]) { bi_v[i] = v; bi[i] = j; }
                if (bj[j] < 0 || v < bj_v[j]) { bj_v[j] = v; bj[j] = i; }
            }
        }
    *pairs_considered = total;
    int64_t n_res = 0, n_pairs = 0;
    sort_key *keys = malloc(sizeof(sort_key) * (n > 0 ? n : 1));
    double *rec = malloc(sizeof(double) * 16 * (n > 0 ? n : 1));       /* dv, kind, pt[2], su[6], ss[6] */
    int64_t *ridx = malloc(sizeof(int64_t) * 2 * (n > 0 ? n : 1));
    if (!keys || !rec || !ridx) return -1;
    for (int64_t i = 0; i < n; ++i) {
        if (bi[i] < 0) continue;
        const int64_t j = bi[i];
        if (!(bj[j] == i && bi_v[i] == bj_v[j])) continue;
        ++n_pairs;
        const int64_t iun = n >= 2 ? nearest(pu, n, i) : -1, jsn = m >= 2 ? nearest(ps, m, j) : -1;
        double xu[6], xs[6], p2[2], dv;
        int refined = 0;
        if (!(iun < 0 || jsn < 0 || iun == i || jsn == j)) {
            const double du = hypot(pu[2 * iun] - pu[2 * i], pu[2 * iun + 1] - pu[2 * i + 1]);
            const double ds = hypot(ps[2 * jsn] - ps[2 * j], ps[2 * jsn + 1] - ps[2 * j + 1]);
            if (!(du > 1e9 || ds > 1e9)) {
                double s, t, px, py, qx, qy;
                closest_on_segments(pu[2 * i], pu[2 * i + 1], pu[2 * iun], pu[2 * iun + 1], ps[2 * j], ps[2 * j + 1],
                                    ps[2 * jsn], ps[2 * jsn + 1], &s, &t, &px, &py, &qx, &qy);
                refined = 1;                                           /* valid and u0 != u1 and s0 != s1 */
                for (int c = 0; c < 6; ++c) {
                    xu[c] = (1.0 - s) * Xu[6 * i + c] + s * Xu[6 * iun + c];
                    xs[c] = (1.0 - t) * Xs[6 * j + c] + t * Xs[6 * jsn + c];
                }
                p2[0] = 0.5 * (px + qx); p2[1] = 0.5 * (py + qy);
            }
        }
        if (!refined) {
            memcpy(xu, Xu + 6 * i, sizeof xu); memcpy(xs, Xs + 6 * j, sizeof xs);
            p2[0] = pu[2 * i]; p2[1] = pu[2 * i + 1];
        }
        const double dvv[3] = {xu[3] - xs[3], xu[4] - xs[4], xu[5] - xs[5]};
        dv = norm3(dvv);
        if (dv <= dv_tol) {
            double *r = rec + 16 * n_res;
            r[0] = dv; r[1] = dv <= bal_tol ? 0.0 : 1.0; r[2] = p2[0]; r[3] = p2[1];
            memcpy(r + 4, xu, sizeof xu); memcpy(r + 10, xs, sizeof xs);
            ridx[2 * n_res] = i; ridx[2 * n_res + 1] = j;
            keys[n_res].dv = dv; keys[n_res].k = n_res;
            ++n_res;
        }
    }
    qsort(keys, (size_t)n_res, sizeof(sort_key), cmp_key);
    for (int64_t q = 0; q < n_res && q < cap; ++q) {
        const int64_t k = keys[q].k;
        const double *r = rec + 16 * k;
        dv_out[q] = r[0]; kind[q] = (int64_t)r[1]; pt[2 * q] = r[2]; pt[2 * q + 1] = r[3];
        memcpy(su + 6 * q, r + 4, 6 * sizeof(double)); memcpy(ss + 6 * q, r + 10, 6 * sizeof(double));
        iu[q] = ridx[2 * k]; is[q] = ridx[2 * k + 1];
    }
    free(bi_v); free(bj_v); free(bi); free(bj); free(keys); free(rec); free(ridx);
    (void)n_pairs;
    return n_res;
}
