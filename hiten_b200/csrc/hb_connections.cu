// hb_connections.cu -- connection search between two sets of section hits.
//
// Reference: hiten/algorithms/connections/backends.py
//   _radius_pairs_2d (:31-171)        all (i, j) with |pu_i - ps_j|^2 <= eps^2          -- O(N*M) double loop
//   mutual-nearest filter (:468-489)   Python dicts: first strict minimum in pair order on both sides
//   _nearest_neighbor_2d (:174-233)    nearest other point of the same set              -- O(N^2) double loop
//   _refine_pairs_on_section (:323-423) + _closest_points_on_segments_2d (:237-320)
//   Delta-V, classification (:507-533)
// The hit buffers of a config-5 run hold millions of points, where the O(N*M) loops are out of reach.  Here both
// sets are binned into a hashed uniform grid of cell size eps (counting sort: histogram, exclusive scan, scatter),
// every point finds its best partner in the other set by scanning the 3 x 3 neighbouring cells, the mutual pairs
// are compacted, and only THEIR nearest same-set neighbours are searched (ring search on the own grid; sparse cases fall
// back to an exact scan by one CTA per pair side; both with the reference's first-minimum tie rule).  Distances are evaluated with the reference's expression and
// rounding, minima with explicit (value, index) ordering, so indices, Delta-V and refined points are bit-identical
// whatever the order in which the grid delivers the candidates.
#include "hb_common.cuh"

#include <cub/cub.cuh>

namespace {

constexpr int CB = 256;

struct Grid {
    int *count;        // [B + 1]  bucket sizes, then exclusive offsets (in place)
    int *cursor;       // [B]
    int *order;        // [n]      point indices grouped by bucket
    int mask;          // B - 1
};

struct ConnParams {
    const double *pu, *ps, *Xu, *Xs;
    long long nu, ns;
    double cell, r2, dv_tol, bal_tol;
    Grid gu, gs;
    int *best_u_j, *best_s_i;        // best partner of every point (or -1)
    double *best_u_v, *best_s_v;     // its squared distance
    int *pairs;                      // [min(nu, ns)][2] mutual pairs
    int *nn;                         // [min(nu, ns)][2] nearest same-set neighbour of each pair member
    unsigned long long *counters;    // [0] pairs considered, [1] mutual pairs, [2] results appended, [3] results dropped
    hb_connection *out;
    long long capacity;
};

HB_DEV long long cell_of(double x, double cell) { return (long long)floor(x / cell); }
HB_DEV int bucket_of(long long cx, long long cy, int mask)
{
    const unsigned long long h = (unsigned long long)cx * 73856093ULL ^ (unsigned long long)cy * 19349663ULL;
    return (int)((h ^ (h >> 17)) & (unsigned long long)mask);
}
// the reference's expression: dx*dx + dy*dy with a = point of U, b = point of S, separately rounded
HB_DEV double dist2(double ax, double ay, double bx, double by)
{
    const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

__global__ void k_bucket_count(const double *p, long long n, double cell, Grid g)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicAdd(&g.count[bucket_of(cell_of(p[2 * i], cell), cell_of(p[2 * i + 1], cell), g.mask)], 1);
}

__global__ void k_bucket_fill(const double *p, long long n, double cell, Grid g)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = bucket_of(cell_of(p[2 * i], cell), cell_of(p[2 * i + 1], cell), g.mask);
    g.order[g.count[b] + atomicAdd(&g.cursor[b], 1)] = (int)i;
}

// Best partner (smallest distance, then smallest index = the reference's first strict minimum in pair order) of every
// query point among the reference set's points within eps.  QUERY_IS_U: queries are U points (also counts the pairs).
template <bool QUERY_IS_U>
__global__ void __launch_bounds__(CB) k_best_partner(const ConnParams p)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nq = QUERY_IS_U ? p.nu : p.ns;
    unsigned long long considered = 0;
    if (q < nq) {
        const double *Q = QUERY_IS_U ? p.pu : p.ps, *R = QUERY_IS_U ? p.ps : p.pu;
        const Grid &g = QUERY_IS_U ? p.gs : p.gu;
        const double qx = Q[2 * q], qy = Q[2 * q + 1];
        const long long cx = cell_of(qx, p.cell), cy = cell_of(qy, p.cell);
        int seen[9], n_seen = 0, best = -1;
        double best_v = 0.0;
        for (int oy = -1; oy <= 1; ++oy)
            for (int ox = -1; ox <= 1; ++ox) {
                const int b = bucket_of(cx + ox, cy + oy, g.mask);
                bool dup = false;
                for (int s = 0; s < n_seen; ++s) dup = dup || (seen[s] == b);      // two cells may share a bucket
                if (dup) continue;
                seen[n_seen++] = b;
                for (int k = g.count[b]; k < g.count[b + 1]; ++k) {
                    const int r = g.order[k];
                    const double v = QUERY_IS_U ? dist2(qx, qy, R[2 * r], R[2 * r + 1]) : dist2(R[2 * r], R[2 * r + 1], qx, qy);
                    if (v <= p.r2) {
                        ++considered;
                        if (best < 0 || v < best_v || (v == best_v && r < best)) { best_v = v; best = r; }
                    }
                }
            }
        if (QUERY_IS_U) { p.best_u_j[q] = best; p.best_u_v[q] = best_v; }
        else { p.best_s_i[q] = best; p.best_s_v[q] = best_v; }
    }
    if (QUERY_IS_U) {
        for (int o = 16; o > 0; o >>= 1) considered += __shfl_down_sync(0xffffffffu, considered, o);
        if ((threadIdx.x & 31) == 0 && considered) atomicAdd(&p.counters[0], considered);
    }
}

__global__ void k_mutual(const ConnParams p)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.nu) return;
    const int j = p.best_u_j[i];
    if (j < 0) return;
    if (p.best_s_i[j] == (int)i && p.best_u_v[i] == p.best_s_v[j]) {
        const unsigned long long k = atomicAdd(&p.counters[1], 1ULL);
        p.pairs[2 * k] = (int)i;
        p.pairs[2 * k + 1] = j;
    }
}

// Nearest same-set neighbour of each pair member, phase 1: one thread per (pair, side) searches its own grid in
// growing square rings.  After ring r every unvisited point is at least r cells away, so the search stops as soon as
// the best distance is strictly inside that radius (ties at the boundary cannot be missed).  Sparse neighbourhoods
// that are not settled within NN_MAX_RING rings are left to the exact scan below (nn = -2).
constexpr int NN_MAX_RING = 6;
__global__ void __launch_bounds__(CB) k_pair_neighbours_grid(const ConnParams p)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * (long long)p.counters[1]) return;
    const int side = (int)(t & 1);
    const double *P = side ? p.ps : p.pu;
    const long long n = side ? p.ns : p.nu;
    const Grid &g = side ? p.gs : p.gu;
    const int self = p.pairs[t];
    if (n < 2) { p.nn[t] = -1; return; }
    const double sx = P[2 * self], sy = P[2 * self + 1];
    const long long cx = cell_of(sx, p.cell), cy = cell_of(sy, p.cell);
    double best_v = 1e300;
    int best = -1;
    bool settled = false;
    for (int r = 0; r <= NN_MAX_RING && !settled; ++r) {
        for (int oy = -r; oy <= r; ++oy)
            for (int ox = -r; ox <= r; ++ox) {
                if (max(abs(ox), abs(oy)) != r) continue;                 // ring r only
                const int b = bucket_of(cx + ox, cy + oy, g.mask);
                for (int k = g.count[b]; k < g.count[b + 1]; ++k) {
                    const int j = g.order[k];
                    if (j == self) continue;
                    const double v = dist2(sx, sy, P[2 * j], P[2 * j + 1]);
                    if (v < best_v || (v == best_v && j < best)) { best_v = v; best = j; }
                }
            }
        const double reach = (double)r * p.cell;
        settled = best >= 0 && best_v < reach * reach * (1.0 - 1e-12);
    }
    p.nn[t] = settled ? best : -2;
}

// Phase 2: one CTA per (pair, side) still open, exact scan of the whole set.
__global__ void __launch_bounds__(CB) k_pair_neighbours(const ConnParams p)
{
    const long long pair = blockIdx.x >> 1;
    const int side = blockIdx.x & 1;
    if (p.nn[2 * pair + side] != -2) return;
    const double *P = side ? p.ps : p.pu;
    const long long n = side ? p.ns : p.nu;
    const int self = p.pairs[2 * pair + side];
    const double sx = P[2 * self], sy = P[2 * self + 1];
    double best_v = 1e300;
    int best = -1;
    for (long long j = threadIdx.x; j < n; j += CB) {
        if (j == self) continue;
        const double v = dist2(sx, sy, P[2 * j], P[2 * j + 1]);
        if (v < best_v) { best_v = v; best = (int)j; }               // ascending j per thread: first minimum kept
    }
    __shared__ double sv[CB];
    __shared__ int si[CB];
    sv[threadIdx.x] = best_v; si[threadIdx.x] = best;
    __syncthreads();
    for (int o = CB / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v = sv[threadIdx.x + o];
            const int i2 = si[threadIdx.x + o];
            const bool take = i2 >= 0 && (si[threadIdx.x] < 0 || v < sv[threadIdx.x] || (v == sv[threadIdx.x] && i2 < si[threadIdx.x]));
            if (take) { sv[threadIdx.x] = v; si[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) p.nn[2 * pair + side] = (n >= 2) ? si[0] : -1;
}

// _closest_points_on_segments_2d (backends.py:237-320)
HB_DEV void closest_on_segments(double a0x, double a0y, double a1x, double a1y, double b0x, double b0y, double b1x,
                                double b1y, double &s, double &t, double &px, double &py, double &qx, double &qy)
{
    const double ux = __dsub_rn(a1x, a0x), uy = __dsub_rn(a1y, a0y), vx = __dsub_rn(b1x, b0x), vy = __dsub_rn(b1y, b0y);
    const double wx = __dsub_rn(a0x, b0x), wy = __dsub_rn(a0y, b0y);
    const double A = __dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)), B = __dadd_rn(__dmul_rn(ux, vx), __dmul_rn(uy, vy));
    const double C = __dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), D = __dadd_rn(__dmul_rn(ux, wx), __dmul_rn(uy, wy));
    const double E = __dadd_rn(__dmul_rn(vx, wx), __dmul_rn(vy, wy));
    const double den = __dsub_rn(__dmul_rn(A, C), __dmul_rn(B, B));
    s = 0.0; t = 0.0;
    if (den > 0.0) {
        s = __ddiv_rn(__dsub_rn(__dmul_rn(B, E), __dmul_rn(C, D)), den);
        t = __ddiv_rn(__dsub_rn(__dmul_rn(A, E), __dmul_rn(B, D)), den);
    }
    if (s < 0.0) { s = 0.0; if (C > 0.0) t = __ddiv_rn(E, C); }
    else if (s > 1.0) { s = 1.0; if (C > 0.0) t = __ddiv_rn(__dadd_rn(E, B), C); }
    if (t < 0.0) {
        t = 0.0;
        if (A > 0.0) { s = __ddiv_rn(-D, A); if (s < 0.0) s = 0.0; else if (s > 1.0) s = 1.0; }
    } else if (t > 1.0) {
        t = 1.0;
        if (A > 0.0) { s = __ddiv_rn(__dsub_rn(B, D), A); if (s < 0.0) s = 0.0; else if (s > 1.0) s = 1.0; }
    }
    px = __dadd_rn(a0x, __dmul_rn(s, ux)); py = __dadd_rn(a0y, __dmul_rn(s, uy));
    qx = __dadd_rn(b0x, __dmul_rn(t, vx)); qy = __dadd_rn(b0y, __dmul_rn(t, vy));
}

// _refine_pairs_on_section + Delta-V (backends.py:323-423, 507-533), one thread per mutual pair
__global__ void k_refine(const ConnParams p)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)p.counters[1]) return;
    const int i = p.pairs[2 * k], j = p.pairs[2 * k + 1];
    const int iu = p.nn[2 * k], js = p.nn[2 * k + 1];
    double xu[6], xs[6], pt0, pt1;
    bool refined = false;
    if (!(iu < 0 || js < 0 || iu == i || js == j)) {
        const double du = hypot(p.pu[2 * iu] - p.pu[2 * i], p.pu[2 * iu + 1] - p.pu[2 * i + 1]);
        const double ds = hypot(p.ps[2 * js] - p.ps[2 * j], p.ps[2 * js + 1] - p.ps[2 * j + 1]);
        if (!(du > 1e9 || ds > 1e9)) {
            double s, t, px, py, qx, qy;
            closest_on_segments(p.pu[2 * i], p.pu[2 * i + 1], p.pu[2 * iu], p.pu[2 * iu + 1], p.ps[2 * j], p.ps[2 * j + 1],
                                p.ps[2 * js], p.ps[2 * js + 1], s, t, px, py, qx, qy);
            refined = true;
            const double oms = __dsub_rn(1.0, s), omt = __dsub_rn(1.0, t);
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                xu[c] = __dadd_rn(__dmul_rn(oms, p.Xu[6LL * i + c]), __dmul_rn(s, p.Xu[6LL * iu + c]));
                xs[c] = __dadd_rn(__dmul_rn(omt, p.Xs[6LL * j + c]), __dmul_rn(t, p.Xs[6LL * js + c]));
            }
            pt0 = __dmul_rn(0.5, __dadd_rn(px, qx));
            pt1 = __dmul_rn(0.5, __dadd_rn(py, qy));
        }
    }
    if (!refined) {
#pragma unroll
        for (int c = 0; c < 6; ++c) { xu[c] = p.Xu[6LL * i + c]; xs[c] = p.Xs[6LL * j + c]; }
        pt0 = p.pu[2 * i]; pt1 = p.pu[2 * i + 1];
    }
    // np.linalg.norm of a 3-vector = sqrt(x.dot(x)): FMA-accumulated dot product (OpenBLAS ddot)
    const double d0 = __dsub_rn(xu[3], xs[3]), d1 = __dsub_rn(xu[4], xs[4]), d2 = __dsub_rn(xu[5], xs[5]);
    const double dv = __dsqrt_rn(__fma_rn(d2, d2, __fma_rn(d1, d1, __fma_rn(d0, d0, 0.0))));
    if (!(dv <= p.dv_tol)) return;
    const unsigned long long slot = atomicAdd(&p.counters[2], 1ULL);
    if ((long long)slot >= p.capacity) { atomicAdd(&p.counters[3], 1ULL); return; }
    hb_connection &o = p.out[slot];
    o.index_u = i; o.index_s = j; o.delta_v = dv; o.point2d[0] = pt0; o.point2d[1] = pt1;
#pragma unroll
    for (int c = 0; c < 6; ++c) { o.state_u[c] = xu[c]; o.state_s[c] = xs[c]; }
    o.kind = (dv <= p.bal_tol) ? 0 : 1;
}

inline int buckets_for(long long n)
{
    long long b = 1024;
    while (b < 2 * n && b < (1LL << 28)) b <<= 1;
    return (int)b;
}
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Layout {
    size_t count_u, cursor_u, order_u, count_s, cursor_s, order_s, best_u_j, best_s_i, best_u_v, best_s_v, pairs, nn,
        counters, cub, total;
    int bu, bs;
    size_t cub_bytes;
};

Layout layout(long long nu, long long ns)
{
    Layout L{};
    L.bu = buckets_for(nu); L.bs = buckets_for(ns);
    const long long np = nu < ns ? nu : ns;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = align256(o + bytes); return at; };
    L.count_u = take(sizeof(int) * ((size_t)L.bu + 1)); L.cursor_u = take(sizeof(int) * (size_t)L.bu); L.order_u = take(sizeof(int) * (size_t)nu);
    L.count_s = take(sizeof(int) * ((size_t)L.bs + 1)); L.cursor_s = take(sizeof(int) * (size_t)L.bs); L.order_s = take(sizeof(int) * (size_t)ns);
    L.best_u_j = take(sizeof(int) * (size_t)nu); L.best_s_i = take(sizeof(int) * (size_t)ns);
    L.best_u_v = take(sizeof(double) * (size_t)nu); L.best_s_v = take(sizeof(double) * (size_t)ns);
    L.pairs = take(sizeof(int) * 2 * (size_t)np); L.nn = take(sizeof(int) * 2 * (size_t)np);
    L.counters = take(sizeof(unsigned long long) * 8);
    size_t cb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cb, (int *)nullptr, (int *)nullptr, (L.bu > L.bs ? L.bu : L.bs) + 1);
    L.cub_bytes = cb;
    L.cub = take(cb);
    L.total = o;
    return L;
}

int build_grid(const double *pts, long long n, double cell, Grid g, int buckets, void *cub_tmp, size_t cub_bytes,
               cudaStream_t st)
{
    HB_CUDA_TRY(cudaMemsetAsync(g.count, 0, sizeof(int) * ((size_t)buckets + 1), st));
    HB_CUDA_TRY(cudaMemsetAsync(g.cursor, 0, sizeof(int) * (size_t)buckets, st));
    const unsigned blocks = (unsigned)((n + CB - 1) / CB);
    k_bucket_count<<<blocks, CB, 0, st>>>(pts, n, cell, g);
    HB_CUDA_TRY(cudaGetLastError());
    HB_CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, g.count, g.count, buckets + 1, st));
    k_bucket_fill<<<blocks, CB, 0, st>>>(pts, n, cell, g);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

}  // namespace

extern "C" int64_t hb_connections_scratch_bytes(int64_t n_u, int64_t n_s)
{
    if (n_u < 0 || n_s < 0 || n_u > 2000000000LL || n_s > 2000000000LL) return -1;
    return (int64_t)layout(n_u, n_s).total;
}

extern "C" int hb_connections(const double *pu, int64_t n_u, const double *ps, int64_t n_s, const double *Xu,
                              const double *Xs, double eps, double dv_tol, double bal_tol, hb_connection *out,
                              int64_t capacity, int64_t *n_out, int64_t *n_dropped, int64_t *pairs_considered,
                              void *scratch, int64_t scratch_bytes, void *stream)
{
    if (n_u < 0 || n_s < 0 || capacity < 0 || !n_out || !(eps > 0.0)) return HB_ERR_BADARG;
    *n_out = 0;
    if (n_dropped) *n_dropped = 0;
    if (pairs_considered) *pairs_considered = 0;
    if (n_u == 0 || n_s == 0) return HB_OK;
    if (n_u > 2000000000LL || n_s > 2000000000LL) return HB_ERR_UNSUPPORTED;
    if (!pu || !ps || !Xu || !Xs || !scratch || (capacity > 0 && !out)) return HB_ERR_BADARG;
    const Layout L = layout(n_u, n_s);
    if ((size_t)scratch_bytes < L.total) return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    char *base = (char *)scratch;
    ConnParams p{};
    p.pu = pu; p.ps = ps; p.Xu = Xu; p.Xs = Xs; p.nu = n_u; p.ns = n_s;
    p.cell = eps * 1.000000001;                       // a hair above eps: neighbours within eps are <= 1 cell apart
    p.r2 = eps * eps; p.dv_tol = dv_tol; p.bal_tol = bal_tol;
    p.gu = Grid{(int *)(base + L.count_u), (int *)(base + L.cursor_u), (int *)(base + L.order_u), L.bu - 1};
    p.gs = Grid{(int *)(base + L.count_s), (int *)(base + L.cursor_s), (int *)(base + L.order_s), L.bs - 1};
    p.best_u_j = (int *)(base + L.best_u_j); p.best_s_i = (int *)(base + L.best_s_i);
    p.best_u_v = (double *)(base + L.best_u_v); p.best_s_v = (double *)(base + L.best_s_v);
    p.pairs = (int *)(base + L.pairs); p.nn = (int *)(base + L.nn);
    p.counters = (unsigned long long *)(base + L.counters);
    p.out = out; p.capacity = capacity;
    HB_CUDA_TRY(cudaMemsetAsync(p.counters, 0, sizeof(unsigned long long) * 8, st));
    int rc = build_grid(pu, n_u, p.cell, p.gu, L.bu, base + L.cub, L.cub_bytes, st);
    if (rc != HB_OK) return rc;
    rc = build_grid(ps, n_s, p.cell, p.gs, L.bs, base + L.cub, L.cub_bytes, st);
    if (rc != HB_OK) return rc;
    k_best_partner<true><<<(unsigned)((n_u + CB - 1) / CB), CB, 0, st>>>(p);
    k_best_partner<false><<<(unsigned)((n_s + CB - 1) / CB), CB, 0, st>>>(p);
    k_mutual<<<(unsigned)((n_u + CB - 1) / CB), CB, 0, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    unsigned long long h[4];
    HB_CUDA_TRY(cudaMemcpyAsync(h, p.counters, sizeof h, cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    if (pairs_considered) *pairs_considered = (int64_t)h[0];
    const long long n_pairs = (long long)h[1];
    if (n_pairs == 0) return HB_OK;
    if (2 * n_pairs > 2147483647LL) return HB_ERR_UNSUPPORTED;
    k_pair_neighbours_grid<<<(unsigned)((2 * n_pairs + CB - 1) / CB), CB, 0, st>>>(p);
    k_pair_neighbours<<<(unsigned)(2 * n_pairs), CB, 0, st>>>(p);
    k_refine<<<(unsigned)((n_pairs + CB - 1) / CB), CB, 0, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    HB_CUDA_TRY(cudaMemcpyAsync(h, p.counters, sizeof h, cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    *n_out = (int64_t)(h[2] < (unsigned long long)capacity ? h[2] : (unsigned long long)capacity);
    if (n_dropped) *n_dropped = (int64_t)h[3];
    return HB_OK;
}
