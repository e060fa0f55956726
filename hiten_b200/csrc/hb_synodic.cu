// hb_synodic.cu -- synodic-section crossing detection on precomputed (dense) trajectories.
//
// Replaces _SynodicDetectionBackend.run / detect_on_trajectory (hiten/algorithms/poincare/synodic/
// backend.py:687-887) and its pure-Python double loop _detect_with_segment_refine (:458-659), linear
// branch -- the one the shipped defaults always select -- plus the
// segment_refine == 0 path (:782-821) and _order_and_dedup_hits (:382-455).
//
// Layout: one warp per trajectory streams its [m][6] row-major samples (the reference's `states`
// array, 48 B per sample -> a warp reads 1536 contiguous bytes per iteration, fully coalesced);
// lane L owns segment k = base + L.  HBM-bound: 48 B per sample read once, hits are rare.
// Segments whose end values have the same strict sign cannot produce a hit under any direction rule
// (every sub-interval value is a convex combination of them), so only sign-changing / on-surface
// segments take the serial path, which every lane executes redundantly so that the de-duplication
// state (last kept hit) stays warp-uniform.  All arithmetic is separately rounded (__d*_rn) in the
// reference's operation order.
#include "hb_common.cuh"

namespace {

struct SynParams {
    const double *states;      // concatenated samples, [sum m_i][6]
    const double *times;       // signed times: concatenated per trajectory, or one shared array of m
    const long long *offsets;  // [n+1] sample offsets, or nullptr when every trajectory has m samples
    int m_uniform;
    int times_shared;
    long long n;
    hb_section sec;
    hb_hit *hits;
    long long capacity;
    int *hits_per_traj;        // optional [n]
    HbWorkspace *ws;
};

struct Dedup {
    double last_t, last_u, last_v;
    int n;
};

HB_DEV double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// _order_and_dedup_hits (backend.py:440-454); returns false when the per-trajectory cap is reached
HB_DEV bool push_hit(const SynParams &p, Dedup &dd, long long traj, double th, const double (&xh)[6], int lane)
{
    const int mh = p.sec.max_hits_per_traj;
    if (mh > 0 && dd.n >= mh) return false;
    const double u = (p.sec.proj_i == 0) ? xh[0] : (p.sec.proj_i == 1) ? xh[1] : (p.sec.proj_i == 2) ? xh[2]
                   : (p.sec.proj_i == 3) ? xh[3] : (p.sec.proj_i == 4) ? xh[4] : xh[5];
    const double v = (p.sec.proj_j == 0) ? xh[0] : (p.sec.proj_j == 1) ? xh[1] : (p.sec.proj_j == 2) ? xh[2]
                   : (p.sec.proj_j == 3) ? xh[3] : (p.sec.proj_j == 4) ? xh[4] : xh[5];
    if (dd.n > 0) {
        if (fabs(__dsub_rn(th, dd.last_t)) <= p.sec.dedup_time_tol) return true;
        const double du = __dsub_rn(u, dd.last_u), dv = __dsub_rn(v, dd.last_v);
        const double d2 = __dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv));
        if (d2 <= __dmul_rn(p.sec.dedup_point_tol, p.sec.dedup_point_tol)) return true;
    }
    if (lane == 0) {
        const unsigned long long slot = atomicAdd(&p.ws->hit_count, 1ULL);
        if ((long long)slot < p.capacity) {
            hb_hit *h = p.hits + slot;
            h->traj = traj; h->seq = dd.n; h->t = th;
#pragma unroll
            for (int d = 0; d < 6; ++d) h->state[d] = xh[d];
        } else {
            atomicAdd(&p.ws->overflow, 1ULL);
        }
    }
    dd.last_t = th; dd.last_u = u; dd.last_v = v;
    dd.n++;
    return true;
}

HB_DEV double pick(const double (&x)[6], int i)
{
    double r = x[0];
#pragma unroll
    for (int d = 1; d < 6; ++d) r = (i == d) ? x[d] : r;
    return r;
}

// Full per-segment logic of _detect_with_segment_refine / the r == 0 path, executed warp-uniformly.
HB_DEV bool process_segment(const SynParams &p, Dedup &dd, long long traj, int lane, bool has_prev, double g_prev,
                            double t0, double t1, const double (&x0)[6], const double (&x1)[6])
{
    const int dir = p.sec.direction;
    const double gk = __dsub_rn(pick(x0, p.sec.idx), p.sec.offset);
    const double gk1 = __dsub_rn(pick(x1, p.sec.idx), p.sec.offset);
    bool accept_left = false;
    if (fabs(gk) < p.sec.tol_on_surface) {
        if (dir == 0) accept_left = true;
        else if (dir > 0) accept_left = (gk1 >= 0.0) || (has_prev && g_prev <= 0.0);
        else accept_left = (gk1 <= 0.0) || (has_prev && g_prev >= 0.0);
    }
    const int r = p.sec.segment_refine;
    double xh[6];
    if (r > 0) {
        if (accept_left && !push_hit(p, dd, traj, t0, x0, lane)) return false;
        const double step = __ddiv_rn(1.0, (double)(r + 1));
        for (int mm = 0; mm <= r; ++mm) {
            const double s_lo = __dmul_rn((double)mm, step), s_hi = __dmul_rn((double)(mm + 1), step);
            if (s_hi > 1.0 + 1e-15) break;
            if (accept_left && mm == 0) continue;
            const double g_lo = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_lo), gk), __dmul_rn(s_lo, gk1));
            const double g_hi = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_hi), gk), __dmul_rn(s_hi, gk1));
            bool crosses;
            if (dir == 0) crosses = (__dmul_rn(g_lo, g_hi) <= 0.0) && (g_lo != g_hi);
            else if (dir > 0) crosses = (g_lo < 0.0) && (g_hi >= 0.0);
            else crosses = (g_lo > 0.0) && (g_hi <= 0.0);
            if (!crosses) continue;
            double s_star;
            if (g_lo == g_hi) s_star = __dmul_rn(0.5, __dadd_rn(s_lo, s_hi));
            else {
                double al = __ddiv_rn(g_lo, __dsub_rn(g_lo, g_hi));
                al = fmin(1.0, fmax(0.0, al));
                s_star = __dadd_rn(s_lo, __dmul_rn(al, __dsub_rn(s_hi, s_lo)));
            }
            const double th = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_star), t0), __dmul_rn(s_star, t1));
#pragma unroll
            for (int d = 0; d < 6; ++d) xh[d] = __dadd_rn(x0[d], __dmul_rn(s_star, __dsub_rn(x1[d], x0[d])));
            if (!push_hit(p, dd, traj, th, xh, lane)) return false;
        }
    } else {
        if (accept_left) return push_hit(p, dd, traj, t0, x0, lane);
        bool crosses;
        if (dir == 0) crosses = (__dmul_rn(gk, gk1) <= 0.0) && (gk != gk1);
        else if (dir > 0) crosses = (gk < 0.0) && (gk1 >= 0.0);
        else crosses = (gk > 0.0) && (gk1 <= 0.0);
        if (!crosses) return true;
        double al = __ddiv_rn(gk, __dsub_rn(gk, gk1));
        al = fmin(1.0, fmax(0.0, al));
        const double th = __dadd_rn(__dmul_rn(__dsub_rn(1.0, al), t0), __dmul_rn(al, t1));
#pragma unroll
        for (int d = 0; d < 6; ++d) xh[d] = __dadd_rn(x0[d], __dmul_rn(al, __dsub_rn(x1[d], x0[d])));
        return push_hit(p, dd, traj, th, xh, lane);
    }
    return true;
}

__global__ void __launch_bounds__(256) k_synodic_detect(const SynParams p)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long traj = warp0; traj < p.n; traj += nwarps) {
        const long long off = p.offsets ? p.offsets[traj] : traj * (long long)p.m_uniform;
        const int m = p.offsets ? (int)(p.offsets[traj + 1] - off) : p.m_uniform;
        const double *X = p.states + off * 6;
        const double *T = p.times_shared ? p.times : p.times + off;
        Dedup dd{0.0, 0.0, 0.0, 0};
        bool alive = true;
        for (int base = 0; alive && base < m - 1; base += 32) {
            const int k = base + lane;
            const bool valid = k < m - 1;
            double x0[6], x1[6];
            const int kk = valid ? k : m - 2;
#pragma unroll
            for (int d = 0; d < 6; ++d) x0[d] = X[(long long)kk * 6 + d];
            // x_{k+1}: from the next lane, the last lane (and the tail) load it
#pragma unroll
            for (int d = 0; d < 6; ++d) {
                const double nb = __shfl_down_sync(0xffffffffu, x0[d], 1);
                x1[d] = (lane == 31 || k + 1 >= m - 1) ? X[(long long)(kk + 1) * 6 + d] : nb;
            }
            const double gk = __dsub_rn(pick(x0, p.sec.idx), p.sec.offset);
            const double gk1 = __dsub_rn(pick(x1, p.sec.idx), p.sec.offset);
            double g_prev = __shfl_up_sync(0xffffffffu, gk, 1);
            if (lane == 0 && k > 0) g_prev = __dsub_rn(X[(long long)(k - 1) * 6 + p.sec.idx], p.sec.offset);
            const bool same_sign = (gk > 0.0 && gk1 > 0.0) || (gk < 0.0 && gk1 < 0.0);
            const bool flagged = valid && (!same_sign || fabs(gk) < p.sec.tol_on_surface);
            unsigned mask = __ballot_sync(0xffffffffu, flagged);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                double a0[6], a1[6];
#pragma unroll
                for (int d = 0; d < 6; ++d) { a0[d] = shfl_d(x0[d], src); a1[d] = shfl_d(x1[d], src); }
                const double gp = shfl_d(g_prev, src);
                const int ks = base + src;
                const double t0 = T[ks], t1 = T[ks + 1];
                if (!process_segment(p, dd, traj, lane, ks > 0, gp, t0, t1, a0, a1)) { alive = false; break; }
            }
        }
        if (p.hits_per_traj && lane == 0) p.hits_per_traj[traj] = dd.n;
    }
}

}  // namespace

extern "C" int hb_synodic_detect(const hb_section *sec, int64_t n_traj, const double *states, const double *times,
                                 const int64_t *offsets, int32_t m_uniform, int32_t times_shared, hb_hit *hits,
                                 int64_t hit_capacity, int32_t *hits_per_traj, void *workspace, void *stream)
{
    if (!sec || n_traj < 0 || !workspace) return HB_ERR_BADARG;
    if (sec->idx < 0 || sec->idx > 5 || sec->proj_i < 0 || sec->proj_i > 5 || sec->proj_j < 0 || sec->proj_j > 5)
        return HB_ERR_BADARG;
    if (sec->segment_refine < 0 || hit_capacity < 0) return HB_ERR_BADARG;
    if (n_traj > 0 && (!states || !times || (!offsets && m_uniform < 0) || (hit_capacity > 0 && !hits)))
        return HB_ERR_BADARG;
    if (times_shared && offsets) return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    if (n_traj == 0) return HB_OK;
    SynParams p{};
    p.states = states; p.times = times; p.offsets = (const long long *)offsets; p.m_uniform = m_uniform;
    p.times_shared = times_shared; p.n = n_traj; p.sec = *sec; p.hits = hits; p.capacity = hit_capacity;
    p.hits_per_traj = hits_per_traj; p.ws = (HbWorkspace *)workspace;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256;
    long long blocks = (n_traj * 32 + threads - 1) / threads;
    const long long cap = (long long)sms * 8;            // 8 CTAs x 8 warps resident per SM
    if (blocks > cap) blocks = cap;
    k_synodic_detect<<<(unsigned)blocks, threads, 0, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

// Reads the hit counter / overflow counter of a finished call (device -> host, synchronises the stream).
extern "C" int hb_read_hit_count(const void *workspace, int64_t *n_hits, int64_t *n_overflow, void *stream)
{
    if (!workspace) return HB_ERR_BADARG;
    HbWorkspace h;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemcpyAsync(&h, workspace, sizeof(h), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    if (n_hits) *n_hits = (int64_t)h.hit_count;
    if (n_overflow) *n_overflow = (int64_t)h.overflow;
    return HB_OK;
}
