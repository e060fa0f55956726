// hb_synodic.cu -- synodic-section crossing detection on precomputed (dense) trajectories.
//
// Replaces _SynodicDetectionBackend.run / detect_on_trajectory (hiten/algorithms/poincare/synodic/
// backend.py:687-887) and its pure-Python double loop _detect_with_segment_refine (:458-659): the linear
// branch -- the one the shipped defaults always select -- in k_synodic_detect, the
// cubic branch (interp_kind == "cubic": Hermite g, Newton on the cubic, cubic hit state; _refine_hits_cubic
// :274-379) in k_synodic_detect_cubic, plus the segment_refine == 0 path (:782-821) and
// _order_and_dedup_hits (:382-455).
//
// Layout: one warp per trajectory streams its [m][6] row-major samples (the reference's `states`
// array, 48 B per sample -> a warp reads 1536 contiguous bytes per iteration, fully coalesced);
// lane L owns segment k = base + L.  HBM-bound: 48 B per sample read once, hits are rare.
// Segments whose end values have the same strict sign cannot produce a hit under any direction rule
// (every sub-interval value is a convex combination of them), so only sign-changing / on-surface
// segments take the serial path, which every lane executes redundantly so that the de-duplication
// state (last kept hit) stays warp-uniform.  All arithmetic is separately rounded (__d*_rn) in the
// reference's operation order.
#include "hb_section.cuh"

namespace {

struct SynParams {
    const double *states;      // concatenated samples, [sum m_i][6]
    const double *times;       // signed times: concatenated per trajectory, or one shared array of m
    const long long *offsets;  // [n+1] sample offsets, or nullptr when every trajectory has m samples
    int m_uniform;
    int times_shared;
    long long n;
    HitSink sink;
    int *hits_per_traj;        // optional [n]
};

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
            const double gk = __dsub_rn(pick(x0, p.sink.sec.idx), p.sink.sec.offset);
            const double gk1 = __dsub_rn(pick(x1, p.sink.sec.idx), p.sink.sec.offset);
            double g_prev = __shfl_up_sync(0xffffffffu, gk, 1);
            if (lane == 0 && k > 0) g_prev = __dsub_rn(X[(long long)(k - 1) * 6 + p.sink.sec.idx], p.sink.sec.offset);
            const bool same_sign = (gk > 0.0 && gk1 > 0.0) || (gk < 0.0 && gk1 < 0.0);
            const bool flagged = valid && (!same_sign || fabs(gk) < p.sink.sec.tol_on_surface);
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
                if (!process_segment(p.sink, dd, traj, lane, ks > 0, gp, t0, t1, a0, a1)) { alive = false; break; }
            }
        }
        if (p.hits_per_traj && lane == 0) p.hits_per_traj[traj] = dd.n;
    }
}

// The cubic request (interp_kind == "cubic").  With segment_refine > 0 the Hermite cubic of g can reach the plane inside a
// segment whose two samples lie on the same side, so the sample signs alone do not clear a segment: it is cleared when the
// four Bezier control values of the cubic (g_k, g_k + d0 dt / 3, g_k+1 - d1 dt / 3, g_k+1 -- the cubic lies in their hull)
// share a strict sign by more than 1e-12 of their scale, three orders above the rounding of the reference's 52 evaluations.
// Flagged segments take the serial warp-uniform path (process_segment_cubic), which reads its samples by index.
__global__ void __launch_bounds__(256) k_synodic_detect_cubic(const SynParams p, const int newton_max_iter)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int ci = p.sink.sec.idx, r = p.sink.sec.segment_refine;
    const double off = p.sink.sec.offset;
    for (long long traj = warp0; traj < p.n; traj += nwarps) {
        const long long o = p.offsets ? p.offsets[traj] : traj * (long long)p.m_uniform;
        const int m = p.offsets ? (int)(p.offsets[traj + 1] - o) : p.m_uniform;
        const double *X = p.states + o * 6;
        const double *T = p.times_shared ? p.times : p.times + o;
        Dedup dd{0.0, 0.0, 0.0, 0};
        bool alive = true;
        for (int base = 0; alive && base < m - 1; base += 32) {
            const int k = base + lane;
            bool flagged = false;
            if (k < m - 1) {
                const double gk = __dsub_rn(X[(long long)k * 6 + ci], off), gk1 = __dsub_rn(X[(long long)(k + 1) * 6 + ci], off);
                const bool same_sign = (gk > 0.0 && gk1 > 0.0) || (gk < 0.0 && gk1 < 0.0);
                flagged = !same_sign || fabs(gk) < p.sink.sec.tol_on_surface;
                const double dt = __dsub_rn(T[k + 1], T[k]);
                if (!flagged && r > 0 && dt > 0.0) {
                    const double gm = k > 0 ? __dsub_rn(X[(long long)(k - 1) * 6 + ci], off) : gk;
                    const double gp = k + 2 < m ? __dsub_rn(X[(long long)(k + 2) * 6 + ci], off) : gk1;
                    const double m0 = (k > 0 ? (gk1 - gm) / (T[k + 1] - T[k - 1]) : (gk1 - gk) / dt) * dt;
                    const double m1 = (k + 2 < m ? (gp - gk) / (T[k + 2] - T[k]) : (gk1 - gk) / dt) * dt;
                    const double b1 = gk + m0 / 3.0, b2 = gk1 - m1 / 3.0;
                    const double margin = 1e-12 * (fabs(gk) + fabs(gk1) + fabs(m0) + fabs(m1));
                    const double lo = fmin(fmin(gk, gk1), fmin(b1, b2)), hi = fmax(fmax(gk, gk1), fmax(b1, b2));
                    flagged = !(lo > margin || hi < -margin);
                }
            }
            unsigned mask = __ballot_sync(0xffffffffu, flagged);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                if (!process_segment_cubic(p.sink, dd, traj, lane, base + src, m, T, X, newton_max_iter)) { alive = false; break; }
            }
        }
        if (p.hits_per_traj && lane == 0) p.hits_per_traj[traj] = dd.n;
    }
}

int detect_launch(const hb_section *sec, int cubic, int32_t newton_max_iter, int64_t n_traj, const double *states,
                  const double *times, const int64_t *offsets, int32_t m_uniform, int32_t times_shared, hb_hit *hits,
                  int64_t hit_capacity, int32_t *hits_per_traj, void *workspace, void *stream)
{
    if (!sec || n_traj < 0 || !workspace) return HB_ERR_BADARG;
    if (sec->idx < 0 || sec->idx > 5 || sec->proj_i < 0 || sec->proj_i > 5 || sec->proj_j < 0 || sec->proj_j > 5)
        return HB_ERR_BADARG;
    if (sec->segment_refine < 0 || hit_capacity < 0 || newton_max_iter < 0) return HB_ERR_BADARG;
    if (n_traj > 0 && (!states || !times || (!offsets && m_uniform < 0) || (hit_capacity > 0 && !hits)))
        return HB_ERR_BADARG;
    if (times_shared && offsets) return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    if (n_traj == 0) return HB_OK;
    SynParams p{};
    p.states = states; p.times = times; p.offsets = (const long long *)offsets; p.m_uniform = m_uniform;
    p.times_shared = times_shared; p.n = n_traj;
    p.sink.sec = *sec; p.sink.hits = hits; p.sink.capacity = hit_capacity; p.sink.ws = (HbWorkspace *)workspace;
    p.hits_per_traj = hits_per_traj;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256;
    long long blocks = (n_traj * 32 + threads - 1) / threads;
    const long long cap = (long long)sms * 8;            // 8 CTAs x 8 warps resident per SM
    if (blocks > cap) blocks = cap;
    if (cubic) k_synodic_detect_cubic<<<(unsigned)blocks, threads, 0, st>>>(p, newton_max_iter);
    else k_synodic_detect<<<(unsigned)blocks, threads, 0, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

}  // namespace

extern "C" int hb_synodic_detect_cubic(const hb_section *sec, int32_t newton_max_iter, int64_t n_traj,
                                       const double *states, const double *times, const int64_t *offsets,
                                       int32_t m_uniform, int32_t times_shared, hb_hit *hits, int64_t hit_capacity,
                                       int32_t *hits_per_traj, void *workspace, void *stream)
{
    return detect_launch(sec, 1, newton_max_iter, n_traj, states, times, offsets, m_uniform, times_shared, hits,
                         hit_capacity, hits_per_traj, workspace, stream);
}

extern "C" int hb_synodic_detect(const hb_section *sec, int64_t n_traj, const double *states, const double *times,
                                 const int64_t *offsets, int32_t m_uniform, int32_t times_shared, hb_hit *hits,
                                 int64_t hit_capacity, int32_t *hits_per_traj, void *workspace, void *stream)
{
    return detect_launch(sec, 0, 0, n_traj, states, times, offsets, m_uniform, times_shared, hits, hit_capacity,
                         hits_per_traj, workspace, stream);
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

extern "C" int hb_read_record_overflow(const void *workspace, int64_t *n_traj, void *stream)
{
    if (!workspace || !n_traj) return HB_ERR_BADARG;
    HbWorkspace h;
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemcpyAsync(&h, workspace, sizeof(h), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    *n_traj = (int64_t)h.rec_overflow;
    return HB_OK;
}
