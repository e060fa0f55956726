// hb_manifold.cu -- the manifold-tube steps either side of the propagation loop, on the device.
//
// Replaces, for a whole tube at once (paths relative to hiten/):
//   _ManifoldDynamicsService._totime                    algorithms/types/services/manifold.py:539-573
//   _ManifoldDynamicsService._compute_manifold_section  algorithms/types/services/manifold.py:470-537
//   the safe-radius test of _run_compute                algorithms/types/services/manifold.py:412-424
//   _max_rel_energy_error                               algorithms/common/energy.py:27-76
//
// hb_manifold_ics   consumes the dense state+STM samples PHI[S][42] where hb_cr3bp_stm_dense left them and writes
//                   the tube's initial conditions straight into the SoA batch array the propagation kernels read:
//                   one warp per fraction finds the STM sample (first minimum of |f*T - |t_k||, explicit
//                   (value, index) order), one thread per (fraction, displacement) forms
//                   x0W = x_k + d * (direction * Phi_k @ eigvec).
// hb_tube_filter    one warp per trajectory streams its stored [m][6] samples (48 B per sample read once, coalesced
//                   1536-byte rows per warp iteration: HBM-bound) and reduces min r1, min r2 and the maximum relative
//                   drift of the Jacobi constant; optionally writes the keep / discard decision of _run_compute.
//
// Arithmetic: every operation separately rounded in the reference's order, IEEE division and square root, with
// the two platform details of the reference's numpy calls reproduced (see oracle/ho_manifold.c): the 6x6 @ complex
// vector product runs through OpenBLAS zgemv (two FMA lanes over elements 0..3, mul + add tail for 4..5) and the
// 3-element norm is an FMA-accumulated dot.  Results are bit-identical to the reference
// (tests/golden/manifold_ics.npz, tests/test_gpu_manifold.py).
#include "hb_tubefilter.cuh"

namespace {

// ---- _totime -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_totime(const double *__restrict__ tt, int S, double period,
                                                const double *__restrict__ fractions, long long K,
                                                int *__restrict__ node_idx)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long k = warp0; k < K; k += nwarps) {
        const double target = __dmul_rn(fractions[k], period);
        double bd = CUDART_INF;
        int bi = 0x7fffffff;
        for (int s = lane; s < S; s += 32) {
            const double d = fabs(__dsub_rn(target, fabs(tt[s])));
            if (d < bd) { bd = d; bi = s; }               // ascending s per lane: first minimum kept
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) node_idx[k] = bi == 0x7fffffff ? 0 : bi;   // all-NaN row: np.argmin returns 0
    }
}

// one row of `phi_frac @ eigvec` (real part) in OpenBLAS zgemv_t order
HB_DEV double zgemv_row6(const double *a, const double *x)
{
    double l0 = __fma_rn(a[0], x[0], 0.0);
    double l1 = __fma_rn(a[1], x[1], 0.0);
    l0 = __fma_rn(a[2], x[2], l0);
    l1 = __fma_rn(a[3], x[3], l1);
    const double head = __dadd_rn(l0, l1);
    double tail = __dadd_rn(0.0, __dmul_rn(a[4], x[4]));
    tail = __dadd_rn(tail, __dmul_rn(a[5], x[5]));
    return __dadd_rn(head, tail);
}

// ---- _compute_manifold_section -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_manifold_ics(const double *__restrict__ phi, const int *__restrict__ node_idx,
                                                      const double *__restrict__ eigvec, double direction,
                                                      const double *__restrict__ disp, long long K, long long D,
                                                      double *__restrict__ x0w_soa)
{
    const long long n = K * D;
    double ev[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) ev[c] = eigvec[c];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long k = i % K, j = i / K;              // displacement-major: row j*K + k
        const double *row = phi + 42ll * node_idx[k];
        double man[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double a[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) a[c] = row[6 * r + c];
            man[r] = __dmul_rn(direction, zgemv_row6(a, ev));
        }
        double sq = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) sq = __fma_rn(man[c], man[c], sq);
        sq = __dadd_rn(sq, 0.0);
        double mag = __dsqrt_rn(sq);
        if (mag < 1e-14) mag = 1.0;
        const double d = __ddiv_rn(disp[j], mag);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double v = __dadd_rn(row[36 + c], __dmul_rn(d, man[c]));
            if ((c == 2 || c == 5) && fabs(v) < 1.0e-15) v = 0.0;
            x0w_soa[(long long)c * n + i] = v;
        }
    }
}

// ---- tube filters (per-sample arithmetic in hb_tubefilter.cuh) ----------------------------------------------------
__global__ void __launch_bounds__(256) k_tube_filter(const double *__restrict__ states, long long n, int m,
                                                     hb_tube_filter_opts o, double *__restrict__ out,
                                                     int *__restrict__ keep)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const double mu1 = __dsub_rn(1.0, o.mu), mu2 = o.mu;
    for (long long traj = warp0; traj < n; traj += nwarps) {
        const double *X = states + traj * (long long)m * 6;
        double s[6];
#pragma unroll
        for (int d = 0; d < 6; ++d) s[d] = X[d];
        const double C0 = jacobi_ref(s, mu1, mu2), absC0 = fabs(C0);
        TubeFilterAcc acc;
        // two samples per lane and iteration, 16-byte loads (a sample is 48 B: 16-byte aligned whenever the array is):
        // six independent loads in flight per lane instead of one dependent group
        if ((reinterpret_cast<unsigned long long>(X) & 15ull) == 0) {
            const double2 *X2 = reinterpret_cast<const double2 *>(X);
            int k = lane;
            for (; k + 32 < m; k += 64) {
                const double2 a0 = X2[3ll * k], a1 = X2[3ll * k + 1], a2 = X2[3ll * k + 2];
                const double2 b0 = X2[3ll * (k + 32)], b1 = X2[3ll * (k + 32) + 1], b2 = X2[3ll * (k + 32) + 2];
                const double sa[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y}, sb[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
                acc.sample(sa, k, o.mu, mu1, mu2, C0);
                acc.sample(sb, k + 32, o.mu, mu1, mu2, C0);
            }
            if (k < m) {
                const double2 a0 = X2[3ll * k], a1 = X2[3ll * k + 1], a2 = X2[3ll * k + 2];
                const double sa[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
                acc.sample(sa, k, o.mu, mu1, mu2, C0);
            }
        } else {
            for (int k = lane; k < m; k += 32) {
#pragma unroll
                for (int d = 0; d < 6; ++d) s[d] = X[(long long)k * 6 + d];
                acc.sample(s, k, o.mu, mu1, mu2, C0);
            }
        }
        acc.warp_reduce();
        if (lane == 0) acc.store(o, traj, absC0, out, keep);
    }
}

int sm_count()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

extern "C" int hb_manifold_ics(const double *phi_dense, const double *tt, int32_t n_samples, double period,
                               const double *eigvec, int32_t direction, const double *fractions, int64_t n_fractions,
                               const double *displacements, int64_t n_displacements, double *x0w_soa,
                               int32_t *node_idx, void *stream)
{
    if (n_fractions < 0 || n_displacements < 0 || n_samples < 1) return HB_ERR_BADARG;
    if (direction != 1 && direction != -1) return HB_ERR_BADARG;
    if (n_fractions == 0 || n_displacements == 0) return HB_OK;
    if (!phi_dense || !tt || !eigvec || !fractions || !displacements || !x0w_soa || !node_idx) return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = sm_count();
    long long blocks = (n_fractions * 32 + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    k_totime<<<(unsigned)blocks, 256, 0, st>>>(tt, n_samples, period, fractions, n_fractions, node_idx);
    HB_CUDA_TRY(cudaGetLastError());
    const long long n = n_fractions * n_displacements;
    blocks = (n + 127) / 128;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    k_manifold_ics<<<(unsigned)blocks, 128, 0, st>>>(phi_dense, node_idx, eigvec, (double)direction, displacements,
                                                     n_fractions, n_displacements, x0w_soa);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

extern "C" int hb_tube_filter(const hb_tube_filter_opts *opts, int64_t n, const double *states, int32_t m,
                              double *out, int32_t *keep, void *stream)
{
    if (!opts || n < 0 || m < 1) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    if (!states || !out) return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    long long blocks = (n * 32 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_tube_filter<<<(unsigned)blocks, 256, 0, st>>>(states, n, m, *opts, out, keep);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}
