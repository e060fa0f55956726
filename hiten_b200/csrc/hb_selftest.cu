// hb_selftest.cu -- device arithmetic self-test entry point (used by tests/test_gpu_arith.py).
// Evaluates the restated div.rn / sqrt.rn fast paths of hb_common.cuh next to the compiler's own
// __ddiv_rn / __dsqrt_rn so the host can compare both with IEEE results.
#include "hb_common.cuh"

namespace {
__global__ void k_selftest(const double *a, const double *b, long long n, double *div_shared, double *div_ref,
                           double *sqrt_fast, double *sqrt_ref, double *pow_out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double y = hb_rcp_refined(b[i]);
    div_shared[i] = hb_div_with(a[i], b[i], y);
    div_ref[i] = __ddiv_rn(a[i], b[i]);
    sqrt_fast[i] = hb_sqrt_rn(b[i]);
    sqrt_ref[i] = __dsqrt_rn(b[i]);
    pow_out[i] = hb_pow(b[i], a[i]);
}
}  // namespace

extern "C" int hb_selftest_arith(const double *a, const double *b, int64_t n, double *div_shared, double *div_ref,
                                 double *sqrt_fast, double *sqrt_ref, double *pow_out, void *stream)
{
    if (n <= 0 || !a || !b || !div_shared || !div_ref || !sqrt_fast || !sqrt_ref || !pow_out) return HB_ERR_BADARG;
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    k_selftest<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(a, b, n, div_shared, div_ref, sqrt_fast,
                                                                        sqrt_ref, pow_out);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}
