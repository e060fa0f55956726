// hb_peak.cu -- FP64 FMA throughput microbenchmark (roofline denominator).
// Register-resident: every thread runs 8 independent DFMA chains; no memory traffic in the loop.
#include "hb_common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_dfma(double *sink, long long iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-9, x2 = x0 + 2e-9, x3 = x0 + 3e-9;
    double x4 = x0 + 4e-9, x5 = x0 + 5e-9, x6 = x0 + 6e-9, x7 = x0 + 7e-9;
    for (long long i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

}  // namespace

extern "C" int hb_dfma_peak(double millis, double *flops_per_s, void *stream)
{
    if (!flops_per_s) return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return HB_ERR_NODEVICE;
    HB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *sink = nullptr;
    HB_CUDA_TRY(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    HB_CUDA_TRY(cudaEventCreate(&e0));
    HB_CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    long long iters = 2000;
    float ms = 0.f;
    int rc = HB_OK;
    for (int pass = 0; pass < 3; ++pass) {
        cudaEventRecord(e0, st);
        k_dfma<<<blocks, threads, 0, st>>>(sink, iters, 0.999999, 1e-6);
        cudaEventRecord(e1, st);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { rc = (int)e; break; }
        cudaEventElapsedTime(&ms, e0, e1);
        if (pass < 2 && ms > 0.f) {
            const double want = (millis > 0 ? millis : 50.0);
            long long next = (long long)(iters * want / ms);
            iters = next < 1000 ? 1000 : next;
        }
    }
    if (rc == HB_OK) {
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        *flops_per_s = flops / (ms * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return rc;
}
