// hb_peer_put: a shard's result -- hit records and end states -- written into the gathering rank's receive buffer over
// NVLink PEER MEMORY by a kernel, with the hit count read ON THE DEVICE.
//
// The sharded form of the path (trajectories shard by index, one exchange at the end: hits + end states to the
// gathering rank) has two ways to move a shard's result.  Large shards use the copy engines (hiten_b200/sharded.py,
// PeerExchange.put: the copies of one tube run under the next tube's persistent propagation kernel, which owns every SM);
// that form needs the hit count on the HOST to size the copy, i.e. a stream synchronisation per tube.  For small shards
// (strong scaling: ~1e5 trajectories per GPU, a 4 ms step) those host round trips are a tenth of the step, the payload is
// a few MB and the SMs are idle once the pipeline ends, so this kernel does the transfer stream-ordered behind the
// pipeline with no host involvement: it reads the counters of the pipeline's workspace, writes an 8-double header into the
// shard's slot of EVERY rank's buffer (so that all ranks agree, after the closing barrier, on whether some shard
// overflowed and the host-sized fallback round is needed) and the payload into the gathering rank's.
// No counterpart in the reference (its worker pools are processes on one host: algorithms/poincare/synodic/engine.py:92-139).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hiten_b200.h"
#include "hb_common.cuh"

namespace {

constexpr int PEER_MAX_WORLD = 16;

struct PeerPutParams {
    double *slot[PEER_MAX_WORLD];   // this shard's slot in rank r's receive buffer (peer-mapped), r < world
    int world, dst;
    const double *hits;             // k records of 9 doubles
    long long hit_slots;            // records the slot has room for
    const double *yf;               // 6 * n_local doubles (SoA)
    long long n_local;
    const HbWorkspace *ws;
};

__device__ __forceinline__ void copy_doubles(double *dst, const double *src, long long n, long long tid, long long nthr)
{
    // both 16-byte aligned by construction (header 64 B, 9 * hit_slots even); 16-byte accesses, scalar tail
    const long long n2 = n >> 1;
    const double2 *s2 = reinterpret_cast<const double2 *>(src);
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    for (long long i = tid; i < n2; i += nthr) d2[i] = s2[i];
    if ((n & 1) && tid == 0) dst[n - 1] = src[n - 1];
}

__global__ void __launch_bounds__(256) k_peer_put(const PeerPutParams p)
{
    const long long k = (long long)p.ws->hit_count, dropped = (long long)p.ws->overflow;
    const long long rec_over = (long long)p.ws->rec_overflow;
    const bool sendable = dropped == 0 && rec_over == 0 && k <= p.hit_slots;
    if (blockIdx.x == 0 && threadIdx.x < p.world) {
        double *h = p.slot[threadIdx.x];
        h[0] = (double)k; h[1] = (double)p.n_local; h[2] = (double)dropped; h[3] = (double)rec_over;
        h[4] = sendable ? 1.0 : 0.0; h[5] = 0.0; h[6] = 0.0; h[7] = 0.0;
    }
    if (!sendable) return;          // the caller falls back to the host-sized exchange (every rank sees h[4] == 0)
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    double *dst = p.slot[p.dst];
    copy_doubles(dst + 8, p.hits, 9 * k, tid, nthr);
    copy_doubles(dst + 8 + 9 * p.hit_slots, p.yf, 6 * p.n_local, tid, nthr);
}

}  // namespace

extern "C" int hb_peer_put(void *const *peer_slots, int32_t world, int32_t dst, const hb_hit *hits, int64_t hit_slots,
                           const double *yf_soa, int64_t n_local, const void *workspace, void *stream)
{
    if (!peer_slots || world < 1 || world > PEER_MAX_WORLD || dst < 0 || dst >= world || hit_slots < 0 || (hit_slots & 1) ||
        n_local < 0 || !workspace || (hit_slots > 0 && !hits) || (n_local > 0 && !yf_soa))
        return HB_ERR_BADARG;
    PeerPutParams p{};
    for (int r = 0; r < world; ++r) {
        if (!peer_slots[r] || ((uintptr_t)peer_slots[r] & 15u)) return HB_ERR_BADARG;
        p.slot[r] = (double *)peer_slots[r];
    }
    if (((uintptr_t)hits & 15u) || ((uintptr_t)yf_soa & 15u)) return HB_ERR_BADARG;
    p.world = world; p.dst = dst; p.hits = (const double *)hits; p.hit_slots = hit_slots;
    p.yf = yf_soa; p.n_local = n_local; p.ws = (const HbWorkspace *)workspace;
    // a few MB over one NVLink port: 64 CTAs of plain 16-byte stores saturate it and fit beside anything still running
    k_peer_put<<<64, 256, 0, (cudaStream_t)stream>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}
