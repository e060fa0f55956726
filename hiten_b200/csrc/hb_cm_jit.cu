// hb_cm_jit.cu -- centre-manifold Poincare map, SPECIALISED at run time for the Hamiltonian at hand.
//
// The polynomial tables of a centre manifold are fixed for the life of the object (they take the reference
// ~34 s to build) while the map evaluates their gradient ~1e9 times.  Instead of interpreting a term table
// (hb_cm.cu: ~60 instructions of decoding per term) this path GENERATES the gradient as straight-line CUDA --
// powers in registers, coefficients as hex-float immediates, terms in the reference's evaluation order -- wraps it
// in the map kernel (Tao symplectic or RK4/6/8 step, crossing test, Hermite hit), compiles it for sm_100a with
// NVRTC and launches it through the driver API on the caller's stream.  Modules are cached per generated source.
// The parity build keeps every mul/add separately rounded, so results stay bit-identical to the reference.
//
// libnvrtc / libcuda are loaded lazily with dlopen: the shared library itself still loads on a machine without a
// GPU driver (the CPU test suite checks exactly that, and compiles a specialised kernel offline).
//
// Reference routines: see hb_cm.cu (same algorithms, same order of operations).
#include "hb_common.cuh"

#include <dlfcn.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct TermHost {
    double coef;
    unsigned long long ex;
};

std::string hexf(double v)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%a", v);
    return std::string(buf);
}

void appendf(std::string &s, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    s += buf;
}

// ---- code generation ---------------------------------------------------------------------------
// gradient: g[q] = sum over degree groups (each summed from 0.0 in term order) of coef * prod_v pt[v]^e_v
std::string gen_grad(const std::vector<TermHost> &terms, const int64_t *ptr)
{
    int maxe[6] = {0, 0, 0, 0, 0, 0};
    for (const TermHost &t : terms)
        for (int v = 0; v < 6; ++v) {
            const int e = (int)((t.ex >> (8 * v)) & 0xff);
            if (e > maxe[v]) maxe[v] = e;
        }
    std::string s = "DEV void grad(const double (&pt)[6], double (&g)[6])\n{\n";
    for (int v = 0; v < 6; ++v)
        for (int e = 1; e <= maxe[v]; ++e) {
            if (e == 1) appendf(s, "    const double w%d_1 = pt[%d];\n", v, v);
            else appendf(s, "    const double w%d_%d = MUL(w%d_%d, pt[%d]);\n", v, e, v, e - 1, v);
        }
    for (int q = 0; q < 6; ++q) {
        appendf(s, "    {\n        double tot = 0.0, acc = 0.0;\n");
        int dcur = -1;
        for (int64_t i = ptr[q]; i < ptr[q + 1]; ++i) {
            const TermHost &t = terms[(size_t)i];
            const int d = (int)(t.ex >> 48);
            if (d != dcur) {
                if (dcur >= 0) s += "        tot = ADD(tot, acc); acc = 0.0;\n";
                dcur = d;
            }
            std::string prod;
            for (int v = 0; v < 6; ++v) {
                const int e = (int)((t.ex >> (8 * v)) & 0xff);
                if (!e) continue;
                char w[32];
                snprintf(w, sizeof w, "w%d_%d", v, e);
                prod = prod.empty() ? std::string(w) : "MUL(" + prod + ", " + w + ")";
            }
            if (prod.empty()) prod = "1.0";
            s += "        acc = MADD(" + hexf(t.coef) + ", " + prod + ", acc);\n";
        }
        if (dcur >= 0) s += "        tot = ADD(tot, acc);\n";
        appendf(s, "        g[%d] = tot;\n    }\n", q);
    }
    s += "}\n";
    return s;
}

// The same gradient split over the four warps of a CTA (cm_map_split below): warp w evaluates the partial derivatives
// assigned to it -- whole partials, each summed exactly as in grad(), so the numbers are the same -- and the warps exchange
// them through shared memory.  Partials are dealt out by descending term count to the least loaded warp.
std::string gen_grad_parts(const std::vector<TermHost> &terms, const int64_t *ptr)
{
    int owner[6], load[4] = {0, 0, 0, 0}, order[6] = {0, 1, 2, 3, 4, 5};
    for (int a = 0; a < 6; ++a)
        for (int b = a + 1; b < 6; ++b)
            if (ptr[order[b] + 1] - ptr[order[b]] > ptr[order[a] + 1] - ptr[order[a]]) { const int t = order[a]; order[a] = order[b]; order[b] = t; }
    for (int a = 0; a < 6; ++a) {
        int w = 0;
        for (int c = 1; c < 4; ++c) if (load[c] < load[w]) w = c;
        owner[order[a]] = w;
        load[w] += (int)(ptr[order[a] + 1] - ptr[order[a]]) + 1;
    }
    std::string s;
    for (int w = 0; w < 4; ++w) {
        int maxe[6] = {0, 0, 0, 0, 0, 0};
        for (int q = 0; q < 6; ++q) {
            if (owner[q] != w) continue;
            for (int64_t i = ptr[q]; i < ptr[q + 1]; ++i)
                for (int v = 0; v < 6; ++v) {
                    const int e = (int)((terms[(size_t)i].ex >> (8 * v)) & 0xff);
                    if (e > maxe[v]) maxe[v] = e;
                }
        }
        appendf(s, "DEV void grad_part%d(const double (&pt)[6], double *gs, int lane)\n{\n", w);
        for (int v = 0; v < 6; ++v)
            for (int e = 1; e <= maxe[v]; ++e) {
                if (e == 1) appendf(s, "    const double w%d_1 = pt[%d];\n", v, v);
                else appendf(s, "    const double w%d_%d = MUL(w%d_%d, pt[%d]);\n", v, e, v, e - 1, v);
            }
        for (int q = 0; q < 6; ++q) {
            if (owner[q] != w) continue;
            appendf(s, "    {\n        double tot = 0.0, acc = 0.0;\n");
            int dcur = -1;
            for (int64_t i = ptr[q]; i < ptr[q + 1]; ++i) {
                const TermHost &t = terms[(size_t)i];
                const int d = (int)(t.ex >> 48);
                if (d != dcur) {
                    if (dcur >= 0) s += "        tot = ADD(tot, acc); acc = 0.0;\n";
                    dcur = d;
                }
                std::string prod;
                for (int v = 0; v < 6; ++v) {
                    const int e = (int)((t.ex >> (8 * v)) & 0xff);
                    if (!e) continue;
                    char wn[32];
                    snprintf(wn, sizeof wn, "w%d_%d", v, e);
                    prod = prod.empty() ? std::string(wn) : "MUL(" + prod + ", " + wn + ")";
                }
                if (prod.empty()) prod = "1.0";
                s += "        acc = MADD(" + hexf(t.coef) + ", " + prod + ", acc);\n";
            }
            if (dcur >= 0) s += "        tot = ADD(tot, acc);\n";
            appendf(s, "        gs[%d * 32 + lane] = tot;\n    }\n", q);
        }
        s += "}\n";
    }
    return s;
}

template <int S>
std::string gen_rk_step(const double (&A)[S][S], const double (&B)[S])
{
    // k[s] = rhs(y + sum_j (h*a_sj) k_j);  yn = y + sum_s (h*b_s) k_s   (backend.py:160-180); k0 = ro (rhs of so)
    std::string s;
    bool used[S];
    for (int i = 0; i < S; ++i) {
        used[i] = B[i] != 0.0;
        for (int r = i + 1; r < S; ++r) used[i] = used[i] || (A[r][i] != 0.0);
    }
    appendf(s, "            double k[%d][6], ys[6];\n", S);
    s += "            UNROLL for (int d = 0; d < 6; ++d) k[0][d] = ro[d];\n";
    for (int i = 1; i < S; ++i) {
        if (!used[i]) continue;
        s += "            UNROLL for (int d = 0; d < 6; ++d) ys[d] = so[d];\n";
        for (int j = 0; j < i; ++j)
            if (A[i][j] != 0.0)
                appendf(s, "            { const double ha = MUL(DT, %s); UNROLL for (int d = 0; d < 6; ++d) ys[d] = MADD(ha, k[%d][d], ys[d]); }\n",
                        hexf(A[i][j]).c_str(), j);
        appendf(s, "            RHS(ys, k[%d]);\n", i);
    }
    s += "            UNROLL for (int d = 0; d < 6; ++d) sn[d] = so[d];\n";
    for (int j = 0; j < S; ++j)
        if (B[j] != 0.0)
            appendf(s, "            { const double hb = MUL(DT, %s); UNROLL for (int d = 0; d < 6; ++d) sn[d] = MADD(hb, k[%d][d], sn[d]); }\n",
                    hexf(B[j]).c_str(), j);
    return s;
}

// steps a seed may take per round before it is parked and re-queued (see cm_map)
constexpr int HB_CM_QUOTA = 160;

const char *PRELUDE = R"SRC(
typedef unsigned long long u64;
struct Ws { u64 cursor, hit_count, overflow, pad[29]; };
#define DEV __device__ __forceinline__
#define UNROLL _Pragma("unroll")
#if PARITY
DEV double ADD(double a, double b) { return __dadd_rn(a, b); }
DEV double SUB(double a, double b) { return __dsub_rn(a, b); }
DEV double MUL(double a, double b) { return __dmul_rn(a, b); }
DEV double DIV(double a, double b) { return __ddiv_rn(a, b); }
DEV double MADD(double a, double b, double c) { return __dadd_rn(c, __dmul_rn(a, b)); }
#else
DEV double ADD(double a, double b) { return a + b; }
DEV double SUB(double a, double b) { return a - b; }
DEV double MUL(double a, double b) { return a * b; }
DEV double DIV(double a, double b) { return a / b; }
DEV double MADD(double a, double b, double c) { return fma(a, b, c); }
#endif
)SRC";

const char *KERNEL_HEAD = R"SRC(
DEV void rhs(const double (&y)[6], double (&dy)[6])          // [dH/dP, -dH/dQ]
{
    double g[6];
    grad(y, g);
    dy[0] = g[3]; dy[1] = g[4]; dy[2] = g[5];
    dy[3] = -g[0]; dy[4] = -g[1]; dy[5] = -g[2];
}
#define GRAD(a, b) grad(a, b)
#define RHS(a, b) rhs(a, b)
DEV double hermite(double s, double y0, double y1, double dy0, double dy1)
{
    const double oms = SUB(1.0, s);
    const double oms2 = MUL(oms, oms), s2 = MUL(s, s);
    const double h00 = MUL(ADD(1.0, MUL(2.0, s)), oms2);
    const double h10 = MUL(s, oms2);
    const double h01 = MUL(s2, SUB(3.0, MUL(2.0, s)));
    const double h11 = MUL(s2, SUB(s, 1.0));
    return ADD(ADD(ADD(MUL(h00, y0), MUL(MUL(h10, dy0), DT)), MUL(h01, y1)), MUL(MUL(h11, dy1), DT));
}
// Round r of the map.  Work items: round 0 = all seeds; later rounds = the seeds that used up their step quota in the
// previous round (list_in, *count_in), resumed from cont[idx][16] = {elapsed, it, so[6], ro[6]}.  A seed that neither
// returns to the section nor reaches MAX_STEPS within QUOTA steps of this round parks its state and is appended to
// list_out.  Return times are heavy-tailed (median 143 steps, max > 1300 at dt = 0.01): without the quota a warp
// idles behind its slowest seed (ncu: 7.5 of 32 lanes active).
extern "C" __global__ void __launch_bounds__(256) cm_map(const double *seeds, long long n, int *flags, double *out,
                                                         double *t_out, u64 *cursor, const int *list_in,
                                                         const int *count_in, int *list_out, int *count_out,
                                                         double *cont)
{
    // ONE flat loop: a lane that is done with its item pulls the next one at the top of the same loop (with the
    // item loop nested around the step loop the lanes of a warp reconverge behind the step loop).
    const long long count = list_in ? (long long)*count_in : n;
    if (list_in && count <= SPLIT_MAX) return;               // a short list: cm_map_split takes this round
    double so[6], sn[6], rn[6], ro[6];
    double elapsed = 0.0;
    long long idx = -1;
    int it = 0, it_stop = 0;
    bool have = false, exhausted = false;
    // First fill WITHOUT the queue: item = rank * 32 + lane, warps ranked warp-index-major across the grid (warp 0 of every
    // CTA first, then warp 1 of every CTA, ...).  A short work list -- the later rounds, when a few thousand slow seeds
    // are left -- then occupies FULL warps, one per CTA and so about one per SM sub-partition, instead of one or two
    // lanes of every resident warp (with the per-lane queue each of the 16 warps of an SM held a couple of seeds and a
    // step cost 16 warps' worth of issue slots).  Refills while a round is running still come from the queue, which
    // starts behind the statically assigned items.
    const long long lanes_total = (long long)gridDim.x * blockDim.x;
    bool first_fill = true;
    for (;;) {
        if (!have && !exhausted) {
            long long k;
            if (first_fill) {
                k = ((long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31);
                first_fill = false;
            } else {
                k = lanes_total + (long long)atomicAdd(cursor, 1ULL);
            }
            if (k < count) {
                if (list_in) {
                    idx = list_in[k];
                    const double *c = cont + idx * 16;
                    elapsed = c[0]; it = (int)c[1];
                    UNROLL for (int d = 0; d < 6; ++d) { so[d] = c[2 + d]; ro[d] = c[8 + d]; }
                    have = true;
                } else {
                    idx = k;
                    const double *sd = seeds + idx * 4;
                    so[0] = 0.0; so[1] = sd[0]; so[2] = sd[2]; so[3] = 0.0; so[4] = sd[1]; so[5] = sd[3];
#if !TAO
                    RHS(so, ro);
#else
                    UNROLL for (int d = 0; d < 6; ++d) ro[d] = 0.0;
#endif
                    elapsed = 0.0; it = 0;
                    have = MAX_STEPS > 0;
                    if (!have) {
                        flags[idx] = 0; t_out[idx] = 0.0;
                        out[idx * 4 + 0] = 0.0; out[idx * 4 + 1] = 0.0; out[idx * 4 + 2] = 0.0; out[idx * 4 + 3] = 0.0;
                    }
                }
                it_stop = it + QUOTA;
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(0xffffffffu, !have && exhausted)) break;
        if (have) {
)SRC";

const char *TAO_STEP = R"SRC(
            double Q[3] = {so[0], so[1], so[2]}, P[3] = {so[3], so[4], so[5]};
            double X[3] = {so[0], so[1], so[2]}, Y[3] = {so[3], so[4], so[5]};
#pragma unroll 1
            for (int j = 0; j < N_SUB; ++j) {
                const double ts = SUB_TS[j], hd = MUL(0.5, ts), c = SUB_COS[j], s = SUB_SIN[j];
#pragma unroll 1
                for (int ph = 0; ph < 5; ++ph) {
                    if (ph == 2) {
                        UNROLL for (int i = 0; i < 3; ++i) {
                            const double qpx = ADD(Q[i], X[i]), qmx = SUB(Q[i], X[i]);
                            const double ppy = ADD(P[i], Y[i]), pmy = SUB(P[i], Y[i]);
                            Q[i] = MUL(0.5, ADD(ADD(qpx, MUL(c, qmx)), MUL(s, pmy)));
                            P[i] = MUL(0.5, ADD(SUB(ppy, MUL(s, qmx)), MUL(c, pmy)));
                            X[i] = MUL(0.5, SUB(SUB(qpx, MUL(c, qmx)), MUL(s, pmy)));
                            Y[i] = MUL(0.5, SUB(ADD(ppy, MUL(s, qmx)), MUL(c, pmy)));
                        }
                    } else {
                        const bool isA = (ph == 0) || (ph == 4);         // phi_a: (Q,Y); phi_b: (X,P)
                        double pt[6], g[6];
                        UNROLL for (int i = 0; i < 3; ++i) { pt[i] = isA ? Q[i] : X[i]; pt[3 + i] = isA ? Y[i] : P[i]; }
                        GRAD(pt, g);
                        UNROLL for (int i = 0; i < 3; ++i) {
                            const double dq = MUL(hd, g[i]), dp = MUL(hd, g[3 + i]);
                            if (isA) { P[i] = SUB(P[i], dq); X[i] = ADD(X[i], dp); }
                            else { Q[i] = ADD(Q[i], dp); Y[i] = SUB(Y[i], dq); }
                        }
                    }
                }
            }
            UNROLL for (int i = 0; i < 3; ++i) { sn[i] = Q[i]; sn[3 + i] = P[i]; }
)SRC";

const char *KERNEL_TAIL = R"SRC(
            RHS(sn, rn);
            const double f_old = so[FIDX], f_new = sn[FIDX];
            bool crossed = false;
            if (!(MUL(f_old, f_new) >= 0.0)) crossed = GOOD_DIR;
            double tc = 0.0, o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0;
            bool done = false;
            if (crossed) {
                const double alpha = DIV(f_old, SUB(f_old, f_new));
#if TAO
                RHS(so, ro);
#endif
                o0 = hermite(alpha, so[1], sn[1], ro[1], rn[1]);
                o1 = hermite(alpha, so[4], sn[4], ro[4], rn[4]);
                o2 = hermite(alpha, so[2], sn[2], ro[2], rn[2]);
                o3 = hermite(alpha, so[5], sn[5], ro[5], rn[5]);
                tc = ADD(elapsed, MUL(alpha, DT));
                done = true;
            } else {
                UNROLL for (int d = 0; d < 6; ++d) { so[d] = sn[d]; ro[d] = rn[d]; }
                elapsed = ADD(elapsed, DT);
                done = ++it >= MAX_STEPS;
            }
            if (done) {
                flags[idx] = crossed ? 1 : 0;
                t_out[idx] = tc;
                out[idx * 4 + 0] = o0; out[idx * 4 + 1] = o1; out[idx * 4 + 2] = o2; out[idx * 4 + 3] = o3;
                have = false;
            } else if (it >= it_stop) {                       // quota used up: park the state for the next round
                double *c = cont + idx * 16;
                c[0] = elapsed; c[1] = (double)it;
                UNROLL for (int d = 0; d < 6; ++d) { c[2 + d] = so[d]; c[8 + d] = ro[d]; }
                list_out[atomicAdd(count_out, 1)] = (int)idx;
                have = false;
            }
        }
    }
}
)SRC";

// cm_map_split: the late rounds of the map.  Return times are heavy-tailed, so the last rounds hold a few thousand seeds
// and the time of a return is the slowest seed's ~1300 sequential steps: a warp alone on an SM sub-partition issues one
// FP64 instruction every 2.9 cycles whatever the rest of the machine does.  Here FOUR warps (one per sub-partition) share
// 32 seeds: every warp carries the full state of the CTA's seeds, evaluates its quarter of the gradient's partial
// derivatives (grad_part<w>) and the warps exchange them through shared memory -- one CTA barrier per gradient, double
// buffered.  Every partial is summed by one warp in the reference's order, so the numbers are those of cm_map; all warps
// apply the same updates and take the same decisions, warp 0 writes.  Barriers must be reached by every thread, so the
// step is executed unconditionally (idle slots integrate a dummy state) and the crossing-only gradient is taken CTA-wide.
const char *SPLIT_HEAD = R"SRC(
#undef GRAD
#undef RHS
DEV void grad_split(const double (&pt)[6], double (&g)[6], double *gsh, int &ph, int w, int lane)
{
    double *gs = gsh + ph * 192;
    if (w == 0) grad_part0(pt, gs, lane);
    else if (w == 1) grad_part1(pt, gs, lane);
    else if (w == 2) grad_part2(pt, gs, lane);
    else grad_part3(pt, gs, lane);
    __syncthreads();
    UNROLL for (int q = 0; q < 6; ++q) g[q] = gs[q * 32 + lane];
    ph ^= 1;
}
DEV void rhs_split(const double (&y)[6], double (&dy)[6], double *gsh, int &ph, int w, int lane)
{
    double g[6];
    grad_split(y, g, gsh, ph, w, lane);
    dy[0] = g[3]; dy[1] = g[4]; dy[2] = g[5];
    dy[3] = -g[0]; dy[4] = -g[1]; dy[5] = -g[2];
}
#define GRAD(a, b) grad_split(a, b, gsh, xbuf_, warp_, lane_)
#define RHS(a, b) rhs_split(a, b, gsh, xbuf_, warp_, lane_)
extern "C" __global__ void __launch_bounds__(128) cm_map_split(const double *seeds, long long n, int *flags, double *out,
                                                               double *t_out, u64 *cursor, const int *list_in,
                                                               const int *count_in, int *list_out, int *count_out,
                                                               double *cont)
{
    const long long count = list_in ? (long long)*count_in : n;
    if (count > SPLIT_MAX) return;                           // a long list: cm_map takes this round
    __shared__ double gsh[2 * 192];
    __shared__ long long item_sh[32];
    const int warp_ = threadIdx.x >> 5, lane_ = threadIdx.x & 31;
    int xbuf_ = 0;
    double so[6], sn[6], rn[6], ro[6];
    UNROLL for (int d = 0; d < 6; ++d) { so[d] = 0.0; ro[d] = 0.0; }
    double elapsed = 0.0;
    long long idx = -1;
    int it = 0, it_stop = 0;
    bool have = false, exhausted = false, first_fill = true;
    for (;;) {
        // refill: warp 0 hands out the items, every warp loads the same seed into its replica of slot `lane_`
        if (warp_ == 0) {
            long long k = -2;                                // -2: slot busy (or done), nothing to hand out
            if (!have && !exhausted) {
                if (first_fill) k = (long long)blockIdx.x * 32 + lane_;
                else k = (long long)gridDim.x * 32 + (long long)atomicAdd(cursor, 1ULL);
                if (k >= count) k = -1;
            }
            item_sh[lane_] = k;
        }
        first_fill = false;
        __syncthreads();
        const long long k = item_sh[lane_];
        if (k >= 0) {
            if (list_in) {
                idx = list_in[k];
                const double *c = cont + idx * 16;
                elapsed = c[0]; it = (int)c[1];
                UNROLL for (int d = 0; d < 6; ++d) { so[d] = c[2 + d]; ro[d] = c[8 + d]; }
                have = true;
            } else {
                idx = k;
                const double *sd = seeds + idx * 4;
                so[0] = 0.0; so[1] = sd[0]; so[2] = sd[2]; so[3] = 0.0; so[4] = sd[1]; so[5] = sd[3];
                UNROLL for (int d = 0; d < 6; ++d) ro[d] = 0.0;
                elapsed = 0.0; it = 0;
                have = MAX_STEPS > 0;
                if (!have && warp_ == 0) {
                    flags[idx] = 0; t_out[idx] = 0.0;
                    out[idx * 4 + 0] = 0.0; out[idx * 4 + 1] = 0.0; out[idx * 4 + 2] = 0.0; out[idx * 4 + 3] = 0.0;
                }
            }
            it_stop = it + QUOTA;
        } else if (k == -1) {
            exhausted = true;
        }
        const bool fresh = (k >= 0) && !list_in;
        if (__syncthreads_and(!have && exhausted)) break;
#if !TAO
        if (__syncthreads_or(fresh)) {                       // rhs of a new seed (fixed-step RK keeps f(so) across steps)
            double r0[6];
            RHS(so, r0);
            if (fresh) { UNROLL for (int d = 0; d < 6; ++d) ro[d] = r0[d]; }
        }
#endif
        {
)SRC";

const char *SPLIT_TAIL = R"SRC(
            RHS(sn, rn);
            const double f_old = so[FIDX], f_new = sn[FIDX];
            bool crossed = false;
            if (have && !(MUL(f_old, f_new) >= 0.0)) crossed = GOOD_DIR;
            double tc = 0.0, o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0;
            bool done = false;
#if TAO
            if (__syncthreads_or(crossed)) {                 // the crossing-only gradient, taken by the whole CTA
                double r0[6];
                RHS(so, r0);
                if (crossed) { UNROLL for (int d = 0; d < 6; ++d) ro[d] = r0[d]; }
            }
#endif
            if (crossed) {
                const double alpha = DIV(f_old, SUB(f_old, f_new));
                o0 = hermite(alpha, so[1], sn[1], ro[1], rn[1]);
                o1 = hermite(alpha, so[4], sn[4], ro[4], rn[4]);
                o2 = hermite(alpha, so[2], sn[2], ro[2], rn[2]);
                o3 = hermite(alpha, so[5], sn[5], ro[5], rn[5]);
                tc = ADD(elapsed, MUL(alpha, DT));
                done = true;
            } else if (have) {
                UNROLL for (int d = 0; d < 6; ++d) { so[d] = sn[d]; ro[d] = rn[d]; }
                elapsed = ADD(elapsed, DT);
                done = ++it >= MAX_STEPS;
            }
            if (have && done) {
                if (warp_ == 0) {
                    flags[idx] = crossed ? 1 : 0;
                    t_out[idx] = tc;
                    out[idx * 4 + 0] = o0; out[idx * 4 + 1] = o1; out[idx * 4 + 2] = o2; out[idx * 4 + 3] = o3;
                }
                have = false;
            } else if (have && it >= it_stop) {               // quota used up: park the state for the next round
                if (warp_ == 0) {
                    double *c = cont + idx * 16;
                    c[0] = elapsed; c[1] = (double)it;
                    UNROLL for (int d = 0; d < 6; ++d) { c[2 + d] = so[d]; c[8 + d] = ro[d]; }
                    list_out[atomicAdd(count_out, 1)] = (int)idx;
                }
                have = false;
            }
            if (!have) { UNROLL for (int d = 0; d < 6; ++d) { so[d] = 0.0; ro[d] = 0.0; } }   // idle slot: dummy state
        }
    }
}
)SRC";

// seeds per round up to which cm_map_split runs it: four 128-thread CTAs per SM, 32 seeds each
constexpr int HB_CM_SPLIT_MAX = 148 * 4 * 32;

std::string gen_source(const std::vector<TermHost> &terms, const int64_t *ptr, const hb_cm_opts &o)
{
    std::string s;
    const bool tao = o.method == HB_SYMPLECTIC;
    appendf(s, "#define PARITY %d\n#define TAO %d\n#define MAX_STEPS %d\n#define DT %s\n#define QUOTA %d\n#define SPLIT_MAX %d\n",
            o.arith == HB_ARITH_PARITY ? 1 : 0, tao ? 1 : 0, o.max_steps, hexf(o.dt).c_str(), HB_CM_QUOTA, HB_CM_SPLIT_MAX);
    static const int fidx[4] = {1, 4, 2, 5};                       // q2, p2, q3, p3 in [q1,q2,q3,p1,p2,p3]
    static const char *good[4] = {"(sn[4] > 0.0)", "(rn[1] > 0.0)", "(sn[5] > 0.0)", "(rn[2] > 0.0)"};
    appendf(s, "#define FIDX %d\n#define GOOD_DIR %s\n", fidx[o.section], good[o.section]);
    if (tao) {
        appendf(s, "#define N_SUB %d\n", o.n_sub);
        for (const char *nm : {"SUB_TS", "SUB_COS", "SUB_SIN"}) {
            const double *arr = !strcmp(nm, "SUB_TS") ? o.sub_ts : !strcmp(nm, "SUB_COS") ? o.sub_cos : o.sub_sin;
            appendf(s, "__device__ const double %s[%d] = {", nm, o.n_sub);
            for (int j = 0; j < o.n_sub; ++j) s += hexf(arr[j]) + (j + 1 < o.n_sub ? ", " : "");
            s += "};\n";
        }
    }
    s += PRELUDE;
    s += gen_grad(terms, ptr);
    s += gen_grad_parts(terms, ptr);
    std::string step;
    if (tao) step = TAO_STEP;
    else if (o.method == HB_RK4) step = gen_rk_step<4>(HB_RK4_A, HB_RK4_B);
    else if (o.method == HB_RK6) step = gen_rk_step<7>(HB_RK6_A, HB_RK6_B);
    else step = gen_rk_step<13>(HB_RK8_A, HB_RK8_B);
    s += KERNEL_HEAD;
    s += step;
    s += KERNEL_TAIL;
    s += SPLIT_HEAD;
    s += step;
    s += SPLIT_TAIL;
    return s;
}

// ---- the Tao integrator over a time grid (_ExtendedSymplectic.integrate, symplectic.py:564-782), specialised --------
// Same generated gradient; the per-interval Tao parameters come from hb_tao_grid_prepare's table in HBM
// (tab[m-1][3][n_sub], warp-uniform loads).  One trajectory per thread; `event` selects the terminal-plane-event form.
const char *GRID_KERNEL = R"SRC(
DEV void rhs(const double (&y)[6], double (&dy)[6])          // [dH/dP, -dH/dQ]
{
    double g[6];
    grad(y, g);
    dy[0] = g[3]; dy[1] = g[4]; dy[2] = g[5];
    dy[3] = -g[0]; dy[4] = -g[1]; dy[5] = -g[2];
}
DEV double pick(const double (&v)[6], int i)
{
    return (i == 0) ? v[0] : (i == 1) ? v[1] : (i == 2) ? v[2] : (i == 3) ? v[3] : (i == 4) ? v[4] : v[5];
}
DEV void hermite6(const double (&y0)[6], const double (&f0)[6], const double (&y1)[6], const double (&f1)[6], double x,
                  double h, double (&out)[6])                 // _hermite_eval_dense_symplectic (symplectic.py:229-279)
{
    const double x2 = MUL(x, x), x3 = MUL(x2, x);
    const double H00 = ADD(SUB(MUL(2.0, x3), MUL(3.0, x2)), 1.0);
    const double H10 = ADD(SUB(x3, MUL(2.0, x2)), x);
    const double H01 = ADD(MUL(-2.0, x3), MUL(3.0, x2));
    const double H11 = SUB(x3, x2);
    UNROLL for (int d = 0; d < 6; ++d)
        out[d] = ADD(ADD(ADD(MUL(H00, y0[d]), MUL(H10, MUL(h, f0[d]))), MUL(H01, y1[d])), MUL(H11, MUL(h, f1[d])));
}
DEV bool ev_crossed(double gp, double gn, int dir)            // utils.py:14-39
{
    if (dir == 0) return (gp < 0.0 && gn > 0.0) || (gp > 0.0 && gn < 0.0) || (gn == 0.0);
    if (dir > 0) return (gp < 0.0 && gn > 0.0) || (gn == 0.0);
    return (gp > 0.0 && gn < 0.0) || (gn == 0.0);
}
DEV bool ev_side(double gl, double gm, int dir)               // utils.py:43-69
{
    if (dir == 0) return (gl < 0.0 && gm > 0.0) || (gl > 0.0 && gm < 0.0);
    if (dir > 0) return (gl < 0.0 && gm > 0.0);
    return (gl > 0.0 && gm < 0.0);
}
extern "C" __global__ void __launch_bounds__(128) symp_grid(const double *y0, long long n, int m, int n_sub,
                                                            const double *tab, const double *t_vals, double *traj,
                                                            int event, int ev_idx, int ev_dir, double ev_off,
                                                            double xtol, double gtol, int *hit, double *t_hit,
                                                            double *y_hit, int *n_rows, u64 *cursor)
{
    for (;;) {
        const long long idx = (long long)atomicAdd(cursor, 1ULL);
        if (idx >= n) break;
        const double *s0 = y0 + idx * 6;
        double Q[3], P[3], X[3], Y[3], yo[6], fo[6], yh[6], g_old = 0.0, th = 0.0;
        UNROLL for (int d = 0; d < 6; ++d) { yo[d] = s0[d]; fo[d] = 0.0; yh[d] = 0.0; }
        UNROLL for (int i = 0; i < 3; ++i) { Q[i] = X[i] = yo[i]; P[i] = Y[i] = yo[3 + i]; }
        double *rows = traj ? traj + idx * (long long)m * 6 : nullptr;
        if (rows) { UNROLL for (int d = 0; d < 6; ++d) rows[d] = yo[d]; }
        if (event) { rhs(yo, fo); g_old = SUB(pick(yo, ev_idx), ev_off); }
        int got = 0, nr = m;
        for (int i = 0; i < m - 1; ++i) {
            const double *tb = tab + (size_t)i * 3 * n_sub;
#pragma unroll 1
            for (int j = 0; j < n_sub; ++j) {
                const double ts = tb[j], hd = MUL(0.5, ts), c = tb[n_sub + j], s = tb[2 * n_sub + j];
#pragma unroll 1
                for (int ph = 0; ph < 5; ++ph) {
                    if (ph == 2) {
                        UNROLL for (int i3 = 0; i3 < 3; ++i3) {
                            const double qpx = ADD(Q[i3], X[i3]), qmx = SUB(Q[i3], X[i3]);
                            const double ppy = ADD(P[i3], Y[i3]), pmy = SUB(P[i3], Y[i3]);
                            Q[i3] = MUL(0.5, ADD(ADD(qpx, MUL(c, qmx)), MUL(s, pmy)));
                            P[i3] = MUL(0.5, ADD(SUB(ppy, MUL(s, qmx)), MUL(c, pmy)));
                            X[i3] = MUL(0.5, SUB(SUB(qpx, MUL(c, qmx)), MUL(s, pmy)));
                            Y[i3] = MUL(0.5, SUB(ADD(ppy, MUL(s, qmx)), MUL(c, pmy)));
                        }
                    } else {
                        const bool isA = (ph == 0) || (ph == 4);         // phi_a: (Q,Y); phi_b: (X,P)
                        double pt[6], g[6];
                        UNROLL for (int i3 = 0; i3 < 3; ++i3) { pt[i3] = isA ? Q[i3] : X[i3]; pt[3 + i3] = isA ? Y[i3] : P[i3]; }
                        grad(pt, g);
                        UNROLL for (int i3 = 0; i3 < 3; ++i3) {
                            const double dq = MUL(hd, g[i3]), dp = MUL(hd, g[3 + i3]);
                            if (isA) { P[i3] = SUB(P[i3], dq); X[i3] = ADD(X[i3], dp); }
                            else { Q[i3] = ADD(Q[i3], dp); Y[i3] = SUB(Y[i3], dq); }
                        }
                    }
                }
            }
            double yn[6] = {Q[0], Q[1], Q[2], P[0], P[1], P[2]};
            if (event) {
                double fn[6];
                rhs(yn, fn);
                const double g_new = SUB(pick(yn, ev_idx), ev_off);
                if (ev_crossed(g_old, g_new, ev_dir)) {
                    const double t0 = t_vals[i], h = SUB(t_vals[i + 1], t0);
                    double a = 0.0, b = 1.0, g_left = g_old, xh = 1.0;
                    bool done = false;
                    for (int it = 0; it < 128; ++it) {
                        const double mid = MUL(0.5, ADD(a, b));
                        hermite6(yo, fo, yn, fn, mid, h, yh);
                        const double g_mid = SUB(pick(yh, ev_idx), ev_off);
                        if (fabs(g_mid) <= gtol) { xh = mid; done = true; break; }
                        if (ev_side(g_left, g_mid, ev_dir)) b = mid;
                        else { a = mid; g_left = g_mid; }
                        if (MUL(SUB(b, a), fabs(h)) <= xtol) break;
                    }
                    if (!done) { xh = b; hermite6(yo, fo, yn, fn, b, h, yh); }
                    th = ADD(t0, MUL(xh, h));
                    got = 1;
                    nr = i + 1;
                    break;
                }
                g_old = g_new;
                UNROLL for (int d = 0; d < 6; ++d) { yo[d] = yn[d]; fo[d] = fn[d]; }
            }
            if (rows) {
                double *o = rows + (long long)(i + 1) * 6;
                UNROLL for (int d = 0; d < 6; ++d) o[d] = yn[d];
            }
        }
        if (event) {
            if (!got) {
                th = t_vals[m - 1];
                UNROLL for (int d = 0; d < 6; ++d) yh[d] = yo[d];
            }
            hit[idx] = got;
            t_hit[idx] = th;
            n_rows[idx] = nr;
            UNROLL for (int d = 0; d < 6; ++d) y_hit[idx * 6 + d] = yh[d];
        }
    }
}
)SRC";

std::string gen_grid_source(const std::vector<TermHost> &terms, const int64_t *ptr, int arith)
{
    std::string s;
    appendf(s, "#define PARITY %d\n", arith == HB_ARITH_PARITY ? 1 : 0);
    s += PRELUDE;
    s += gen_grad(terms, ptr);
    s += GRID_KERNEL;
    return s;
}

// ---- NVRTC + driver API through dlopen -----------------------------------------------------------
typedef int (*nvrtcCreate_t)(void **, const char *, const char *, int, const char *const *, const char *const *);
typedef int (*nvrtcCompile_t)(void *, int, const char *const *);
typedef int (*nvrtcSize_t)(void *, size_t *);
typedef int (*nvrtcGet_t)(void *, char *);
typedef int (*nvrtcDestroy_t)(void **);
typedef int (*cuModuleLoadData_t)(void **, const void *);
typedef int (*cuModuleGetFunction_t)(void **, void *, const char *);
typedef int (*cuLaunchKernel_t)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *,
                                void **, void **);
typedef int (*cuOccupancy_t)(int *, void *, int, size_t);

struct Api {
    void *nvrtc = nullptr, *cuda = nullptr;
    nvrtcCreate_t create = nullptr;
    nvrtcCompile_t compile = nullptr;
    nvrtcSize_t cubin_size = nullptr, log_size = nullptr;
    nvrtcGet_t cubin = nullptr, log = nullptr;
    nvrtcDestroy_t destroy = nullptr;
    cuModuleLoadData_t load = nullptr;
    cuModuleGetFunction_t getfn = nullptr;
    cuLaunchKernel_t launch = nullptr;
    cuOccupancy_t occ = nullptr;
};

std::mutex g_mu;
Api g_api;
// (device ordinal, generated source) -> CUmodule, (the same, kernel name) -> CUfunction.  A module belongs to the context
// that was current at cuModuleLoadData -- the primary context of the current device -- so every device gets its own.
std::unordered_map<std::string, void *> g_modules;
std::unordered_map<std::string, void *> g_functions;
inline std::string fn_key(const std::string &src)
{
    int dev = 0;
    cudaGetDevice(&dev);
    return std::to_string(dev) + "|" + src;
}
std::string g_last_log;

bool load_nvrtc()
{
    if (g_api.create) return true;
    for (const char *nm : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"}) {
        g_api.nvrtc = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (g_api.nvrtc) break;
    }
    if (!g_api.nvrtc) return false;
    g_api.create = (nvrtcCreate_t)dlsym(g_api.nvrtc, "nvrtcCreateProgram");
    g_api.compile = (nvrtcCompile_t)dlsym(g_api.nvrtc, "nvrtcCompileProgram");
    g_api.cubin_size = (nvrtcSize_t)dlsym(g_api.nvrtc, "nvrtcGetCUBINSize");
    g_api.cubin = (nvrtcGet_t)dlsym(g_api.nvrtc, "nvrtcGetCUBIN");
    g_api.log_size = (nvrtcSize_t)dlsym(g_api.nvrtc, "nvrtcGetProgramLogSize");
    g_api.log = (nvrtcGet_t)dlsym(g_api.nvrtc, "nvrtcGetProgramLog");
    g_api.destroy = (nvrtcDestroy_t)dlsym(g_api.nvrtc, "nvrtcDestroyProgram");
    return g_api.create && g_api.compile && g_api.cubin_size && g_api.cubin && g_api.destroy;
}

bool load_driver()
{
    if (g_api.launch) return true;
    g_api.cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!g_api.cuda) return false;
    g_api.load = (cuModuleLoadData_t)dlsym(g_api.cuda, "cuModuleLoadData");
    g_api.getfn = (cuModuleGetFunction_t)dlsym(g_api.cuda, "cuModuleGetFunction");
    g_api.launch = (cuLaunchKernel_t)dlsym(g_api.cuda, "cuLaunchKernel");
    g_api.occ = (cuOccupancy_t)dlsym(g_api.cuda, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
    return g_api.load && g_api.getfn && g_api.launch;
}

int compile_cubin(const std::string &src, std::vector<char> &cubin)
{
    if (!load_nvrtc()) return HB_ERR_UNSUPPORTED;
    void *prog = nullptr;
    if (g_api.create(&prog, src.c_str(), "hb_cm_specialised.cu", 0, nullptr, nullptr) != 0) return HB_ERR_BADARG;
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    const int rc = g_api.compile(prog, 4, opts);
    if (rc != 0) {
        size_t ls = 0;
        if (g_api.log_size && g_api.log && g_api.log_size(prog, &ls) == 0 && ls > 1) {
            g_last_log.resize(ls);
            g_api.log(prog, &g_last_log[0]);
        }
        g_api.destroy(&prog);
        return HB_ERR_BADARG;
    }
    size_t n = 0;
    g_api.cubin_size(prog, &n);
    cubin.resize(n);
    g_api.cubin(prog, cubin.data());
    g_api.destroy(&prog);
    return HB_OK;
}

int check_opts(const hb_polyham *ham, const hb_cm_opts *o)
{
    if (!ham || !o) return HB_ERR_BADARG;
    if (ham->n_dof != 3 || ham->max_deg < 0 || ham->max_deg > 60) return HB_ERR_UNSUPPORTED;
    if (o->section < 0 || o->section > 3 || o->max_steps < 0) return HB_ERR_BADARG;
    if (o->arith != HB_ARITH_PARITY && o->arith != HB_ARITH_FAST) return HB_ERR_BADARG;
    if (o->method == HB_SYMPLECTIC) { if (o->n_sub <= 0 || o->n_sub > HB_MAX_TAO_SUBSTEPS) return HB_ERR_BADARG; }
    else if (o->method != HB_RK4 && o->method != HB_RK6 && o->method != HB_RK8) return HB_ERR_UNSUPPORTED;
    return HB_OK;
}

}  // namespace

// Host-only: generate and compile the specialised kernel for a term table in HOST memory; returns the cubin size.
// Needs libnvrtc but no GPU -- used by the CPU test-suite and by tools that want to inspect the generated code.
extern "C" int hb_cm_jit_compile_host(const void *terms_host, const int64_t *ptr, int32_t max_deg, const hb_cm_opts *opts,
                                      int64_t *cubin_bytes, char *source_out, int64_t source_cap)
{
    hb_polyham h{};
    h.n_dof = 3; h.max_deg = max_deg;
    int rc = check_opts(&h, opts);
    if (rc != HB_OK || !ptr || (ptr[6] > 0 && !terms_host)) return rc != HB_OK ? rc : HB_ERR_BADARG;
    std::vector<TermHost> terms((size_t)ptr[6]);
    if (ptr[6]) memcpy(terms.data(), terms_host, sizeof(TermHost) * (size_t)ptr[6]);
    const std::string src = gen_source(terms, ptr, *opts);
    if (source_out && source_cap > 0) {
        const size_t n = src.size() < (size_t)source_cap - 1 ? src.size() : (size_t)source_cap - 1;
        memcpy(source_out, src.data(), n);
        source_out[n] = 0;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    std::vector<char> cubin;
    rc = compile_cubin(src, cubin);
    if (rc == HB_OK && cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    if (rc != HB_OK && source_out && source_cap > 0 && !g_last_log.empty()) {
        const size_t n = g_last_log.size() < (size_t)source_cap - 1 ? g_last_log.size() : (size_t)source_cap - 1;
        memcpy(source_out, g_last_log.data(), n);
        source_out[n] = 0;
    }
    return rc;
}

// source -> CUfunction (compiled and loaded once per process, device and source)
static int get_function(const std::string &src, const char *name, void **fn_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    const std::string mkey = fn_key(src), key = std::string(name) + "|" + mkey;
    auto it = g_functions.find(key);
    if (it != g_functions.end()) { *fn_out = it->second; return HB_OK; }
    if (!load_driver()) return HB_ERR_NODEVICE;
    void *mod = nullptr, *fn = nullptr;
    auto im = g_modules.find(mkey);
    if (im != g_modules.end()) mod = im->second;
    else {
        std::vector<char> cubin;
        const int rc = compile_cubin(src, cubin);
        if (rc != HB_OK) return rc;
        const int e = g_api.load(&mod, cubin.data());
        if (e != 0) return 1000 + e;
        g_modules.emplace(mkey, mod);
    }
    const int e = g_api.getfn(&fn, mod, name);
    if (e != 0) return 1000 + e;
    g_functions.emplace(key, fn);
    *fn_out = fn;
    return HB_OK;
}

// Host-only: generate + compile the specialised GRID kernel for a term table in HOST memory (no GPU needed).
extern "C" int hb_symp_jit_compile_host(const void *terms_host, const int64_t *ptr, int32_t arith, int64_t *cubin_bytes)
{
    if (!ptr || (ptr[6] > 0 && !terms_host) || (arith != HB_ARITH_PARITY && arith != HB_ARITH_FAST)) return HB_ERR_BADARG;
    std::vector<TermHost> terms((size_t)ptr[6]);
    if (ptr[6]) memcpy(terms.data(), terms_host, sizeof(TermHost) * (size_t)ptr[6]);
    const std::string src = gen_grid_source(terms, ptr, arith);
    std::lock_guard<std::mutex> lk(g_mu);
    std::vector<char> cubin;
    const int rc = compile_cubin(src, cubin);
    if (rc == HB_OK && cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    return rc;
}

// hb_ham_symplectic_dense / hb_ham_symplectic_event with the run-time specialised gradient: same arguments, same
// results (bit-identical in the parity variant).  `ev` == NULL selects the grid form.
extern "C" int hb_ham_symplectic_jit(const hb_polyham *ham, const hb_symp_opts *o, const hb_event *ev, int64_t n,
                                     const double *y0, const double *t_vals_signed, const double *tao_tab, double *traj,
                                     int32_t *hit, double *t_hit, double *y_hit, int32_t *n_rows, void *workspace,
                                     void *stream)
{
    if (!ham || !o || !workspace || n < 0) return HB_ERR_BADARG;
    if (ham->n_dof != 3 || ham->max_deg < 0 || ham->max_deg > 60) return HB_ERR_UNSUPPORTED;
    if (o->order < 2 || (o->order % 2) != 0 || o->order > 8) return HB_ERR_UNSUPPORTED;
    if (o->m < 2 || o->n_sub <= 0 || o->n_sub > HB_MAX_TAO_SUBSTEPS) return HB_ERR_BADARG;
    if (o->arith != HB_ARITH_PARITY && o->arith != HB_ARITH_FAST) return HB_ERR_BADARG;
    if (n > 0 && (!y0 || !tao_tab || (ham->ptr[6] > 0 && !ham->terms))) return HB_ERR_BADARG;
    if (ev) {
        if (ev->idx < 0 || ev->idx > 5) return HB_ERR_BADARG;
        if (n > 0 && (!t_vals_signed || !hit || !t_hit || !y_hit || !n_rows)) return HB_ERR_BADARG;
    } else if (n > 0 && !traj) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<TermHost> terms((size_t)ham->ptr[6]);
    if (ham->ptr[6]) {
        HB_CUDA_TRY(cudaMemcpyAsync(terms.data(), ham->terms, sizeof(TermHost) * terms.size(), cudaMemcpyDeviceToHost, st));
        HB_CUDA_TRY(cudaStreamSynchronize(st));
    }
    void *fn = nullptr;
    int rc = get_function(gen_grid_source(terms, ham->ptr, o->arith), "symp_grid", &fn);
    if (rc != HB_OK) return rc;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    int dev = 0, sms = 148, per_sm = 2;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_api.occ && g_api.occ(&per_sm, fn, 128, 0) != 0) per_sm = 2;
    if (per_sm < 1) per_sm = 1;
    long long blocks = (n + 127) / 128;
    const long long cap = (long long)sms * per_sm;
    if (blocks > cap) blocks = cap;
    long long nn = n;
    int m = o->m, n_sub = o->n_sub, event = ev ? 1 : 0, ev_idx = ev ? ev->idx : 0, ev_dir = ev ? ev->direction : 0;
    double ev_off = ev ? ev->offset : 0.0, xtol = ev ? ev->xtol : 0.0, gtol = ev ? ev->gtol : 0.0;
    unsigned long long *cursor = (unsigned long long *)workspace;
    void *args[] = {(void *)&y0, (void *)&nn, (void *)&m, (void *)&n_sub, (void *)&tao_tab, (void *)&t_vals_signed,
                    (void *)&traj, (void *)&event, (void *)&ev_idx, (void *)&ev_dir, (void *)&ev_off, (void *)&xtol,
                    (void *)&gtol, (void *)&hit, (void *)&t_hit, (void *)&y_hit, (void *)&n_rows, (void *)&cursor};
    const int e = g_api.launch(fn, (unsigned)blocks, 1, 1, 128, 1, 1, 0, (void *)st, args, nullptr);
    return e == 0 ? HB_OK : 1000 + e;
}

// Same contract as hb_cm_poincare_map (hb_cm.cu), specialised kernel.
extern "C" int hb_cm_poincare_map_jit(const hb_polyham *ham, const hb_cm_opts *opts, int64_t n, const double *seeds,
                                      int32_t *flags, double *out, double *t_out, void *workspace, void *stream)
{
    int rc = check_opts(ham, opts);
    if (rc != HB_OK) return rc;
    if (!workspace || n < 0) return HB_ERR_BADARG;
    if (n > 0 && (!seeds || !flags || !out || !t_out || (ham->ptr[6] > 0 && !ham->terms))) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<TermHost> terms((size_t)ham->ptr[6]);
    if (ham->ptr[6]) {
        HB_CUDA_TRY(cudaMemcpyAsync(terms.data(), ham->terms, sizeof(TermHost) * terms.size(), cudaMemcpyDeviceToHost, st));
        HB_CUDA_TRY(cudaStreamSynchronize(st));
    }
    const std::string src = gen_source(terms, ham->ptr, *opts);
    void *fn = nullptr, *fn_split = nullptr;
    rc = get_function(src, "cm_map", &fn);
    if (rc != HB_OK) return rc;
    rc = get_function(src, "cm_map_split", &fn_split);
    if (rc != HB_OK) return rc;
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    int dev = 0, sms = 148, per_sm = 2;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_api.occ && g_api.occ(&per_sm, fn, 256, 0) != 0) per_sm = 2;
    if (per_sm < 1) per_sm = 1;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (n > 2147483647LL) return HB_ERR_UNSUPPORTED;
    // cm_map_split (four warps per 32 seeds) takes the rounds whose work list is short -- and a small batch from the start
    int per_sm_split = 4;
    if (g_api.occ && g_api.occ(&per_sm_split, fn_split, 128, 0) != 0) per_sm_split = 4;
    if (per_sm_split < 1) per_sm_split = 1;
    if (per_sm_split > 4) per_sm_split = 4;
    const long long n_split = n < HB_CM_SPLIT_MAX ? n : HB_CM_SPLIT_MAX;
    long long blocks_split = (n_split + 31) / 32;
    if (blocks_split > (long long)sms * per_sm_split) blocks_split = (long long)sms * per_sm_split;
    // rounds of HB_CM_QUOTA steps; round r resumes what round r-1 parked (device-side lists and counts: no host
    // synchronisation, rounds without work return at once)
    const int rounds = opts->max_steps > 0 ? (opts->max_steps + HB_CM_QUOTA - 1) / HB_CM_QUOTA : 1;
    const size_t b_cont = sizeof(double) * 16 * (size_t)n, b_list = sizeof(int) * (size_t)n;
    const size_t b_ctr = 16 * (size_t)(rounds + 1);                  // per round: u64 cursor + int count (+pad)
    char *tmp = nullptr;
    {
        // keep the stream-ordered pool's memory across calls (the default releases it at every synchronisation and
        // the next call pays for a fresh 100+ MB allocation)
        static std::atomic<unsigned long long> pool_set{0ULL};       // one bit per device ordinal
        const unsigned long long bit = 1ULL << (dev & 63);
        if (!(pool_set.load() & bit)) {
            cudaMemPool_t pool = nullptr;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
                unsigned long long keep = ~0ULL;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_set.fetch_or(bit);
        }
    }
    HB_CUDA_TRY(cudaMallocAsync((void **)&tmp, b_cont + 2 * b_list + b_ctr, st));
    double *cont = (double *)tmp;
    int *lists[2] = {(int *)(tmp + b_cont), (int *)(tmp + b_cont + b_list)};
    char *ctr = tmp + b_cont + 2 * b_list;
    HB_CUDA_TRY(cudaMemsetAsync(ctr, 0, b_ctr, st));
    long long nn = n;
    int e = 0;
    for (int r = 0; r < rounds && e == 0; ++r) {
        unsigned long long *cursor = (unsigned long long *)(ctr + 16 * (size_t)r);
        const int *list_in = r ? lists[(r - 1) & 1] : nullptr;
        const int *count_in = r ? (const int *)(ctr + 16 * (size_t)(r - 1) + 8) : nullptr;
        int *list_out = lists[r & 1];
        int *count_out = (int *)(ctr + 16 * (size_t)r + 8);
        void *args[] = {(void *)&seeds, (void *)&nn, (void *)&flags, (void *)&out, (void *)&t_out, (void *)&cursor,
                        (void *)&list_in, (void *)&count_in, (void *)&list_out, (void *)&count_out, (void *)&cont};
        // each kernel looks at the round's count and returns at once when the other form takes it; round 0's count is n
        if (r > 0 || n > HB_CM_SPLIT_MAX) e = g_api.launch(fn, (unsigned)blocks, 1, 1, 256, 1, 1, 0, (void *)st, args, nullptr);
        if (e == 0 && (r > 0 || n <= HB_CM_SPLIT_MAX))
            e = g_api.launch(fn_split, (unsigned)blocks_split, 1, 1, 128, 1, 1, 0, (void *)st, args, nullptr);
    }
    cudaFreeAsync(tmp, st);
    return e == 0 ? HB_OK : 1000 + e;
}
