// hb_cr3bp_common.cuh -- pieces shared by the 6-state DOP853 kernels (hb_cr3bp.cu, hb_cr3bp_section.cu):
// kernel parameters, the CR3BP vector field, initial-step heuristic, host-side parameter filling.
#pragma once
#include "hb_dop853.cuh"
#include "hb_section.cuh"

namespace hbc {




// 256 resident threads per SM (8 warps) is what 240+ registers per thread allow; one 256-thread CTA
// per SM measured ~8% faster than two of 128 (gpurun probe, round 1).
#ifndef HB_BLOCK
#define HB_BLOCK 256
#endif
#ifndef HB_MINBLOCKS
#define HB_MINBLOCKS 1
#endif

struct PropParams {
    double mu, om;
    unsigned negmask;  // bit d set -> derivative d negated (fwd == -1 wrapper)
    double rtol, atol, max_step, min_step;
    long long max_attempts;
    long long n;
    const double *y0;
    double t0, tf;
    const double *tf_arr;
    double *yf;
    int *nacc, *nrej, *status;
    HbWorkspace *ws;
    const double *t_eval;
    int m;
    double *dense_out;
    int ev_idx, ev_dir;
    double ev_off, xtol, gtol;
    double *t_hit;
    HitSink sink;          // MODE_SECTION: detector settings + hit buffer
    int *hits_per_traj;
    double tsign;          // sign applied to grid times for the detector (times = forward * t_eval)
    double inv_grid_dt;    // (m-1)/(t_eval[m-1]-t_eval[0]): first guess when locating grid samples
    const double *h0;      // first step sizes from k_first_steps (null: computed inline when a trajectory starts)
    double *rec;           // MODE_RECORD / MODE_RECORD_NEAR: per-step stage records [n][rec_cap][HB_REC_DOUBLES]
    int rec_cap;
    int *nrec;             // MODE_RECORD_NEAR: records written per trajectory (only steps near the section plane + neighbours)
    int max_ctas;          // host side only: cap on the persistent launch's grid (hb_integ.max_ctas; 0 = every SM)
    const int *order;      // hb_integ.order: trajectory handed out q-th (null: q itself)
};

// One accepted step as stored by MODE_RECORD (hb_cr3bp.cu) and consumed by the scan kernels (hb_section_scan.cu):
//   header (3 sectors): [0] t_old [1] t_new [2] hseg [3] y_old[sidx] [4..10] F[0..6][sidx]   (event component)
//   body              : [11..16] y_old  [17..58] F[7][6] row-major  [59] pad
// MODE_RECORD step record (512 B = 16 x 32-byte sectors, written/read with 256-bit accesses): everything the dense
// output of one accepted step is a function of.  Stages 2..5 do not enter it (the extra-stage rows and D have zero
// columns 1..4) and k1 = f(y_old) is recomputed by the consumer.
#define HB_REC_DOUBLES 64
#define HB_REC_T 0        // t_old, t_new
#define HB_REC_YOLD 2
#define HB_REC_YNEW 8
#define HB_REC_K5 14      // k[5..11] (7 x 6)
#define HB_REC_K12 56     // k[12] = f(y_new)
#define HB_REC_META 62    // sparse records: int32 step number (0-based), int32 flags (HB_REC_LAST); [63] unused
#define HB_REC_LAST 1     // the trajectory's last accepted step (owns every grid sample that is left)
#define HB_REC_ROW_BYTES (HB_REC_DOUBLES * 8 + 16)   // shared-memory staging row, padded: conflict-free 16 B accesses

HB_DEV double hb_flip_sign(double x)      // -x, bit for bit (also for zeros and NaNs), without the FP64 pipe
{
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}

HB_DEV void hb_st4(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
HB_DEV void hb_ld4(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// ---------------------------------------------------------------------------------------------
// Vector field.  Parity form keeps rtbp.py:65-74's operation order:
//   r1 = sqrt((x+mu)**2 + y**2 + z**2);  r**3 -> r*(r*r)
//   ax = 2*vy + x - (1-mu)*(x+mu)/r1**3 - mu*(x-1+mu)/r2**3   (left to right)
// Fast form: rsqrt-based, 2 MUFU + Newton instead of 2 sqrt + 6 div.
// ---------------------------------------------------------------------------------------------
// NEG: 0 = forward (no sign change), 1 = every derivative negated (fwd = -1, flip all, the manifold
// case base.py:296-300), 2 = generic per-component mask.
template <class AR, int NEG>
HB_DEV void crtbp_rhs(const double (&s)[6], const PropParams &p, double (&out)[6])
{
    const double x = s[0], y = s[1], z = s[2], vx = s[3], vy = s[4], vz = s[5];
    const double mu = p.mu, om = p.om;
    double ax, ay, az;
    if constexpr (AR::parity) {
        const double xm = AR::add(x, mu);
        const double xo = AR::sub(x, om);
        const double yy = AR::mul(y, y), zz = AR::mul(z, z);
        const double r1 = AR::sqrt(AR::add(AR::add(AR::mul(xm, xm), yy), zz));
        const double r2 = AR::sqrt(AR::add(AR::add(AR::mul(xo, xo), yy), zz));
        const double r1c = AR::mul(r1, AR::mul(r1, r1));
        const double r2c = AR::mul(r2, AR::mul(r2, r2));
        // three quotients per denominator share one refined reciprocal (each stays correctly rounded)
        const double i1 = hb_rcp_refined(r1c), i2 = hb_rcp_refined(r2c);
        const double xq = AR::add(AR::sub(x, 1.0), mu);  // (x - 1 + mu)
        ax = AR::sub(AR::sub(AR::add(AR::mul(2.0, vy), x), hb_div_with(AR::mul(om, xm), r1c, i1)),
                     hb_div_with(AR::mul(mu, xq), r2c, i2));
        ay = AR::sub(AR::sub(AR::add(AR::mul(-2.0, vx), y), hb_div_with(AR::mul(om, y), r1c, i1)),
                     hb_div_with(AR::mul(mu, y), r2c, i2));
        az = AR::sub(hb_div_with(AR::mul(-om, z), r1c, i1), hb_div_with(AR::mul(mu, z), r2c, i2));
    } else {
        const double xm = x + mu;
        const double xo = x - om;
        const double yz = fma(y, y, z * z);
        const double i1 = hb_rsqrt_fast(fma(xm, xm, yz));
        const double i2 = hb_rsqrt_fast(fma(xo, xo, yz));
        const double c1 = om * (i1 * i1 * i1);
        const double c2 = mu * (i2 * i2 * i2);
        const double cs = c1 + c2;
        ax = fma(2.0, vy, x) - fma(c1, xm, c2 * xo);
        ay = fma(-2.0, vx, y) - cs * y;
        az = -cs * z;
    }
    if constexpr (NEG == 0) {
        out[0] = vx; out[1] = vy; out[2] = vz; out[3] = ax; out[4] = ay; out[5] = az;
    } else if constexpr (NEG == 1) {
        if constexpr (AR::parity) {
            // sign flips on the integer pipe: the compiler otherwise spends ~50 DADD (-0 - x) per step on the FP64 pipe,
            // which bounds this kernel, wherever it cannot fold the sign into a consumer
            out[0] = hb_flip_sign(vx); out[1] = hb_flip_sign(vy); out[2] = hb_flip_sign(vz);
            out[3] = hb_flip_sign(ax); out[4] = hb_flip_sign(ay); out[5] = hb_flip_sign(az);
        } else {
            out[0] = -vx; out[1] = -vy; out[2] = -vz; out[3] = -ax; out[4] = -ay; out[5] = -az;
        }
    } else {
        out[0] = (p.negmask & 1u) ? -vx : vx;
        out[1] = (p.negmask & 2u) ? -vy : vy;
        out[2] = (p.negmask & 4u) ? -vz : vz;
        out[3] = (p.negmask & 8u) ? -ax : ax;
        out[4] = (p.negmask & 16u) ? -ay : ay;
        out[5] = (p.negmask & 32u) ? -az : az;
    }
}

template <class AR, int NEG>
struct Cr3bpRhs {
    const PropParams &p;
    HB_DEV void operator()(const double (&y)[6], double (&dy)[6]) const { crtbp_rhs<AR, NEG>(y, p, dy); }
};

// Component select without dynamic register-array indexing (which would force local memory).
HB_DEV double pick6(const double (&v)[6], int i)
{
    double r = v[0];
#pragma unroll
    for (int d = 1; d < 6; ++d) r = (i == d) ? v[d] : r;
    return r;
}

// scale0, d0, d1, h0 (rk.py:2445-2448; utils.py:127-157).  The reference's np.linalg.norm is OpenBLAS dnrm2
// (x87 extended accumulation), emulated bit for bit by hb_x87_norm2.
HB_DEV double norm2_ext6(const double (&v)[6])
{
    double w[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) w[d] = v[d];
    return hb_x87_norm2(w, 6);       // bit-exact x87 dnrm2 (hb_x87.cuh)
}

template <class AR>
HB_DEV double initial_step(const double (&y)[6], const double (&f)[6], const PropParams &p)
{
    double a[6], b[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        const double sc = AR::madd(p.rtol, fabs(y[d]), p.atol);
        a[d] = AR::div(y[d], sc);
        b[d] = AR::div(f[d], sc);
    }
    const double sq = AR::sqrt(6.0);
    double d0, d1;
    if constexpr (AR::parity) {
        d0 = AR::div(norm2_ext6(a), sq);
        d1 = AR::div(norm2_ext6(b), sq);
    } else {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int d = 0; d < 6; ++d) { s0 = fma(a[d], a[d], s0); s1 = fma(b[d], b[d], s1); }
        d0 = ::sqrt(s0) / sq;
        d1 = ::sqrt(s1) / sq;
    }
    double h = (d0 < 1.0e-5 || d1 < 1.0e-5) ? 1.0e-6 : AR::div(AR::mul(0.01, d0), d1);
    if (h > p.max_step) h = p.max_step;
    if (h < p.min_step) h = p.min_step;
    return h;
}


// First step sizes of all trajectories in one convergent pass.  Inside the persistent kernels a trajectory start
// is executed by whichever lanes just ran out of work while the rest of the warp waits; in the parity build the
// start (12 divisions + two emulated x87 norms) costs about half a step, ~15 % of the warp's time.  The values
// are parked in the FIRST ROW of the output array yf, the accelerations of f(y0) in rows 3..5 (each thread reads its
// entries before it writes that trajectory's end state), so no scratch is needed; callers whose yf overlaps y0 keep
// the inline path.
template <class AR, int NEG>
__global__ void __launch_bounds__(128) k_first_steps(const PropParams p, double *h0)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.n) return;
    double y[6], f[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) y[d] = p.y0[(long long)d * p.n + idx];
    crtbp_rhs<AR, NEG>(y, p, f);
    h0[idx] = initial_step<AR>(y, f, p);
    // the accelerations of f(y0) ride along in rows 3..5 (k_dop853_6 then starts a trajectory with loads only: inside the
    // persistent kernel the vector field of a refill runs at ~2 of 32 lanes); f[0..2] are the (signed) velocities
#pragma unroll
    for (int d = 3; d < 6; ++d) h0[(long long)d * p.n + idx] = f[d];
}

template <class AR>
inline int first_steps_prepass(PropParams &p, cudaStream_t st)
{
    p.h0 = nullptr;
    if (!p.yf || p.n < 128) return HB_OK;                                   // tiny batches: not worth a launch
    const double *a0 = p.y0, *a1 = p.y0 + 6 * p.n, *b0 = p.yf, *b1 = p.yf + 6 * p.n;
    if (b0 < a1 && a0 < b1) return HB_OK;                                   // in-place call: yf rows are live input
    const unsigned grid = (unsigned)((p.n + 127) / 128);
    if (p.negmask == 0u) k_first_steps<AR, 0><<<grid, 128, 0, st>>>(p, p.yf);
    else if (p.negmask == 63u) k_first_steps<AR, 1><<<grid, 128, 0, st>>>(p, p.yf);
    else k_first_steps<AR, 2><<<grid, 128, 0, st>>>(p, p.yf);
    HB_CUDA_TRY(cudaGetLastError());
    p.h0 = p.yf;
    return HB_OK;
}

inline int fill_params(const hb_cr3bp *sys, const hb_integ *integ, PropParams &p)
{
    if (!sys || !integ) return HB_ERR_BADARG;
    if (integ->method != HB_DOP853 && integ->method != HB_RK45 && integ->method != HB_RK4 &&
        integ->method != HB_RK6 && integ->method != HB_RK8)
        return HB_ERR_UNSUPPORTED;
    if (integ->arith != HB_ARITH_PARITY && integ->arith != HB_ARITH_FAST) return HB_ERR_BADARG;
    p.mu = sys->mu;
    p.om = 1.0 - sys->mu;
    p.negmask = 0;
    if (sys->fwd < 0) {
        int lo = sys->flip_lo, hi = sys->flip_hi;
        if (lo < 0) { lo = 0; hi = 6; }
        if (hi > 6 || lo > hi) return HB_ERR_BADARG;
        for (int d = lo; d < hi; ++d) p.negmask |= 1u << d;
    }
    p.rtol = integ->rtol; p.atol = integ->atol;
    p.max_step = integ->max_step; p.min_step = integ->min_step;
    p.max_attempts = integ->max_attempts > 0 ? integ->max_attempts : 2147483647LL;
    p.max_ctas = integ->max_ctas > 0 ? integ->max_ctas : 0;
    p.order = integ->order;
    return HB_OK;
}

inline int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}


}  // namespace hbc

// hb_cr3bp_rk.cu: RK45 and fixed-step RK4/6/8 (mode 0 end state, 1 dense grid, 2 terminal event)
int hb_rk_dispatch(const hbc::PropParams &p, int method, int arith, int mode, int n_fixed, cudaStream_t st);
