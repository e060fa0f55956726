// hb_x87.cuh -- bit-exact emulation of the x87 extended-precision 2-norm the reference gets from
// np.linalg.norm (OpenBLAS dnrm2 on x86-64 accumulates squares in 80-bit registers, takes fsqrt there and
// rounds to double once; measured equal to `s = 0L; s += (long double)v*v; (double)sqrtl(s)` on 2000/2000
// random 6-vectors).  It decides the initial step size (rk.py:2445-2448) and RK45's error norm (rk.py:1333),
// so reproducing it bit for bit is part of reproducing the reference's step sequence.
//
// Every x87 operation rounds to a 64-bit significand (round to nearest even); values here are positive,
// normal and far from the exponent limits, so an (exponent, 64-bit significand) pair with 128-bit integer
// intermediates is enough.  Plain C++ so that tools/check_x87.c can compile the same code on the host.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define HB_HD __host__ __device__ inline
#define HB_SQRT_D(x) sqrt(x)
#else
#define HB_HD static inline
#define HB_SQRT_D(x) __builtin_sqrt(x)
#endif

typedef unsigned __int128 hb_u128;

struct hb_ext {           // value = m * 2^(e - 63), m in [2^63, 2^64), or m == 0
    uint64_t m;
    int e;
};

HB_HD uint64_t hb_bits(double x)
{
    union { double d; uint64_t u; } c;
    c.d = x;
    return c.u;
}
HB_HD double hb_from_bits(uint64_t u)
{
    union { double d; uint64_t u; } c;
    c.u = u;
    return c.d;
}

// round a 128-bit significand (top bit anywhere) to 64 bits, nearest-even; `sticky` = bits already lost
HB_HD hb_ext hb_ext_round(hb_u128 v, int e_of_bit127, int sticky)
{
    hb_ext r;
    if (v == 0) { r.m = 0; r.e = 0; return r; }
    int lz = 0;
    uint64_t hi = (uint64_t)(v >> 64);
    if (hi == 0) { lz = 64; uint64_t lo = (uint64_t)v; while (!(lo >> 63)) { lo <<= 1; ++lz; } }
    else { while (!(hi >> 63)) { hi <<= 1; ++lz; } }
    v <<= lz;
    int e = e_of_bit127 - lz;
    uint64_t m = (uint64_t)(v >> 64);
    const uint64_t low = (uint64_t)v;
    const int round_bit = (int)(low >> 63);
    const int rest = ((low << 1) != 0) || sticky;
    if (round_bit && (rest || (m & 1))) {
        ++m;
        if (m == 0) { m = 0x8000000000000000ULL; ++e; }
    }
    r.m = m; r.e = e;
    return r;
}

// (long double)a * (long double)b for finite non-zero doubles a, b (sign dropped: squares / positive operands)
HB_HD hb_ext hb_ext_mul_dd(double a, double b)
{
    const uint64_t ua = hb_bits(a) & 0x7fffffffffffffffULL, ub = hb_bits(b) & 0x7fffffffffffffffULL;
    hb_ext z; z.m = 0; z.e = 0;
    if (ua == 0 || ub == 0) return z;
    int ea = (int)(ua >> 52), eb = (int)(ub >> 52);
    uint64_t ma = ua & 0xfffffffffffffULL, mb = ub & 0xfffffffffffffULL;
    if (ea == 0) { ea = 1; while (!(ma >> 52)) { ma <<= 1; --ea; } } else ma |= 1ULL << 52;
    if (eb == 0) { eb = 1; while (!(mb >> 52)) { mb <<= 1; --eb; } } else mb |= 1ULL << 52;
    const hb_u128 p = (hb_u128)ma * mb;                       // < 2^106, value = p * 2^(ea+eb-2*1075)
    // bit 127 of p would have weight 2^(ea+eb-2150+127)
    return hb_ext_round(p, ea + eb - 2150 + 127, 0);
}

HB_HD hb_ext hb_ext_add(hb_ext a, hb_ext b)                    // both >= 0
{
    if (a.m == 0) return b;
    if (b.m == 0) return a;
    if (a.e < b.e) { hb_ext t = a; a = b; b = t; }
    const int d = a.e - b.e;
    hb_u128 va = (hb_u128)a.m << 63;                          // bit 126 = leading bit, weight 2^(a.e)
    hb_u128 vb = (hb_u128)b.m << 63;
    int sticky = 0;
    if (d >= 127) { sticky = 1; vb = 0; }
    else if (d > 0) { sticky = (vb & (((hb_u128)1 << d) - 1)) != 0; vb >>= d; }
    const hb_u128 s = va + vb;                                // < 2^128
    return hb_ext_round(s, a.e + 1, sticky);
}

HB_HD hb_ext hb_ext_sqrt(hb_ext a)
{
    hb_ext r; r.m = 0; r.e = 0;
    if (a.m == 0) return r;
    // a = m * 2^(e-63); make the exponent of the 128-bit radicand even: M = m << s, sqrt(a) = sqrt(M) * 2^((e-63-s)/2)
    int s = ((a.e - 63) & 1) ? 63 : 64;                       // e-63-s even
    const hb_u128 M = (hb_u128)a.m << s;                      // in [2^126, 2^128)
    // floor(sqrt(M)) by Newton from a double estimate
    uint64_t q = (uint64_t)HB_SQRT_D((double)a.m * (s == 64 ? 18446744073709551616.0 : 9223372036854775808.0));
    if (q == 0) q = 1;
    for (int it = 0; it < 4; ++it) {
        const hb_u128 qn = ((hb_u128)q + M / q) >> 1;
        q = (qn >> 64) ? 0xffffffffffffffffULL : (uint64_t)qn;
    }
    while ((hb_u128)q * q > M) --q;
    while (((hb_u128)q + 1) * ((hb_u128)q + 1) <= M) { if (q == 0xffffffffffffffffULL) break; ++q; }
    const hb_u128 rem = M - (hb_u128)q * q;
    // q has 64 bits (M >= 2^126 -> q >= 2^63).  Round to nearest: (q + 1/2)^2 = q^2 + q + 1/4
    uint64_t m = q;
    int e = (a.e - 63 - s) / 2 + 63;
    if (rem > (hb_u128)q) { ++m; if (m == 0) { m = 0x8000000000000000ULL; ++e; } }
    r.m = m; r.e = e;
    return r;
}

HB_HD double hb_ext_to_double(hb_ext a)                        // FST: 64 -> 53 bits, nearest-even (normal range)
{
    if (a.m == 0) return 0.0;
    uint64_t m = a.m >> 11;
    const uint64_t low = a.m & 0x7ff;
    int e = a.e;
    if (low > 0x400 || (low == 0x400 && (m & 1))) {
        ++m;
        if (m >> 53) { m >>= 1; ++e; }
    }
    const int be = e + 1023;
    if (be <= 0 || be >= 2047) return (be <= 0) ? 0.0 : hb_from_bits(0x7ff0000000000000ULL);
    return hb_from_bits(((uint64_t)be << 52) | (m & 0xfffffffffffffULL));
}

// np.linalg.norm(v) for a float64 vector of length n (n <= 64), as computed by the reference platform.
// Out of line on the device: it runs once per trajectory, and inlining its 128-bit integer code into the
// persistent integration loops cost them 20-35 % (instruction-cache footprint).
#ifdef __CUDACC__
static __host__ __device__ __noinline__ double hb_x87_norm2(const double *v, int n)
#else
static inline double hb_x87_norm2(const double *v, int n)
#endif
{
    hb_ext s; s.m = 0; s.e = 0;
    for (int i = 0; i < n; ++i) s = hb_ext_add(s, hb_ext_mul_dd(v[i], v[i]));
    return hb_ext_to_double(hb_ext_sqrt(s));
}
