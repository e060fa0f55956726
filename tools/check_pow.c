/* Host check of the pow() restatement used on the GPU: same tables, same FMA sequence, compared with libm pow()
 * bit for bit.  Build: gcc -O2 -ffp-contract=off -mfma -o /tmp/check_pow tools/check_pow.c -lm
 * (tables included from the generated header with the device qualifiers stripped). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define __device__
#define constexpr const
#include "../hiten_b200/csrc/hb_libm_pow_tables.h"

static inline uint64_t asu(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double asd(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* valid for normal x > 0 and |y*log(x)| in [2^-54, 512); returns 0 when outside (caller falls back) */
static int pow_libm(double x, double y, double *res)
{
    const uint64_t ix = asu(x);
    const uint32_t topx = ix >> 52, topy = (asu(y) >> 52) & 0x7ff;
    if (topx - 1 > 0x7fd || topy - 0x3be > 0x7f) return 0;
    const uint64_t tmp = ix - 0x3fe6955500000000ULL;
    const int i = (tmp >> 45) & 0x7f;
    const int k = (int)((int64_t)tmp >> 52);
    const double z = asd(ix - (tmp & 0xfff0000000000000ULL));
    const double kd = (double)k;
    const double t1 = fma(kd, HB_POW_LN2HI, HB_POW_LOGC[i]);
    const double lo1 = fma(kd, HB_POW_LN2LO, HB_POW_LOGCTAIL[i]);
    const double r = fma(z, HB_POW_INVC[i], -1.0);
    const double ar = r * HB_POW_A[0];
    const double q12 = fma(r, HB_POW_A[2], HB_POW_A[1]);
    const double q34 = fma(r, HB_POW_A[4], HB_POW_A[3]);
    const double t2 = r + t1;
    const double lo2 = (t1 - t2) + r;
    const double ar2 = r * ar;
    const double ar3 = r * ar2;
    const double lo3 = fma(ar, r, -ar2);
    const double hi = t2 + ar2;
    const double q56 = fma(r, HB_POW_A[6], HB_POW_A[5]);
    const double lo4 = (t2 - hi) + ar2;
    const double q = fma(ar2, fma(q56, ar2, q34), q12);
    const double lo = fma(ar3, q, ((lo1 + lo2) + lo3) + lo4);
    const double lhi = hi + lo;
    const double ltail = (hi - lhi) + lo;
    const double ehi = y * lhi;
    const double elo = fma(y, ltail, fma(lhi, y, -ehi));
    const uint32_t abstop = (asu(ehi) >> 52) & 0x7ff;
    if (abstop - 0x3c9 > 0x3e) return 0;
    const double kds = fma(ehi, HB_EXP_INVLN2N, HB_EXP_SHIFT);
    const uint64_t ki = asu(kds);
    const double kd2 = kds - HB_EXP_SHIFT;
    double rr = fma(kd2, HB_EXP_NEGLN2LON, fma(kd2, HB_EXP_NEGLN2HIN, ehi));
    rr = elo + rr;
    const unsigned idx = 2 * (ki & 0x7f);
    const uint64_t sbits = HB_EXP_T[idx + 1] + (ki << 45);
    const double tail = asd(HB_EXP_T[idx]);
    const double r2 = rr * rr;
    const double p23 = fma(rr, HB_EXP_C[1], HB_EXP_C[0]);
    const double p45 = fma(rr, HB_EXP_C[3], HB_EXP_C[2]);
    const double tmpv = fma(p45, r2 * r2, fma(p23, r2, rr + tail));
    const double scale = asd(sbits);
    *res = fma(tmpv, scale, scale);
    return 1;
}

int main(void)
{
    const double ys[6] = { -1.0 / 9.0, 0.4 * (1.0 / 9.0), -1.0 / 8.0, -1.0 / 6.0, 0.4 * (1.0 / 6.0), -1.0 / 5.0 };
    uint64_t s = 88172645463325252ULL;
    long n = 0, bad = 0, skipped = 0;
    for (long it = 0; it < 20000000; ++it) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double u01 = (s >> 11) * (1.0 / 9007199254740992.0);
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const double x = pow(10.0, -12.0 + 18.0 * u01) * (1.0 + (s >> 11) * (1.0 / 9007199254740992.0));
        const double y = (it % 7 == 6) ? (-3.0 + 6.0 * u01) : ys[it % 6];
        double r;
        if (!pow_libm(x, y, &r)) { ++skipped; continue; }
        ++n;
        if (asu(r) != asu(pow(x, y))) { if (bad < 5) printf("mismatch x=%a y=%a got %a want %a\n", x, y, r, pow(x, y)); ++bad; }
    }
    printf("checked %ld, skipped %ld, mismatches %ld\n", n, skipped, bad);
    return bad != 0;
}
