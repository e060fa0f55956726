/* Host check of hb_x87.cuh against real x87 long double arithmetic.
 * Build: g++ -O2 -o /tmp/check_x87 -x c++ tools/check_x87.c -lm */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "../hiten_b200/csrc/hb_x87.cuh"

int main(void)
{
    uint64_t s = 88172645463325252ULL;
    long bad = 0, n = 0;
    for (long it = 0; it < 3000000; ++it) {
        double v[6];
        const int len = 1 + (int)(it % 6);
        for (int i = 0; i < len; ++i) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            const double u = (s >> 11) * (1.0 / 9007199254740992.0);
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            const double e = -20.0 + 36.0 * ((s >> 11) * (1.0 / 9007199254740992.0));
            v[i] = (u - 0.5) * pow(10.0, e);
            if (it % 97 == 0 && i == 0) v[i] = 0.0;
        }
        long double acc = 0.0L;
        for (int i = 0; i < len; ++i) acc += (long double)v[i] * (long double)v[i];
        const double want = (double)sqrtl(acc);
        const double got = hb_x87_norm2(v, len);
        ++n;
        if (memcmp(&want, &got, 8) != 0) { if (bad < 5) printf("mismatch len %d: got %a want %a\n", len, got, want); ++bad; }
    }
    printf("checked %ld, mismatches %ld\n", n, bad);
    return bad != 0;
}
