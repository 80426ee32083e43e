// Host build of csrc/glibc_trig.cuh for tests/test_glibc_trig.py: counts the arguments on
// which the restated sin/cos differ from the host libm's.  g++ -O2 -ffp-contract=off -mfma
#include "../navigation-by-deja-vu_b200/csrc/glibc_trig.cuh"

extern "C" long long nvb_trig_mismatches(const double *x, long long n, double *first_bad)
{
    long long bad = 0;
    for (long long i = 0; i < n; i++) {
        double s, c;
        nvb_glibc_sincos(x[i], &s, &c);
        const double rs = sin(x[i]), rc = cos(x[i]);
        if (memcmp(&s, &rs, 8) != 0 || memcmp(&c, &rc, 8) != 0) {
            if (bad == 0 && first_bad) *first_bad = x[i];
            bad++;
        }
    }
    return bad;
}

extern "C" void nvb_trig_eval(const double *x, long long n, double *s, double *c)
{
    for (long long i = 0; i < n; i++) nvb_glibc_sincos(x[i], s + i, c + i);
}
