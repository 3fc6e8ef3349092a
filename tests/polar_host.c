/* Host restatement of shb_polar (csrc/shb_kernels.cu) for tests/test_polar_host.py: the same operations with libm's fma();
 * the MUFU.RSQ64H seed of the device is modelled by a float reciprocal square root cut to 20 bits (the cubic step that
 * follows takes either to full double precision).  Table: tab.inc, extracted from the .cu by the test.
 *   polar_check(n, seed) -> max |theta - atan2l| in ulp of the result; *sqrt_bad = samples whose r is not sqrt(x*x + y*y). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static const double tab[23][4] = {
#include "tab.inc"
};
static double seed_rsqrt(double a) { float f = 1.0f / sqrtf((float)a); uint32_t u; memcpy(&u, &f, 4); u &= 0xfffffff8u; memcpy(&f, &u, 4); return (double)f; }
static void polar(double x, double y, double* th, double* r) {
    const double s2 = x * x + y * y;
    const double y0 = seed_rsqrt(s2);
    const double e = fma(s2, -(y0 * y0), 1.0);
    const double y1 = fma(fma(e, 0.375, 0.5), y0 * e, y0);
    const double g = s2 * y1;
    *r = fma(fma(g, -g, s2), 0.5 * y1, g);
    const double c = fabs(x) * y1, s = fabs(y) * y1;
    const int swap = s > c;
    const double a = swap ? c : s, b = swap ? s : c;
    unsigned i = (unsigned)(a * 32.0);
    if (i > 22) i = 22;
    const double u = fma(a, tab[i][1], -(b * tab[i][0]));
    const double w = u * u;
    const double p = fma(w, fma(w, fma(w, 35.0 / 1152.0, 5.0 / 112.0), 3.0 / 40.0), 1.0 / 6.0);
    double at = tab[i][2] + fma(u * w, p, u);
    if (swap) at = 1.57079632679489655800e+00 - (at - 6.12323399573676603587e-17);
    if (x < 0.0) at = 3.1415926535897931160e+00 - (at - 1.2246467991473531772e-16);
    *th = copysign(at, y);
}
double polar_check(long n, long seed, long* sqrt_bad) {
    srand48(seed);
    double maxulp = 0.0;
    *sqrt_bad = 0;
    for (long k = 0; k < n; ++k) {
        double x, y;
        switch (k % 4) {
            case 0: x = (drand48() - 0.5) * 100; y = (drand48() - 0.5) * 100; break;                                   /* a bone section's box */
            case 1: { const double t = (drand48() - 0.5) * 6.283185307179586, R = exp((drand48() - 0.5) * 20); x = R * cos(t); y = R * sin(t); } break;
            case 2: x = (drand48() - 0.5) * 100; y = x * (drand48() < 0.5 ? 1 : -1) * (1 + (drand48() - 0.5) * 1e-6); break;   /* the diagonals */
            default: x = (drand48() - 0.5) * 100; y = (drand48() - 0.5) * 1e-8 * x; if (drand48() < 0.5) { const double t = x; x = y; y = t; } break;   /* the axes */
        }
        double th, r;
        polar(x, y, &th, &r);
        const long double ref = atan2l((long double)y, (long double)x);
        const double d = fabs((double)((long double)th - ref));
        const double ulp = nextafter(fabs((double)ref), INFINITY) - fabs((double)ref);
        if (d / ulp > maxulp) maxulp = d / ulp;
        if (r != sqrt(x * x + y * y)) ++*sqrt_bad;
    }
    return maxulp;
}
