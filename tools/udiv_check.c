/* udiv_check.c — exhaustive CPU check of the uniform-divisor division used by map_coord on the device
 * (div_uniform in gyroflow_b200/csrc/warp_kernel.cuh, map_apply_x2 in warp_kernel_x2.cuh):
 *
 *     r  = RN(1 / d)              (host, once per frame)
 *     q0 = RN(a * r)
 *     r0 = fma(-d, q0, a)         (exact: q0 is within one ulp of a / d)
 *     q  = RN(q0 + r0 * r)  ==  RN(a / d)
 *
 * This is Markstein's division theorem (P. Markstein, IBM J. R&D 34(1), 1990; Muller et al., "Handbook of Floating-Point
 * Arithmetic", 2nd ed., Thm. 4.8): with y the CORRECTLY ROUNDED reciprocal of b and q a faithful approximation of a / b,
 * RN(q + RN(a - b q) y) is the correctly rounded quotient, provided no intermediate under/overflows.  The window the
 * kernels use (2^-80 < |a| < 2^60, 2^-40 <= |d| <= 2^40) keeps every intermediate normal.
 *
 * The tool confirms it by brute force: for each divisor on the command line (default: the frame sizes and odd values the
 * tests use) every float a in the window is compared with the IEEE quotient.  Build: gcc -O2 -fopenmp -ffp-contract=off
 * tools/udiv_check.c -o /tmp/udiv_check -lm;  run: /tmp/udiv_check [d ...].  Exit code 0 = no mismatch.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static unsigned long long check_divisor(float d, int stride) {
    const float r = 1.0f / d;
    unsigned long long bad = 0;
    /* biased exponents 47 (2^-80) .. 186 (2^59): every mantissa, both signs */
    #pragma omp parallel for reduction(+ : bad) schedule(dynamic, 1)
    for (int e = 47; e <= 186; ++e) {
        for (uint32_t m = 0; m < (1u << 23); m += (uint32_t)stride) {
            for (uint32_t s = 0; s < 2; ++s) {
                const float a = u2f((s << 31) | ((uint32_t)e << 23) | m);
                if (!(fabsf(a) > 0x1p-80f && fabsf(a) < 0x1p60f)) continue;
                const float q0 = a * r;
                const float r0 = fmaf(-d, q0, a);
                const float q = fmaf(r0, r, q0);
                const float want = a / d;
                uint32_t qa, qb; memcpy(&qa, &q, 4); memcpy(&qb, &want, 4);
                if (qa != qb) ++bad;
            }
        }
    }
    return bad;
}

int main(int argc, char** argv) {
    static const float defaults[] = { 3840.0f, 2160.0f, 1920.0f, 1080.0f, 7680.0f, 4320.0f, 640.0f, 360.0f, 1280.0f, 720.0f,
                                      203.0f, 117.0f, 3.0f, 7.0f, 1000.0f, 0.3333333f, 1e-9f, 8.5e11f, -49.0f, 16383.0f };
    int stride = 1;
    unsigned long long total = 0;
    int n = 0;
    for (int i = 1; i < argc; ++i) {
        if (!strncmp(argv[i], "--stride=", 9)) { stride = atoi(argv[i] + 9); if (stride < 1) stride = 1; continue; }
        const float d = strtof(argv[i], NULL);
        const unsigned long long b = check_divisor(d, stride);
        printf("d = %-14.9g mismatches %llu\n", (double)d, b);
        total += b; ++n;
    }
    if (n == 0) {
        for (size_t i = 0; i < sizeof(defaults) / sizeof(defaults[0]); ++i) {
            const unsigned long long b = check_divisor(defaults[i], stride);
            printf("d = %-14.9g mismatches %llu\n", (double)defaults[i], b);
            total += b;
        }
    }
    printf("total mismatches %llu (stride %d)\n", total, stride);
    return total ? 1 : 0;
}
