// gf_math.cuh — scalar float semantics shared by every kernel of the warp.
//
// The reference CPU path is Rust: strict IEEE f32 (no FMA contraction), `as` casts that
// truncate/saturate/NaN->0, f32::round = half away from zero, and transcendental functions
// taken from the platform libm (atanf / tanf / sinf / cosf — what Rust's std calls on Linux).
// To be bit-exact with that path this header
//   * is compiled with -fmad=false (no contraction), default -prec-div/-prec-sqrt/-ftz=false;
//   * restates glibc 2.39's float algorithms operation by operation, so the device returns
//     the very bits `atanf()` etc. return on the host of this image.  gf_atanf is verified
//     against libm over all 2^32 inputs (tests/test_gf_math.py, tools/gf_math_check.cu).
// Everything here is __host__ __device__ so the same code can be checked on the CPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#  define GF_HD __host__ __device__ __forceinline__
#else
#  define GF_HD static inline
#endif

namespace gf {

GF_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; __builtin_memcpy(&f, &u, 4); return f;
#endif
}
GF_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; __builtin_memcpy(&u, &f, 4); return u;
#endif
}

// `x as i32` (Rust): truncate toward zero, saturate, NaN -> 0.  cvt.rzi.s32.f32 does exactly that.
GF_HD int32_t as_i32(float x) {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(x);
#else
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
#endif
}
// `x as usize` for f64, only ever used on values that are then clamped to < 16
GF_HD uint32_t as_usize_small(double x) {
    if (!(x > 0.0)) return 0u;            // NaN, negatives, zero
    if (x >= 4294967040.0) return 0xffffffffu;
    return (uint32_t)x;
}

// f32::round — half away from zero.  trunc(x) is exact, x - trunc(x) is exact (same binade or smaller).
GF_HD float rs_round(float x) {
    float t = truncf(x);
    float d = x - t;
    if (fabsf(d) >= 0.5f) t += copysignf(1.0f, x);
    return t;
}
// f32::max / f32::min (NaN-ignoring) == fmaxf / fminf
GF_HD float rs_max(float a, float b) { return fmaxf(a, b); }
GF_HD float rs_min(float a, float b) { return fminf(a, b); }
GF_HD float rs_clamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

// util::map_coord — src/core/util.rs:144-147; operation order is part of the contract
GF_HD float map_coord(float x, float in_min, float in_max, float out_min, float out_max) {
    return (x - in_min) * (out_max - out_min) / (in_max - in_min) + out_min;
}

// ------------------------------------------------------------------------------------------
// atanf — glibc 2.39 sysdeps/ieee754/flt-32/s_atanf.c (the fdlibm float algorithm: argument
// reduction to |x| < 7/16 around 0.5 / 1 / 1.5 / inf, then an 11-term odd/even split polynomial).
// Constants are the bit patterns glibc's decimal literals convert to (aT[0] is 0x3eaaaaab, the huge-
// argument cut-off is 2^25); with them this function equals libm's atanf on all 2^32 inputs.
// ------------------------------------------------------------------------------------------
GF_HD float gf_atanf(float x) {
    const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
    float hi, lo;
    bool reduced = true;
    if (ix >= 0x4c000000u) {                       // |x| >= 2^25
        if (ix > 0x7f800000u) return x + x;        // NaN
        const float r = u2f(0x3fc90fdau) + u2f(0x33a22168u);
        return (hx >> 31) ? -r : r;
    }
    if (ix < 0x3ee00000u) {                        // |x| < 0.4375
        if (ix < 0x31000000u) return x;            // |x| < 2^-29
        reduced = false; hi = 0.0f; lo = 0.0f;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000u) {                    // |x| < 1.1875
            if (ix < 0x3f300000u) { hi = u2f(0x3eed6338u); lo = u2f(0x31ac3769u); x = (2.0f * x - 1.0f) / (2.0f + x); }   // 7/16 <= |x| < 11/16
            else                  { hi = u2f(0x3f490fdau); lo = u2f(0x33222168u); x = (x - 1.0f) / (x + 1.0f); }          // 11/16 <= |x| < 19/16
        } else {
            if (ix < 0x401c0000u) { hi = u2f(0x3f7b985eu); lo = u2f(0x33140fb4u); x = (x - 1.5f) / (1.0f + 1.5f * x); }   // |x| < 2.4375
            else                  { hi = u2f(0x3fc90fdau); lo = u2f(0x33a22168u); x = -1.0f / x; }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (u2f(0x3eaaaaabu) + w * (u2f(0x3e124925u) + w * (u2f(0x3dba2e6eu) + w * (u2f(0x3d886b35u) + w * (u2f(0x3d4bda59u) + w * u2f(0x3c8569d7u))))));
    const float s2 = w * (u2f(0xbe4ccccdu) + w * (u2f(0xbde38e38u) + w * (u2f(0xbd9d8795u) + w * (u2f(0xbd6ef16bu) + w * u2f(0xbd15a221u)))));
    if (!reduced) return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return (hx >> 31) ? -r : r;
}

// ------------------------------------------------------------------------------------------
// sinf / cosf — glibc 2.39 sysdeps/ieee754/flt-32/{s_sinf.c,s_cosf.c,sincosf.h} as built for the
// x86-64 FMA ifunc variant (__sinf_fma/__cosf_fma, selected on every AVX2+FMA CPU): double-precision
// polynomials with the exact contraction pattern of that build (read off libm.so.6 2.39-0ubuntu8.5).
// ------------------------------------------------------------------------------------------
GF_HD double gf_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
struct SinCosTab { double c0, c1, s1, c2, s2, c3, s3, c4; };
GF_HD SinCosTab sincos_tab(bool neg) {
    SinCosTab t;
    t.c0 =  0x1.0000000000000p+0;  t.c1 = -0x1.ffffffd0c621cp-2;  t.c2 = 0x1.55553e1068f19p-5;
    t.c3 = -0x1.6c087e89a359dp-10; t.c4 =  0x1.99343027bf8c3p-16;
    t.s1 = -0x1.555545995a603p-3;  t.s2 =  0x1.1107605230bc4p-7;  t.s3 = -0x1.994eb3774cf24p-13;
    if (neg) { t.c0 = -t.c0; t.c1 = -t.c1; t.c2 = -t.c2; t.c3 = -t.c3; t.c4 = -t.c4; }
    return t;
}
// sinf_poly (sincosf.h): n even -> sine polynomial, n odd -> cosine polynomial
GF_HD float sinf_poly(double x, double x2, const SinCosTab& p, int n) {
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double s1 = gf_fma(x2, p.s3, p.s2);
        const double x7 = x3 * x2;
        const double s  = gf_fma(x3, p.s1, x);
        return (float)gf_fma(s1, x7, s);
    } else {
        const double x4 = x2 * x2;
        const double c1 = gf_fma(x2, p.c1, p.c0);
        const double c2 = gf_fma(x2, p.c4, p.c3);
        const double x6 = x4 * x2;
        const double c  = gf_fma(x4, p.c2, c1);
        return (float)gf_fma(c2, x6, c);
    }
}
// reduce_fast: |x| < 120.  hpi_inv is 2/pi * 2^24 so the quadrant lands in bits 24..31.
GF_HD double reduce_fast(double x, int& n, bool fused) {
    const double r = x * 0x1.45f306dc9c883p+23;
    n = ((int32_t)r + 0x800000) >> 24;
    return fused ? gf_fma(-(double)n, 0x1.921fb54442d18p+0, x) : x - (double)n * 0x1.921fb54442d18p+0;
}
// reduce_large: |x| >= 120, 4/pi in 32-bit chunks (__inv_pio4)
GF_HD double reduce_large(uint32_t xi, int& np) {
    const uint32_t inv_pio4[24] = {
        0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u,
        0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u,
        0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u };
    const uint32_t* arr = &inv_pio4[(xi >> 26) & 15];
    const int shift = (xi >> 23) & 7;
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    uint64_t res0 = (uint32_t)(xi * arr[0]);
    const uint64_t res1 = (uint64_t)xi * arr[4];
    const uint64_t res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    const uint64_t n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    np = (int)n;
    return (double)(int64_t)res0 * 0x1.921fb54442d18p-62;
}
GF_HD float gf_sinf(float y) {
    const uint32_t top = (f2u(y) >> 20) & 0x7ffu;
    double x = (double)y;
    int n;
    if (top < 0x3f4u) {                       // |y| < 0.75 (abstop12(pio4f))
        if (top < 0x398u) return y;           // |y| < 2^-12
        return sinf_poly(x, x * x, sincos_tab(false), 0);
    } else if (top < 0x42fu) {                // |y| < 120
        x = reduce_fast(x, n, true);
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   // sign[] = {1,-1,-1,1}
        return sinf_poly(x * s, x * x, sincos_tab((n & 2) != 0), n);
    } else if (top < 0x7f8u) {
        const uint32_t xi = f2u(y); const int sign = (int)(xi >> 31);
        x = reduce_large(xi, n);
        const int q = (n + sign) & 3;
        const double s = (q == 1 || q == 2) ? -1.0 : 1.0;
        return sinf_poly(x * s, x * x, sincos_tab(((n + sign) & 2) != 0), n);
    }
    return y - y;                             // inf / NaN -> NaN (__math_invalidf)
}
GF_HD float gf_cosf(float y) {
    const uint32_t top = (f2u(y) >> 20) & 0x7ffu;
    double x = (double)y;
    int n;
    if (top < 0x3f4u) {
        if (top < 0x398u) return 1.0f;
        return sinf_poly(x, x * x, sincos_tab(false), 1);
    } else if (top < 0x42fu) {
        x = reduce_fast(x, n, true);
        const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        return sinf_poly(x * s, x * x, sincos_tab((n & 2) != 0), n ^ 1);
    } else if (top < 0x7f8u) {
        const uint32_t xi = f2u(y); const int sign = (int)(xi >> 31);
        x = reduce_large(xi, n);
        const int q = (n + sign) & 3;
        const double s = (q == 1 || q == 2) ? -1.0 : 1.0;
        return sinf_poly(x * s, x * x, sincos_tab(((n + sign) & 2) != 0), n ^ 1);
    }
    return y - y;
}

// ------------------------------------------------------------------------------------------
// tanf — glibc 2.39 sysdeps/ieee754/flt-32/{s_tanf.c,k_tanf.c}: the fdlibm float kernel (13-term
// polynomial, pio4 - x fold above 0.6744) behind the same double-precision pi/2 reduction as sinf
// (not FMA-contracted: tanf is not an ifunc).
// ------------------------------------------------------------------------------------------
GF_HD float kernel_tanf(float x, float y, int iy) {
    const uint32_t hxu = f2u(x);
    const int32_t hx = (int32_t)hxu;
    const uint32_t ix = hxu & 0x7fffffffu;
    if (ix < 0x39000000u) {                                  // |x| < 2^-13
        if ((int)x == 0) {
            if ((ix | (uint32_t)(iy + 1)) == 0) return 1.0f / fabsf(x);
            else if (iy == 1) return x;
            else return -1.0f / x;
        }
    }
    if (ix >= 0x3f2ca140u) {                                 // |x| >= 0.6744
        if (hx < 0) { x = -x; y = -y; }
        const float z0 = u2f(0x3f490fdau) - x;
        const float w0 = u2f(0x33222168u) - y;
        x = z0 + w0; y = 0.0f;
        if (fabsf(x) < u2f(0x39000000u))
            return (float)((1 - ((hx >> 30) & 2)) * iy) * (1.0f - (float)(2 * iy) * x);
    }
    const float z = x * x;
    const float w = z * z;
    const float r0 = u2f(0x3e088889u) + w * (u2f(0x3cb327a4u) + w * (u2f(0x3b6b6916u) + w * (u2f(0x3a1a26c8u) + w * (u2f(0x38a3f445u) + w * u2f(0xb79bae5fu)))));
    const float v0 = z * (u2f(0x3d5d0dd1u) + w * (u2f(0x3c11371fu) + w * (u2f(0x3abede48u) + w * (u2f(0x398137b9u) + w * (u2f(0x3895c07au) + w * u2f(0x37d95384u))))));
    const float s = z * x;
    float r = y + z * (s * (r0 + v0) + y);
    r += u2f(0x3eaaaaabu) * s;
    const float ww = x + r;
    if (ix >= 0x3f2ca140u) {
        const float v = (float)iy;
        return (float)(1 - ((hx >> 30) & 2)) * (v - 2.0f * (x - (ww * ww / (ww + v) - r)));
    }
    if (iy == 1) return ww;
    // -1/(x+r) with extra care
    const float zz = u2f(f2u(ww) & 0xfffff000u);
    const float vv = r - (zz - x);
    const float a = -1.0f / ww;
    const float t = u2f(f2u(a) & 0xfffff000u);
    const float ss = 1.0f + t * zz;
    return t + a * (ss + t * vv);
}
GF_HD float gf_tanf(float x) {
    const uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
    if (ix <= 0x3f490fdau) return kernel_tanf(x, 0.0f, 1);
    if (ix >= 0x7f800000u) return x - x;                     // inf / NaN
    int n; double dx = (double)x;
    if (((hx >> 20) & 0x7ffu) < 0x42fu) dx = reduce_fast(dx, n, false);
    else { dx = reduce_large(hx, n); if (hx >> 31) dx = -dx; }
    const float y0 = (float)dx;
    const float y1 = (float)(dx - (double)y0);
    return kernel_tanf(y0, y1, 1 - ((n & 1) << 1));
}

} // namespace gf
