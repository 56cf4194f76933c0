/* gf_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see gf_oracle.h for the rules).
 *
 * Restates, function by function, the reference CPU path.  Citations are relative to
 * /root/reference/src/core (gyroflow @ b5e8828).  PARITY UNPINNED — see gf_oracle.h.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 */
#define _GNU_SOURCE
#include "gf_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * Rust scalar semantics
 * ---------------------------------------------------------------------------------------- */
/* `x as i32` for f32: truncate toward zero, saturate, NaN -> 0 */
static inline int32_t rs_f32_as_i32(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}
/* `x as usize` for f64 */
static inline size_t rs_f64_as_usize(double x) {
    if (x != x) return 0;
    if (x <= 0.0) return 0;
    if (x >= 18446744073709551616.0) return SIZE_MAX;
    return (size_t)x;
}
static inline uint8_t rs_f32_as_u8(float x) {
    if (x != x) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 255.0f) return 255;
    return (uint8_t)x;
}
static inline uint16_t rs_f32_as_u16(float x) {
    if (x != x) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 65535.0f) return 65535;
    return (uint16_t)x;
}
/* f32::round — half away from zero (glibc roundf is exact) */
static inline float rs_round(float x) { return roundf(x); }
/* f32::max / f32::min — NaN-ignoring, like fmaxf/fminf */
static inline float rs_max(float a, float b) { return fmaxf(a, b); }
static inline float rs_min(float a, float b) { return fminf(a, b); }
static inline double rs_maxd(double a, double b) { return fmax(a, b); }
static inline double rs_mind(double a, double b) { return fmin(a, b); }
/* f32::clamp */
static inline float rs_clamp(float x, float lo, float hi) {
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
/* half::f16 <-> f32 (half 2.7.1: round-to-nearest-even) */
static inline float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, u;
    if (exp == 0) {
        if (man == 0) u = sign;
        else { int e = -1; do { e++; man <<= 1; } while (!(man & 0x400u)); u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13); }
    } else if (exp == 31) u = sign | 0x7f800000u | (man << 13);
    else u = sign | ((exp + 112) << 23) | (man << 13);
    float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t f32_to_f16(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u, man = x & 0x7fffffu; int32_t exp = (int32_t)((x >> 23) & 0xff);
    if (exp == 255) return (uint16_t)(sign | 0x7c00u | (man ? (0x200u | (man >> 13)) : 0));
    int32_t e = exp - 127 + 15;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - e);
        uint32_t half = man >> shift, rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)e << 10) | (man >> 13), rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
    return (uint16_t)(sign | half);
}

/* util::map_coord — util.rs:144-147 (operation order preserved) */
static inline float map_coord(float x, float in_min, float in_max, float out_min, float out_max) {
    return (x - in_min) * (out_max - out_min) / (in_max - in_min) + out_min;
}

typedef struct { float x, y; } v2;

/* ------------------------------------------------------------------------------------------
 * Lens models — distortion_models/ (one .rs per model)
 * ---------------------------------------------------------------------------------------- */
#define RS_PI_F 3.14159274101257324f  /* std::f32::consts::PI */

/* opencv_fisheye.rs:12-70 */
static int fisheye_undistort(v2 p, const gf_kernel_params* P, v2* o) {
    const float* k = P->k;
    if (k[0] == 0.0f && k[1] == 0.0f && k[2] == 0.0f && k[3] == 0.0f) { *o = p; return 1; }
    const float EPS = 1e-6f;
    float theta_d = sqrtf(p.x * p.x + p.y * p.y);
    theta_d = rs_min(rs_max(theta_d, -RS_PI_F), RS_PI_F);
    int converged = 0;
    float theta = theta_d, scale = 0.0f;
    if (fabsf(theta_d) > EPS) {
        theta = 0.0f;
        for (int i = 0; i < 10; ++i) {
            float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta6 * theta2;
            float k0_theta2 = k[0] * theta2, k1_theta4 = k[1] * theta4, k2_theta6 = k[2] * theta6, k3_theta8 = k[3] * theta8;
            float theta_fix = (theta * (1.0f + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d)
                            / (1.0f + 3.0f * k0_theta2 + 5.0f * k1_theta4 + 7.0f * k2_theta6 + 9.0f * k3_theta8);
            theta_fix = rs_min(rs_max(theta_fix, -0.9f), 0.9f);
            theta = theta - theta_fix;
            if (fabsf(theta_fix) < EPS) { converged = 1; break; }
        }
        scale = tanf(theta) / theta_d;
    } else {
        converged = 1;
    }
    int theta_flipped = (theta_d < 0.0f && theta > 0.0f) || (theta_d > 0.0f && theta < 0.0f);
    if (converged && !theta_flipped) { o->x = p.x * scale; o->y = p.y * scale; return 1; }
    return 0;
}
/* opencv_fisheye.rs:72-93 */
static v2 fisheye_distort(float x, float y, float z, const gf_kernel_params* P) {
    const float* k = P->k;
    x = x / z; y = y / z;
    if (k[0] == 0.0f && k[1] == 0.0f && k[2] == 0.0f && k[3] == 0.0f) return (v2){x, y};
    float r = sqrtf(x * x + y * y);
    float theta = atanf(r);
    float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
    float theta_d = theta * (1.0f + k[0] * theta2 + k[1] * theta4 + k[2] * theta6 + k[3] * theta8);
    float scale = r == 0.0f ? 1.0f : theta_d / r;
    return (v2){x * scale, y * scale};
}

/* opencv_standard.rs:12-30 */
static int standard_undistort(v2 p, const gf_kernel_params* P, v2* o) {
    const float* k = P->k;
    float x = p.x, y = p.y, x0 = p.x, y0 = p.y;
    for (int i = 0; i < 20; ++i) {
        float r2 = x * x + y * y;
        float icdist = (1.0f + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1.0f + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0.0f) return 0;
        float delta_x = 2.0f * k[2] * x * y + k[3] * (r2 + 2.0f * x * x) + k[8] * r2 + k[9] * r2 * r2;
        float delta_y = k[2] * (r2 + 2.0f * y * y) + 2.0f * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - delta_x) * icdist;
        y = (y0 - delta_y) * icdist;
    }
    o->x = x; o->y = y; return 1;
}
/* opencv_standard.rs:32-48 */
static v2 standard_distort(float x, float y, float z, const gf_kernel_params* P) {
    const float* k = P->k;
    x = x / z; y = y / z;
    float r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    float a1 = 2.0f * x * y, a2 = r2 + 2.0f * x * x, a3 = r2 + 2.0f * y * y;
    float cdist = 1.0f + k[0] * r2 + k[1] * r4 + k[4] * r6;
    float icdist2 = 1.0f / (1.0f + k[5] * r2 + k[6] * r4 + k[7] * r6);
    float xd0 = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
    float yd0 = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
    return (v2){xd0, yd0};
}

#define NEWTON_EPS 0.00001f
/* poly3.rs:14-52 */
static int poly3_undistort(v2 p, const gf_kernel_params* P, v2* o) {
    float inv_k1 = 1.0f / P->k[0];
    float rd = sqrtf(p.x * p.x + p.y * p.y);
    if (rd == 0.0f) return 0;
    float rd_div_k1 = rd * inv_k1;
    float ru = rd;
    for (int i = 0; i < 10; ++i) {
        float fru = ru * ru * ru + ru * inv_k1 - rd_div_k1;
        if (fru >= -NEWTON_EPS && fru < NEWTON_EPS) break;
        if (i > 5) return 0;
        ru = ru - (fru / (3.0f * ru * ru + inv_k1));
    }
    if (ru < 0.0f) return 0;
    ru = ru / rd;
    o->x = p.x * ru; o->y = p.y * ru; return 1;
}
/* poly3.rs:54-63 */
static v2 poly3_distort(float x, float y, float z, const gf_kernel_params* P) {
    x = x / z; y = y / z;
    float poly2 = P->k[0] * (x * x + y * y) + 1.0f;
    return (v2){x * poly2, y * poly2};
}
/* poly5.rs:14-41 */
static int poly5_undistort(v2 p, const gf_kernel_params* P, v2* o) {
    const float* k = P->k;
    float rd = sqrtf(p.x * p.x + p.y * p.y);
    if (rd == 0.0f) return 0;
    float ru = rd;
    for (int i = 0; i < 10; ++i) {
        float ru2 = ru * ru;
        float fru = ru * (1.0f + k[0] * ru2 + k[1] * ru2 * ru2) - rd;
        if (fru >= -NEWTON_EPS && fru < NEWTON_EPS) break;
        if (i > 5) return 0;
        ru = ru - (fru / (1.0f + 3.0f * k[0] * ru2 + 5.0f * k[1] * ru2 * ru2));
    }
    if (ru < 0.0f) return 0;
    ru = ru / rd;
    o->x = p.x * ru; o->y = p.y * ru; return 1;
}
/* poly5.rs:43-53 */
static v2 poly5_distort(float x, float y, float z, const gf_kernel_params* P) {
    const float* k = P->k;
    x = x / z; y = y / z;
    float ru2 = x * x + y * y;
    float poly4 = 1.0f + k[0] * ru2 + k[1] * ru2 * ru2;
    return (v2){x * poly4, y * poly4};
}
/* ptlens.rs:14-40 */
static int ptlens_undistort(v2 p, const gf_kernel_params* P, v2* o) {
    const float* k = P->k;
    float rd = sqrtf(p.x * p.x + p.y * p.y);
    if (rd == 0.0f) return 0;
    float ru = rd;
    for (int i = 0; i < 10; ++i) {
        float fru = ru * (k[0] * ru * ru * ru + k[1] * ru * ru + k[2] * ru + 1.0f) - rd;
        if (fru >= -NEWTON_EPS && fru < NEWTON_EPS) break;
        if (i > 5) return 0;
        ru = ru - (fru / (4.0f * k[0] * ru * ru * ru + 3.0f * k[1] * ru * ru + 2.0f * k[2] * ru + 1.0f));
    }
    if (ru < 0.0f) return 0;
    ru = ru / rd;
    o->x = p.x * ru; o->y = p.y * ru; return 1;
}
/* ptlens.rs:42-53 */
static v2 ptlens_distort(float x, float y, float z, const gf_kernel_params* P) {
    const float* k = P->k;
    x = x / z; y = y / z;
    float ru2 = x * x + y * y;
    float r = sqrtf(ru2);
    float poly3 = k[0] * ru2 * r + k[1] * ru2 + k[2] * r + 1.0f;
    return (v2){x * poly3, y * poly3};
}
/* insta360.rs:27-48 */
static v2 insta360_distort(float x, float y, float z, const gf_kernel_params* P) {
    float k1 = P->k[0], k2 = P->k[1], k3 = P->k[2], p1 = P->k[3], p2 = P->k[4], xi = P->k[5];
    float len = sqrtf(x * x + y * y + z * z);
    x = (x / len) / ((z / len) + xi);
    y = (y / len) / ((z / len) + xi);
    float r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    return (v2){
        x * (1.0f + k1 * r2 + k2 * r4 + k3 * r6) + 2.0f * p1 * x * y + p2 * (r2 + 2.0f * x * x),
        y * (1.0f + k1 * r2 + k2 * r4 + k3 * r6) + 2.0f * p2 * x * y + p1 * (r2 + 2.0f * y * y)
    };
}
/* insta360.rs:10-25 */
static int insta360_undistort(v2 p, const gf_kernel_params* P, v2* o) {
    float px = p.x, py = p.y;
    for (int i = 0; i < 200; ++i) {
        v2 dp = insta360_distort(px, py, 1.0f, P);
        float dx = dp.x - p.x, dy = dp.y - p.y;
        if (fabsf(dx) < 1e-6f && fabsf(dy) < 1e-6f) break;
        px -= dx; py -= dy;
    }
    o->x = px; o->y = py; return 1;
}
/* sony.rs:10-63 */
static int sony_undistort(v2 p, const gf_kernel_params* P, v2* o) {
    const float* k = P->k;
    if (k[0] == 0.0f && k[1] == 0.0f && k[2] == 0.0f && k[3] == 0.0f) { *o = p; return 1; }
    const float EPS = 1e-6f;
    float theta_d = sqrtf(p.x * p.x + p.y * p.y);
    int converged = 0;
    float theta = theta_d, scale = 0.0f;
    if (fabsf(theta_d) > EPS) {
        theta = 0.0f;
        for (int i = 0; i < 10; ++i) {
            float theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta2 * theta3;
            float k0 = k[0], k1_theta1 = k[1] * theta, k2_theta2 = k[2] * theta2, k3_theta3 = k[3] * theta3,
                  k4_theta4 = k[4] * theta4, k5_theta5 = k[5] * theta5;
            float theta_fix = (theta * (k0 + k1_theta1 + k2_theta2 + k3_theta3 + k4_theta4 + k5_theta5) - theta_d)
                            / (k0 + 2.0f * k1_theta1 + 3.0f * k2_theta2 + 4.0f * k3_theta3 + 5.0f * k4_theta4 + 6.0f * k5_theta5);
            theta = theta - theta_fix;
            if (fabsf(theta_fix) < EPS) { converged = 1; break; }
        }
        scale = tanf(theta) / theta_d;
    } else {
        converged = 1;
    }
    int theta_flipped = (theta_d < 0.0f && theta > 0.0f) || (theta_d > 0.0f && theta < 0.0f);
    if (converged && !theta_flipped) { o->x = p.x * scale; o->y = p.y * scale; return 1; }
    return 0;
}
/* sony.rs:65-89 */
static v2 sony_distort(float x, float y, float z, const gf_kernel_params* P) {
    const float* k = P->k;
    x = x / z; y = y / z;
    if (k[0] == 0.0f && k[1] == 0.0f && k[2] == 0.0f && k[3] == 0.0f) return (v2){x, y};
    float r = sqrtf(x * x + y * y);
    float theta = atanf(r);
    float theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta2 * theta3, theta6 = theta3 * theta3;
    float theta_d = theta * k[0] + theta2 * k[1] + theta3 * k[2] + theta4 * k[3] + theta5 * k[4] + theta6 * k[5];
    float scale = r == 0.0f ? 1.0f : theta_d / r;
    return (v2){x * scale, y * scale};
}
static int genpoly_all_zero(const float* k) {
    for (int i = 0; i < 12; ++i) if (!(k[i] == 0.0f)) return 0;
    return 1;
}
/* generic_polynomial.rs:18-81 */
static int genpoly_undistort(v2 p, const gf_kernel_params* P, v2* o) {
    const float* k = P->k;
    if (genpoly_all_zero(k)) { *o = p; return 1; }
    const float EPS = 1e-6f;
    float theta_d = sqrtf(p.x * p.x + p.y * p.y);
    int converged = 0;
    float theta = theta_d, scale = 0.0f;
    if (fabsf(theta_d) > EPS) {
        theta = 0.0f;
        for (int i = 0; i < 10; ++i) {
            float theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta2 * theta3,
                  theta6 = theta3 * theta3, theta7 = theta3 * theta4, theta8 = theta4 * theta4, theta9 = theta4 * theta5,
                  theta10 = theta5 * theta5, theta11 = theta5 * theta6;
            float k0 = k[0], k1t = k[1] * theta, k2t = k[2] * theta2, k3t = k[3] * theta3, k4t = k[4] * theta4, k5t = k[5] * theta5,
                  k6t = k[6] * theta6, k7t = k[7] * theta7, k8t = k[8] * theta8, k9t = k[9] * theta9, k10t = k[10] * theta10, k11t = k[11] * theta11;
            float theta_fix = (theta * (k0 + k1t + k2t + k3t + k4t + k5t + k6t + k7t + k8t + k9t + k10t + k11t) - theta_d)
                            / (k0 + 2.0f * k1t + 3.0f * k2t + 4.0f * k3t + 5.0f * k4t + 6.0f * k5t + 7.0f * k6t + 8.0f * k7t + 9.0f * k8t + 10.0f * k9t + 11.0f * k10t + 12.0f * k11t);
            theta = theta - theta_fix;
            if (fabsf(theta_fix) < EPS) { converged = 1; break; }
        }
        scale = tanf(theta) / theta_d;
    } else {
        converged = 1;
    }
    int theta_flipped = (theta_d < 0.0f && theta > 0.0f) || (theta_d > 0.0f && theta < 0.0f);
    if (converged && !theta_flipped) { o->x = p.x * scale; o->y = p.y * scale; return 1; }
    return 0;
}
/* generic_polynomial.rs:83-122 */
static v2 genpoly_distort(float x, float y, float z, const gf_kernel_params* P) {
    const float* k = P->k;
    x = x / z; y = y / z;
    if (genpoly_all_zero(k)) return (v2){x, y};
    float r = sqrtf(x * x + y * y);
    float theta = atanf(r);
    float theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta2 * theta3,
          theta6 = theta3 * theta3, theta7 = theta3 * theta4, theta8 = theta4 * theta4, theta9 = theta4 * theta5,
          theta10 = theta5 * theta5, theta11 = theta5 * theta6, theta12 = theta6 * theta6;
    float theta_d = theta * k[0] + theta2 * k[1] + theta3 * k[2] + theta4 * k[3] + theta5 * k[4] + theta6 * k[5]
                  + theta7 * k[6] + theta8 * k[7] + theta9 * k[8] + theta10 * k[9] + theta11 * k[10] + theta12 * k[11];
    float scale = r == 0.0f ? 1.0f : theta_d / r;
    return (v2){x * scale, y * scale};
}
/* gopro.rs:19-38 */
static inline float gopro_poly_eval(float p, const float* k) {
    return k[0] + p * (k[1] + p * (k[2] + p * (k[3] + p * (k[4] + p * (k[5] + p * k[6])))));
}
static inline float gopro_poly_deriv(float p, const float* k) {
    return k[1] + p * (2.0f * k[2] + p * (3.0f * k[3] + p * (4.0f * k[4] + p * (5.0f * k[5] + p * (6.0f * k[6])))));
}
static float gopro_poly_invert(float theta, const float* k) {
    float p = (theta - k[0]) / k[1];
    for (int i = 0; i < 10; ++i) {
        float d = gopro_poly_deriv(p, k);
        if (fabsf(d) < 1e-12f) break;
        float fix = (gopro_poly_eval(p, k) - theta) / d;
        p -= fix;
        if (fabsf(fix) < 1e-7f) break;
    }
    return p;
}
#define GOPRO_TMAX 1.5533f
/* gopro.rs:42-57 */
static int gopro_undistort(v2 pt, const gf_kernel_params* P, v2* o) {
    const float* k = P->k;
    if (k[1] == 0.0f) { *o = pt; return 1; }
    float r_norm = sqrtf(pt.x * pt.x + pt.y * pt.y);
    if (r_norm < 1e-9f) { *o = pt; return 1; }
    float p = r_norm / k[1];
    float theta = gopro_poly_eval(p, k);
    float tt = tanf(GOPRO_TMAX);
    float rr = theta < GOPRO_TMAX ? tanf(theta) : tt + (theta - GOPRO_TMAX) * (1.0f + tt * tt);
    float scale = rr / r_norm;
    o->x = pt.x * scale; o->y = pt.y * scale; return 1;
}
/* gopro.rs:61-74 */
static v2 gopro_distort(float x, float y, float z, const gf_kernel_params* P) {
    const float* k = P->k;
    v2 pos = {x / z, y / z};
    if (k[1] == 0.0f) return pos;
    float r = sqrtf(pos.x * pos.x + pos.y * pos.y);
    float tt = tanf(GOPRO_TMAX);
    float theta = r < tt ? atanf(r) : GOPRO_TMAX + (r - tt) / (1.0f + tt * tt);
    float p = gopro_poly_invert(theta, k);
    float r_norm = k[1] * p;
    float scale = r < 1e-9f ? 1.0f : r_norm / r;
    return (v2){pos.x * scale, pos.y * scale};
}

/* ---- digital lenses -------------------------------------------------------------------- */
/* gopro_superview.rs:12-19 */
static v2 superview_fn(v2 uv) {
    float x2 = uv.x * uv.x, y2 = uv.y * uv.y;
    return (v2){
        uv.x * (1.2100393f + x2 * (-1.2758402f + x2 * 1.7751845f)),
        uv.y * (0.9364505f + (0.4465308f - 0.7683315f * y2) * y2 + (-0.3574087f + 1.1584653f * y2 + 0.3529348f * x2) * x2)
    };
}
/* gopro6_superview.rs:12-17 */
static v2 superview6_fn(v2 uv) {
    uv.x *= 1.0f - 0.48f * fabsf(uv.x);
    uv.x *= 0.943396f * (1.0f + 0.157895f * fabsf(uv.x));
    uv.y *= 0.943396f * (1.0f + 0.060000f * fabsf(uv.y * 2.0f));
    return uv;
}
/* gopro_hyperview.rs:10-17 */
static v2 hyperview_fn(v2 uv) {
    float x2 = uv.x * uv.x, y2 = uv.y * uv.y;
    return (v2){
        uv.x * (1.5805143f + x2 * (-8.1668825f + x2 * (74.5198746f + x2 * (-451.5002441f + x2 * (1551.2922363f + x2 * (-2735.5422363f + x2 * 1923.1572266f))))) + y2 * -0.1086027f),
        uv.y * (1.0238225f + y2 * -0.1025671f + x2 * (-0.2639930f + x2 * 0.2979266f))
    };
}
/* gopro_warp.rs:22-39 */
static v2 gopro_map(v2 uv, const float* p) {
    float x = rs_clamp(uv.x, -0.5f, 0.5f), y = rs_clamp(uv.y, -0.5f, 0.5f);
    float x2 = x * x, y2 = y * y;
    float poly_x = p[0] + x2 * (p[1] + x2 * (p[2] + x2 * (p[3] + x2 * (p[4] + x2 * (p[5] + x2 * p[6])))));
    return (v2){
        x * (poly_x + p[7] * y2) + (uv.x - x),
        y * (p[8] + p[9] * y2 + p[10] * y2 * y2 + x2 * (p[11] + p[12] * y2 + p[13] * x2)) + (uv.y - y)
    };
}

/* "uv range (0,0)..(width,height)": {superview,superview6,hyperview}.rs undistort_point */
static int digital_undistort(int model, v2 uv, const gf_kernel_params* P, v2* o) {
    v2 out_c2 = {(float)P->output_width, (float)P->output_height};
    switch (model) {
    case GF_LENS_GOPRO_SUPERVIEW:   /* gopro_superview.rs:23-34 */
        uv.x = (uv.x / out_c2.x) - 0.5f; uv.y = (uv.y / out_c2.y) - 0.5f;
        uv = superview_fn(uv);
        uv.x = uv.x / 1.333333333f;
        o->x = (uv.x + 0.5f) * out_c2.x; o->y = (uv.y + 0.5f) * out_c2.y; return 1;
    case GF_LENS_GOPRO6_SUPERVIEW:  /* gopro6_superview.rs:21-30 */
        uv.x = (uv.x / out_c2.x) - 0.5f; uv.y = (uv.y / out_c2.y) - 0.5f;
        uv = superview6_fn(uv);
        o->x = (uv.x + 0.5f) * out_c2.x; o->y = (uv.y + 0.5f) * out_c2.y; return 1;
    case GF_LENS_GOPRO_HYPERVIEW:   /* gopro_hyperview.rs:21-32 */
        uv.x = (uv.x / out_c2.x) - 0.5f; uv.y = (uv.y / out_c2.y) - 0.5f;
        uv = hyperview_fn(uv);
        uv.x = uv.x / 1.555555555f;
        o->x = (uv.x + 0.5f) * out_c2.x; o->y = (uv.y + 0.5f) * out_c2.y; return 1;
    case GF_LENS_GOPRO_WARP: {      /* gopro_warp.rs:43-56 */
        const float* p = P->digital_lens_params;
        float factor = p[14] != 0.0f ? p[14] : 1.0f;
        uv.x = (uv.x / out_c2.x) - 0.5f; uv.y = (uv.y / out_c2.y) - 0.5f;
        uv = gopro_map(uv, p);
        uv.x = uv.x / factor;
        o->x = (uv.x + 0.5f) * out_c2.x; o->y = (uv.y + 0.5f) * out_c2.y; return 1;
    }
    case GF_LENS_DIGITAL_STRETCH:   /* digital_stretch.rs:12-15 */
        o->x = uv.x / P->digital_lens_params[0]; o->y = uv.y / P->digital_lens_params[1]; return 1;
    default: return -1;
    }
}
typedef v2 (*warp_fn)(v2);
static v2 fixed_point_invert(warp_fn f, float x, float y) {   /* the 12-iteration loops of *_view.rs distort_point */
    v2 pp = {x, y};
    for (int i = 0; i < 12; ++i) {
        v2 dp = f(pp);
        float dx = dp.x - x, dy = dp.y - y;
        if (fabsf(dx) < 1e-6f && fabsf(dy) < 1e-6f) break;
        pp.x -= dx; pp.y -= dy;
    }
    return pp;
}
static int digital_distort(int model, float x, float y, const gf_kernel_params* P, v2* o) {
    v2 size = {(float)P->width, (float)P->height};
    switch (model) {
    case GF_LENS_GOPRO_SUPERVIEW: {  /* gopro_superview.rs:38-57 */
        x = (x / size.x) - 0.5f; y = (y / size.y) - 0.5f;
        x = x * 1.333333333f;
        v2 pp = fixed_point_invert(superview_fn, x, y);
        o->x = (pp.x + 0.5f) * size.x; o->y = (pp.y + 0.5f) * size.y; return 1;
    }
    case GF_LENS_GOPRO6_SUPERVIEW: { /* gopro6_superview.rs:34-51 */
        x = (x / size.x) - 0.5f; y = (y / size.y) - 0.5f;
        v2 pp = fixed_point_invert(superview6_fn, x, y);
        o->x = (pp.x + 0.5f) * size.x; o->y = (pp.y + 0.5f) * size.y; return 1;
    }
    case GF_LENS_GOPRO_HYPERVIEW: {  /* gopro_hyperview.rs:36-55 */
        x = (x / size.x) - 0.5f; y = (y / size.y) - 0.5f;
        x = x * 1.555555555f;
        v2 pp = fixed_point_invert(hyperview_fn, x, y);
        o->x = (pp.x + 0.5f) * size.x; o->y = (pp.y + 0.5f) * size.y; return 1;
    }
    case GF_LENS_GOPRO_WARP: {       /* gopro_warp.rs:60-94 */
        const float* p = P->digital_lens_params;
        float factor = p[14] != 0.0f ? p[14] : 1.0f;
        x = (x / size.x) - 0.5f; y = (y / size.y) - 0.5f;
        v2 target = {x * factor, y};
        v2 pp = {x, y};
        for (int i = 0; i < 12; ++i) {
            v2 dp = gopro_map(pp, p);
            float dx = dp.x - target.x, dy = dp.y - target.y;
            if (fabsf(dx) < 1e-6f && fabsf(dy) < 1e-6f) break;
            pp.x -= dx; pp.y -= dy;
        }
        v2 res = gopro_map(pp, p);
        if (fabsf(res.x - target.x) > 0.02f || fabsf(res.y - target.y) > 0.02f) { o->x = -99999.0f; o->y = -99999.0f; return 1; }
        o->x = (pp.x + 0.5f) * size.x; o->y = (pp.y + 0.5f) * size.y; return 1;
    }
    case GF_LENS_DIGITAL_STRETCH:    /* digital_stretch.rs:19-22 */
        o->x = x * P->digital_lens_params[0]; o->y = y * P->digital_lens_params[1]; return 1;
    default: return -1;
    }
}

/* DistortionModel::undistort_point / distort_point — distortion_models/mod.rs:36-45.
 * Digital models are also reachable here (the enum does not distinguish). */
static int lens_undistort(int model, v2 p, const gf_kernel_params* P, v2* o) {
    switch (model) {
    case GF_LENS_OPENCV_FISHEYE:     return fisheye_undistort(p, P, o);
    case GF_LENS_OPENCV_STANDARD:    return standard_undistort(p, P, o);
    case GF_LENS_POLY3:              return poly3_undistort(p, P, o);
    case GF_LENS_POLY5:              return poly5_undistort(p, P, o);
    case GF_LENS_PTLENS:             return ptlens_undistort(p, P, o);
    case GF_LENS_INSTA360:           return insta360_undistort(p, P, o);
    case GF_LENS_SONY:               return sony_undistort(p, P, o);
    case GF_LENS_GENERIC_POLYNOMIAL: return genpoly_undistort(p, P, o);
    case GF_LENS_GOPRO:              return gopro_undistort(p, P, o);
    default: { int r = digital_undistort(model, p, P, o); if (r < 0) { *o = p; return 1; } return r; }
    }
}
static v2 lens_distort(int model, float x, float y, float z, const gf_kernel_params* P) {
    switch (model) {
    case GF_LENS_OPENCV_FISHEYE:     return fisheye_distort(x, y, z, P);
    case GF_LENS_OPENCV_STANDARD:    return standard_distort(x, y, z, P);
    case GF_LENS_POLY3:              return poly3_distort(x, y, z, P);
    case GF_LENS_POLY5:              return poly5_distort(x, y, z, P);
    case GF_LENS_PTLENS:             return ptlens_distort(x, y, z, P);
    case GF_LENS_INSTA360:           return insta360_distort(x, y, z, P);
    case GF_LENS_SONY:               return sony_distort(x, y, z, P);
    case GF_LENS_GENERIC_POLYNOMIAL: return genpoly_distort(x, y, z, P);
    case GF_LENS_GOPRO:              return gopro_distort(x, y, z, P);
    default: { v2 o = {x, y}; digital_distort(model, x, y, P, &o); return o; }
    }
}

int gf_oracle_lens_undistort_point(int model, float x, float y, const gf_kernel_params* P, float* ox, float* oy) {
    v2 o = {0, 0}; int r = lens_undistort(model, (v2){x, y}, P, &o); *ox = o.x; *oy = o.y; return r;
}
void gf_oracle_lens_distort_point(int model, float x, float y, float z, const gf_kernel_params* P, float* ox, float* oy) {
    v2 o = lens_distort(model, x, y, z, P); *ox = o.x; *oy = o.y;
}

/* ------------------------------------------------------------------------------------------
 * Mesh correction — gyro_source/splines.rs:88-176, sony.rs:557-563 (all f64)
 * ---------------------------------------------------------------------------------------- */
#define MAX_GRID_SIZE 9

/* splines.rs:100-124 */
static void cubic_spline_coefficients(const double* mesh, size_t step, size_t offset, double size, size_t n,
                                      double* a, double* b, double* c, double* d, double* alpha, double* mu, double* z) {
    double h = size / (double)(n - 1);
    double inv_h = 1.0 / h;
    double three_inv_h = 3.0 * inv_h;
    double h_over_3 = h / 3.0;
    double inv_3h = 1.0 / (3.0 * h);
    for (size_t i = 0; i < n; ++i) a[i] = mesh[(i + offset) * step];
    for (size_t i = 1; i + 1 < n; ++i) alpha[i] = three_inv_h * (a[i + 1] - 2.0 * a[i] + a[i - 1]);
    mu[0] = 0.0; z[0] = 0.0;
    for (size_t i = 1; i + 1 < n; ++i) {
        mu[i] = 1.0 / (4.0 - mu[i - 1]);
        z[i] = (alpha[i] * inv_h - z[i - 1]) * mu[i];
    }
    c[n - 1] = 0.0;
    for (size_t jj = n - 1; jj-- > 0;) {
        c[jj] = z[jj] - mu[jj] * c[jj + 1];
        b[jj] = (a[jj + 1] - a[jj]) * inv_h - h_over_3 * (c[jj + 1] + 2.0 * c[jj]);
        d[jj] = (c[jj + 1] - c[jj]) * inv_3h;
    }
}
void gf_oracle_cubic_spline_coefficients(const double* mesh, size_t step, size_t offset, double size, size_t n,
                                         double* a, double* b, double* c, double* d) {
    double alpha[MAX_GRID_SIZE] = {0}, mu[MAX_GRID_SIZE] = {0}, z[MAX_GRID_SIZE] = {0};
    for (int i = 0; i < MAX_GRID_SIZE; ++i) { a[i] = b[i] = c[i] = d[i] = 0.0; }
    cubic_spline_coefficients(mesh, step, offset, size, n, a, b, c, d, alpha, mu, z);
}
/* splines.rs:126-139 */
static double cubic_spline_interpolate(const double* a, const double* b, const double* c, const double* d, size_t n, double x, double size) {
    if (x <= 0.0) return a[0] + b[0] * x;
    if (x >= size) {
        double h = size / (double)(n - 1);
        double slope = b[n - 2] + 2.0 * c[n - 2] * h + 3.0 * d[n - 2] * h * h;
        return a[n - 1] + slope * (x - size);
    }
    size_t i = rs_f64_as_usize(((double)n - 1.0) * x / size);
    if (i > n - 2) i = n - 2;
    double dx = x - size * (double)i / (double)(n - 1);
    return a[i] + b[i] * dx + c[i] * dx * dx + d[i] * dx * dx * dx;
}
/* splines.rs:141-176 */
static double bivariate_interpolate(size_t n_x, size_t n_y, double size_x, double size_y, const double* mesh, size_t mesh_offset, double x, double y) {
    double iv[MAX_GRID_SIZE] = {0}, a[MAX_GRID_SIZE] = {0}, b[MAX_GRID_SIZE] = {0}, c[MAX_GRID_SIZE] = {0}, d[MAX_GRID_SIZE] = {0};
    double alpha[MAX_GRID_SIZE] = {0}, mu[MAX_GRID_SIZE] = {0}, z[MAX_GRID_SIZE] = {0};
    size_t i = rs_f64_as_usize(((double)n_x - 1.0) * x / size_x);
    if (i > n_x - 2) i = n_x - 2;
    double dx = x - size_x * (double)i / (double)(n_x - 1);
    double dx2 = dx * dx;
    size_t grid = MAX_GRID_SIZE, raw_mesh_len = n_x * n_y * 2, block = grid * 4;
    size_t coeff_base = 9 + raw_mesh_len + (mesh_offset * n_y * block);
    size_t offs = coeff_base + i;
    for (size_t j = 0; j < n_y; ++j) {
        size_t rb = offs + j * block;
        iv[j] = mesh[rb + grid * 0] + mesh[rb + grid * 1] * dx + mesh[rb + grid * 2] * dx2 + mesh[rb + grid * 3] * dx2 * dx;
    }
    cubic_spline_coefficients(iv, 1, 0, size_y, n_y, a, b, c, d, alpha, mu, z);
    return cubic_spline_interpolate(a, b, c, d, n_y, y, size_y);
}
/* sony.rs:557-563 — size = (mesh[3], mesh[4]) at the call site cpu_undistort.rs:170,179 */
void gf_oracle_interpolate_mesh(double x, double y, const double* mesh, double* ox, double* oy) {
    size_t n_x = rs_f64_as_usize(mesh[1]), n_y = rs_f64_as_usize(mesh[2]);
    *ox = bivariate_interpolate(n_x, n_y, mesh[3], mesh[4], mesh, 0, x, y);
    *oy = bivariate_interpolate(n_x, n_y, mesh[3], mesh[4], mesh, 1, x, y);
}

/* ------------------------------------------------------------------------------------------
 * rotate_and_distort — cpu_undistort.rs:133-228
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const gf_kernel_params* P;
    const float* matrices;       /* rows x 14 */
    size_t matrix_rows;
    int model, digital;          /* digital: GF_LENS_NONE for Option::None */
    float r_limit_sq;
    const double* mesh;          /* widened copy, cpu_undistort.rs:539 */
    size_t mesh_len;
} warp_ctx;

static int rotate_and_distort(v2 pos, size_t idx, const warp_ctx* W, v2* out) {
    const gf_kernel_params* P = W->P;
    const float* m = W->matrices + idx * GF_MATRIX_STRIDE;
    const double* mesh = W->mesh; const size_t mesh_len = W->mesh_len;
    float _x = (pos.x * m[0]) + (pos.y * m[1]) + m[2] + P->translation3d[0];
    float _y = (pos.x * m[3]) + (pos.y * m[4]) + m[5] + P->translation3d[1];
    float _w = (pos.x * m[6]) + (pos.y * m[7]) + m[8] + P->translation3d[2];
    if (_w > 0.0f) {
        if (W->r_limit_sq > 0.0f && (_x * _x + _y * _y) > W->r_limit_sq * _w) return 0;   /* :139 (sic: * _w) */

        if (P->light_refraction_coefficient != 1.0f && P->light_refraction_coefficient > 0.0f) {   /* :143-152 */
            if (_w != 0.0f) {
                float r = sqrtf(_x * _x + _y * _y) / _w;
                float sin_theta_d = (r / sqrtf(1.0f + r * r)) * P->light_refraction_coefficient;
                float r_d = sin_theta_d / sqrtf(1.0f - sin_theta_d * sin_theta_d);
                if (r_d != 0.0f) _w *= r / r_d;
            }
        }

        v2 uv = lens_distort(W->model, _x, _y, _w, P);           /* :154 */
        uv.x = uv.x * P->f[0]; uv.y = uv.y * P->f[1];             /* :155 */

        if (m[9] != 0.0f || m[10] != 0.0f || m[11] != 0.0f || m[12] != 0.0f || m[13] != 0.0f) {   /* :157-165 */
            float ang_rad = m[11];
            float cos_a = cosf(-ang_rad), sin_a = sinf(-ang_rad);
            v2 t = { cos_a * uv.x - sin_a * uv.y - m[9]  + m[12],
                     sin_a * uv.x + cos_a * uv.y - m[10] + m[13] };
            uv = t;
        }

        uv.x = uv.x + P->c[0]; uv.y = uv.y + P->c[1];             /* :167 */

        if (mesh_len > 0 && mesh[0] > 10.0) {                      /* :169-185 */
            float origin_x = (float)mesh[5], origin_y = (float)mesh[6];
            float crop_w = (float)mesh[7], crop_h = (float)mesh[8];
            if ((P->flags & 128) == 128) uv.y = (float)P->height - uv.y;
            uv.x = map_coord(uv.x, 0.0f, (float)P->width,  origin_x, origin_x + crop_w);
            uv.y = map_coord(uv.y, 0.0f, (float)P->height, origin_y, origin_y + crop_h);
            double nx, ny;
            gf_oracle_interpolate_mesh((double)uv.x, (double)uv.y, mesh, &nx, &ny);
            uv.x = map_coord((float)nx, origin_x, origin_x + crop_w, 0.0f, (float)P->width);
            uv.y = map_coord((float)ny, origin_y, origin_y + crop_h, 0.0f, (float)P->height);
            if ((P->flags & 128) == 128) uv.y = (float)P->height - uv.y;
        }

        /* FocalPlaneDistortion :188-214.  The reference indexes mesh_data[mesh_data[0]] unchecked (would panic
         * when no FPD block follows); here a missing block simply means "no FPD". */
        if (mesh_len > 0 && mesh[0] > 0.0 && rs_f64_as_usize(mesh[0]) < mesh_len && mesh[rs_f64_as_usize(mesh[0])] > 0.0) {
            size_t o = rs_f64_as_usize(mesh[0]);
            double mesh_size_y = mesh[4];
            float origin_x = (float)mesh[5], origin_y = (float)mesh[6];
            float crop_w = (float)mesh[7], crop_h = (float)mesh[8];
            double stblz_grid = mesh_size_y / 8.0;
            if ((P->flags & 128) == 128) uv.y = (float)P->height - uv.y;
            uv.x = map_coord(uv.x, 0.0f, (float)P->width,  origin_x, origin_x + crop_w);
            uv.y = map_coord(uv.y, 0.0f, (float)P->height, origin_y, origin_y + crop_h);
            size_t idx2 = rs_f64_as_usize(rs_mind(rs_maxd(floor((double)uv.y / stblz_grid), 0.0), 7.0));
            double delta = (double)uv.y - stblz_grid * (double)idx2;
            uv.x -= (float)(mesh[o + 4 + idx2 * 2 + 0] * delta);
            uv.y -= (float)(mesh[o + 4 + idx2 * 2 + 1] * delta);
            for (size_t j = 0; j < idx2; ++j) {
                uv.x -= (float)(mesh[o + 4 + j * 2 + 0] * stblz_grid);
                uv.y -= (float)(mesh[o + 4 + j * 2 + 1] * stblz_grid);
            }
            uv.x = map_coord(uv.x, origin_x, origin_x + crop_w, 0.0f, (float)P->width);
            uv.y = map_coord(uv.y, origin_y, origin_y + crop_h, 0.0f, (float)P->height);
            if ((P->flags & 128) == 128) uv.y = (float)P->height - uv.y;
        }

        if ((P->flags & 2) == 2 && W->digital != GF_LENS_NONE) {   /* :216-220 */
            uv = lens_distort(W->digital, uv.x, uv.y, 1.0f, P);
        }

        if (P->input_horizontal_stretch > 0.001f) uv.x /= P->input_horizontal_stretch;   /* :222-223 */
        if (P->input_vertical_stretch   > 0.001f) uv.y /= P->input_vertical_stretch;

        *out = uv;
        return 1;
    }
    return 0;
}

/* cpu_undistort.rs:262-265 */
static v2 rotate_point(v2 pos, float angle, v2 origin, v2 origin2) {
    return (v2){ cosf(angle) * (pos.x - origin.x) - sinf(angle) * (pos.y - origin.y) + origin2.x,
                 sinf(angle) * (pos.x - origin.x) + cosf(angle) * (pos.y - origin.y) + origin2.y };
}

/* undistort_coord — cpu_undistort.rs:421-517 */
static int undistort_coord(v2 out_pos, const warp_ctx* W, v2 out_c, v2 out_f, v2* result) {
    const gf_kernel_params* P = W->P;
    out_pos.x = map_coord(out_pos.x, (float)P->output_rect[0], (float)(P->output_rect[0] + P->output_rect[2]), 0.0f, (float)P->output_width);
    out_pos.y = map_coord(out_pos.y, (float)P->output_rect[1], (float)(P->output_rect[1] + P->output_rect[3]), 0.0f, (float)P->output_height);
    out_pos.x += P->translation2d[0];
    out_pos.y += P->translation2d[1];

    if (P->lens_correction_amount < 1.0f) {                               /* :429-460 */
        v2 np = out_pos;
        if ((P->flags & 2) == 2 && W->digital != GF_LENS_NONE) {
            v2 uz = { (np.x - out_c.x) * P->fov + out_c.x, (np.y - out_c.y) * P->fov + out_c.y };
            v2 pt;
            if (lens_undistort(W->digital, uz, P, &pt)) {
                np.x = (pt.x - out_c.x) / P->fov + out_c.x;
                np.y = (pt.y - out_c.y) / P->fov + out_c.y;
            }
        }
        np.x = (np.x - out_c.x) / out_f.x; np.y = (np.y - out_c.y) / out_f.y;
        v2 pt;
        if (lens_undistort(W->model, np, P, &pt)) np = pt;
        if (P->light_refraction_coefficient != 1.0f && P->light_refraction_coefficient > 0.0f) {
            float r = sqrtf(np.x * np.x + np.y * np.y);
            if (r != 0.0f) {
                float sin_theta_d = (r / sqrtf(1.0f + r * r)) / P->light_refraction_coefficient;
                float r_d = sin_theta_d / sqrtf(1.0f - sin_theta_d * sin_theta_d);
                float factor = r_d / r;
                np.x *= factor; np.y *= factor;
            }
        }
        np.x = (np.x * out_f.x) + out_c.x; np.y = (np.y * out_f.y) + out_c.y;
        float ia = 1.0f - P->lens_correction_amount;
        out_pos.x = np.x * ia + (out_pos.x * P->lens_correction_amount);
        out_pos.y = np.y * ia + (out_pos.y * P->lens_correction_amount);
    }

    /* :465-479 rolling-shutter row */
    const int hrs = (P->flags & 16) == 16;
    int32_t sy_i = hrs ? rs_f32_as_i32(rs_round(out_pos.x)) : rs_f32_as_i32(rs_round(out_pos.y));
    { int32_t lim = hrs ? P->width : P->height; if (sy_i > lim) sy_i = lim; if (sy_i < 0) sy_i = 0; }
    size_t sy = (size_t)sy_i;
    if (P->matrix_count > 1) {
        size_t idx = (size_t)P->matrix_count / 2;
        v2 pt;
        if (rotate_and_distort(out_pos, idx, W, &pt)) {
            int32_t v = hrs ? rs_f32_as_i32(rs_round(pt.x)) : rs_f32_as_i32(rs_round(pt.y));
            int32_t lim = hrs ? P->width : P->height;
            if (v > lim) v = lim; if (v < 0) v = 0;
            sy = (size_t)v;
        }
    }

    size_t idx = sy; { size_t last = (size_t)P->matrix_count - 1; if (idx > last) idx = last; }   /* :482 */
    v2 uv;
    if (!rotate_and_distort(out_pos, idx, W, &uv)) return 0;                                        /* :483 */
    v2 frame_size = {(float)P->width, (float)P->height};
    if (P->input_rotation != 0.0f) {                                                                 /* :485-491 */
        float rotation = P->input_rotation * (RS_PI_F / 180.0f);
        v2 size = frame_size;
        frame_size = rotate_point(size, rotation, (v2){0.0f, 0.0f}, (v2){0.0f, 0.0f});
        frame_size.x = rs_round(fabsf(frame_size.x)); frame_size.y = rs_round(fabsf(frame_size.y));
        uv = rotate_point(uv, rotation, (v2){size.x / 2.0f, size.y / 2.0f}, (v2){frame_size.x / 2.0f, frame_size.y / 2.0f});
    }

    float width_f = (float)P->width, height_f = (float)P->height;
    if (P->background_mode == 1) {                /* edge repeat :495-499 */
        uv.x = rs_min(rs_max(uv.x, 3.0f), width_f - 3.0f);
        uv.y = rs_min(rs_max(uv.y, 3.0f), height_f - 3.0f);
    } else if (P->background_mode == 2) {         /* edge mirror :500-509 */
        float rx = rs_round(uv.x), ry = rs_round(uv.y);
        float width3 = width_f - 3.0f, height3 = height_f - 3.0f;
        if (rx > width3)  uv.x = width3  - (rx - width3);
        if (rx < 3.0f)    uv.x = 3.0f + width_f - (width3 + rx);
        if (ry > height3) uv.y = height3 - (ry - height3);
        if (ry < 3.0f)    uv.y = 3.0f + height_f - (height3 + ry);
    }
    if (P->background_mode != 3) {                /* :510-515 */
        uv.x = map_coord(uv.x, 0.0f, frame_size.x, (float)P->source_rect[0], (float)(P->source_rect[0] + P->source_rect[2]));
        uv.y = map_coord(uv.y, 0.0f, frame_size.y, (float)P->source_rect[1], (float)(P->source_rect[1] + P->source_rect[3]));
    }
    *result = uv;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Pixel formats — pixel_formats.rs.  (count, scalar kind) is all the kernel needs.
 * ---------------------------------------------------------------------------------------- */
enum { SC_U8 = 0, SC_U16 = 1, SC_F32 = 2, SC_F16 = 3 };
static int pix_layout(int pixel_type, int* count, int* scalar) {
    switch (pixel_type) {
    case GF_PIX_LUMA8:   *count = 1; *scalar = SC_U8;  return 1;
    case GF_PIX_LUMA16:  *count = 1; *scalar = SC_U16; return 1;
    case GF_PIX_RGB8:    *count = 3; *scalar = SC_U8;  return 1;
    case GF_PIX_RGBA8:   *count = 4; *scalar = SC_U8;  return 1;
    case GF_PIX_BGRA8:   *count = 4; *scalar = SC_U8;  return 1;
    case GF_PIX_RGB16:   *count = 3; *scalar = SC_U16; return 1;
    case GF_PIX_RGBA16:  *count = 4; *scalar = SC_U16; return 1;
    case GF_PIX_AYUV16:  *count = 4; *scalar = SC_U16; return 1;
    case GF_PIX_RGBAF:   *count = 4; *scalar = SC_F32; return 1;
    case GF_PIX_RGBAF16: *count = 4; *scalar = SC_F16; return 1;
    case GF_PIX_R32F:    *count = 1; *scalar = SC_F32; return 1;
    case GF_PIX_UV8:     *count = 2; *scalar = SC_U8;  return 1;
    case GF_PIX_UV16:    *count = 2; *scalar = SC_U16; return 1;
    default: return 0;
    }
}
typedef struct { float v[4]; } v4;

static inline v4 pix_to_float(const uint8_t* p, int count, int scalar) {   /* PixelType::to_float */
    v4 r = {{0.0f, 0.0f, 0.0f, 0.0f}};
    for (int i = 0; i < count; ++i) {
        switch (scalar) {
        case SC_U8:  r.v[i] = (float)p[i]; break;
        case SC_U16: { uint16_t t; memcpy(&t, p + 2 * i, 2); r.v[i] = (float)t; } break;
        case SC_F32: { float t; memcpy(&t, p + 4 * i, 4); r.v[i] = t; } break;
        default:     { uint16_t t; memcpy(&t, p + 2 * i, 2); r.v[i] = f16_to_f32(t); } break;
        }
    }
    return r;
}
static inline void pix_from_float(uint8_t* p, v4 val, int count, int scalar) {   /* PixelType::from_float (`as` casts) */
    for (int i = 0; i < count; ++i) {
        switch (scalar) {
        case SC_U8:  p[i] = rs_f32_as_u8(val.v[i]); break;
        case SC_U16: { uint16_t t = rs_f32_as_u16(val.v[i]); memcpy(p + 2 * i, &t, 2); } break;
        case SC_F32: { float t = val.v[i]; memcpy(p + 4 * i, &t, 4); } break;
        default:     { uint16_t t = f32_to_f16(val.v[i]); memcpy(p + 2 * i, &t, 2); } break;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Preview overlays — src/core/gpu/opencl_undistort.cl:109-154 (draw_pixel, draw_safe_area; colours / alphas :109-120).
 * DATA_CONVERT is convert_<T>_sat with OpenCL's default round-toward-zero = the truncating, saturating `as` cast.
 * ---------------------------------------------------------------------------------------- */
static const float OVL_COLORS[9][4] = { {0, 0, 0, 0}, {255, 0, 0, 255}, {0, 255, 0, 255}, {0, 0, 255, 255}, {254, 251, 71, 255},
                                        {200, 200, 0, 255}, {255, 0, 255, 255}, {0, 128, 255, 255}, {0, 200, 200, 255} };
static const float OVL_ALPHAS[4] = { 1.0f, 0.75f, 0.50f, 0.25f };
void gf_oracle_draw_overlays(uint8_t* buf, size_t len, int width, int height, int stride, const gf_kernel_params* P, int pixel_type,
                             int is_input, const uint8_t* drawing, size_t drawing_len) {
    int count, scalar;
    if (!pix_layout(pixel_type, &count, &scalar)) return;
    const int sb = scalar == SC_U8 ? 1 : (scalar == SC_F32 ? 4 : 2), bpp = count * sb;
    const int dw = P->width > P->output_width ? P->width : P->output_width;               /* max(params->width, params->output_width) */
    for (int y = 0; y < height; ++y) for (int x = 0; x < width; ++x) {
        const size_t off = (size_t)y * (size_t)stride + (size_t)x * (size_t)bpp;
        if (off + (size_t)bpp > len) continue;
        uint8_t* px = buf + off;
        if ((P->flags & 8) && drawing && drawing_len) {                                      /* draw_pixel :121-140 */
            const float fpos = rs_round(floorf((float)y / P->canvas_scale) * ((float)dw / P->canvas_scale) + floorf((float)x / P->canvas_scale));
            const int32_t pos = rs_f32_as_i32(fpos);
            if (pos >= 0 && (size_t)pos < drawing_len) {
                const uint8_t data = drawing[pos];
                if (data > 0) {
                    const int color = (data & 0xF8) >> 3, alpha = (data & 0x06) >> 1, stage = data & 1;
                    if (((stage == 0 && is_input) || (stage == 1 && !is_input)) && color < 9 && alpha < 4) {
                        v4 v = pix_to_float(px, count, scalar);
                        const float af = OVL_ALPHAS[alpha];
                        for (int c = 0; c < count; ++c) v.v[c] = OVL_COLORS[color][c] * af + v.v[c] * (1.0f - af);
                        pix_from_float(px, v, count, scalar);
                    }
                }
            }
        }
        if (!is_input) {                                                                     /* draw_safe_area :141-154 */
            const float fx = (float)x, fy = (float)y;
            const float* r = P->safe_area_rect;
            const int safe = fx >= r[0] && fx <= r[2] && fy >= r[1] && fy <= r[3];
            if (!safe) {
                static const float factor[4] = { 0.5f, 0.5f, 0.5f, 1.0f };
                v4 v = pix_to_float(px, count, scalar);
                for (int c = 0; c < count; ++c) v.v[c] = v.v[c] * factor[c];
                pix_from_float(px, v, count, scalar);
                const int border = fx >= r[0] - 5.0f && fx <= r[2] + 5.0f && fy >= r[1] - 5.0f && fy <= r[3] + 5.0f;
                if (border) {
                    v = pix_to_float(px, count, scalar);
                    for (int c = 0; c < count; ++c) v.v[c] = v.v[c] * factor[c];
                    pix_from_float(px, v, count, scalar);
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * sample_input_at — cpu_undistort.rs:329-419, separable branch (I <= 8) :370-412.
 * Bilinear weights: COEFFS[0..64] (cpu_undistort.rs:14-19) are exactly (1 - i/32, i/32).
 * ---------------------------------------------------------------------------------------- */
#include "gf_coeffs.inc"   /* generated by oracle/gen_coeffs.py: bicubic + lanczos4 5-bit tables */

static inline const float* coeff_row(int I, uint32_t frac, float* tmp) {
    if (I == 2) { tmp[0] = 1.0f - (float)frac / 32.0f; tmp[1] = (float)frac / 32.0f; return tmp; }
    if (I == 4) return &GF_COEFFS_BICUBIC[frac << 2];
    return &GF_COEFFS_LANCZOS4[frac << 3];
}

/* EWA (Elliptical Weighted Average) CubicBC helpers — cpu_undistort.rs:271-327.  jac = (dx/du, dx/dv, dy/du, dy/dv) as v4. */
static v2 affine_bbox(v4 jac) {                                                                  /* :272-277 */
    v2 r;
    r.x = 2.0f * rs_max(rs_max(fabsf(jac.v[0] + jac.v[1]), fabsf(jac.v[0] - jac.v[1])), 1.0f);
    r.y = 2.0f * rs_max(rs_max(fabsf(jac.v[2] + jac.v[3]), fabsf(jac.v[2] - jac.v[3])), 1.0f);
    return r;
}
static void clamped_ellipse(v4 jac, float* oa, float* ob, float* oc) {                           /* :279-315 */
    const float jx = jac.v[0], jy = jac.v[1], jz = jac.v[2], jw = jac.v[3];
    float f0 = fabsf(jx * jw - jy * jz);
    float f = rs_max(f0 * f0, 0.1f);
    float a = (jz * jz + jw * jw) / f;
    float b = -2.0f * (jx * jz + jy * jw) / f;
    float c = (jx * jx + jy * jy) / f;
    float vx = c - a, vy = -b;
    float lv = sqrtf(vx * vx + vy * vy);                                                         /* nalgebra norm(): sqrt of the dot product */
    float v0 = lv > 0.01f ? vx / lv : 1.0f;
    float cc = sqrtf(rs_max(1.0f + v0, 0.0f) / 2.0f);
    float s  = sqrtf(rs_max(1.0f - v0, 0.0f) / 2.0f);
    float a0 = a * cc * cc - b * cc * s + c * s * s;
    float c0 = a * s * s + b * cc * s + c * cc * cc;
    float bt1 = b * (cc * cc - s * s);
    float bt2 = 2.0f * (a - c) * cc * s;
    float b0 = bt1 + bt2;
    float b0v2 = bt1 - bt2;
    if (fabsf(b0) > fabsf(b0v2)) { s = -s; b0 = b0v2; }
    a0 = rs_min(a0, 1.0f);
    c0 = rs_min(c0, 1.0f);
    float sn = -s;
    *oa = a0 * cc * cc - b0 * cc * sn + c0 * sn * sn;
    *ob = 2.0f * a0 * cc * sn + b0 * cc * cc - b0 * sn * sn - 2.0f * c0 * cc * sn;
    *oc = a0 * sn * sn + b0 * cc * sn + c0 * cc * cc;
}
static float bc2(float x, const gf_kernel_params* P) {                                           /* :316-326 */
    x = fabsf(x);
    float x2 = x * x;
    if (x < 1.0f) return P->ewa_coeffs_p[0] + P->ewa_coeffs_p[1] * x + P->ewa_coeffs_p[2] * x2 + P->ewa_coeffs_p[3] * x2 * x;
    if (x < 2.0f) return P->ewa_coeffs_q[0] + P->ewa_coeffs_q[1] * x + P->ewa_coeffs_q[2] * x2 + P->ewa_coeffs_q[3] * x2 * x;
    return 0.0f;
}

static v4 sample_input_at(int I, v2 uv, v4 jac, const uint8_t* input, size_t in_len, const gf_kernel_params* P, v4 bg, int count, int scalar, int* oob) {
    v4 sum = {{0.0f, 0.0f, 0.0f, 0.0f}};
    if (I > 8) {                                                                                  /* :331-369 */
        v2 trans_size = affine_bbox(jac);
        int32_t b0 = rs_f32_as_i32(floorf(uv.x - trans_size.x)), b1 = rs_f32_as_i32(ceilf(uv.x + trans_size.x));
        int32_t b2 = rs_f32_as_i32(floorf(uv.y - trans_size.y)), b3 = rs_f32_as_i32(ceilf(uv.y + trans_size.y));
        float sum_div = 0.0f;
        int64_t src_index = (int64_t)(int32_t)((uint32_t)b2 * (uint32_t)P->stride);               /* i32 multiply, wrapping in release builds */
        float ea, eb, ec;
        clamped_ellipse(jac, &ea, &eb, &ec);
        for (int64_t in_y = b2; in_y <= (int64_t)b3; ++in_y) {
            float in_fy = (float)(int32_t)in_y - uv.y;
            float in_fy2 = in_fy * eb;
            float in_fy3 = in_fy * in_fy * ec;
            for (int64_t in_x = b0; in_x <= (int64_t)b1; ++in_x) {
                float in_fx = (float)(int32_t)in_x - uv.x;
                float dr = in_fx * in_fx * ea + in_fx * in_fy2 + in_fy3;
                float k = bc2(sqrtf(dr), P);
                if (k == 0.0f) continue;
                v4 pixel;
                if (in_y >= P->source_rect[1] && in_y < P->source_rect[1] + P->source_rect[3] && in_x >= P->source_rect[0] && in_x < P->source_rect[0] + P->source_rect[2]) {
                    int64_t off = src_index + (int64_t)P->bytes_per_pixel * in_x;
                    if (off < 0 || (uint64_t)off + (uint64_t)P->bytes_per_pixel > in_len) { *oob = 1; pixel = bg; }
                    else pixel = pix_to_float(input + off, count, scalar);
                } else {
                    pixel = bg;
                }
                for (int ch = 0; ch < 4; ++ch) sum.v[ch] += k * pixel.v[ch];
                sum_div += k;
            }
            src_index += P->stride;
        }
        for (int ch = 0; ch < 4; ++ch) sum.v[ch] /= sum_div;
        for (int ch = 0; ch < 4; ++ch) sum.v[ch] = rs_min(sum.v[ch], P->pixel_value_limit);
        return sum;
    }
    const float offset = (I == 2) ? 0.0f : (I == 4 ? 1.0f : 3.0f);
    float u = uv.x - offset, v = uv.y - offset;
    int32_t sx0 = rs_f32_as_i32(rs_round(u * 32.0f));
    int32_t sy0 = rs_f32_as_i32(rs_round(v * 32.0f));
    int32_t sx = sx0 >> 5, sy = sy0 >> 5;
    float tx[2], ty[2];
    const float* coeffs_x = coeff_row(I, (uint32_t)sx0 & 31u, tx);
    const float* coeffs_y = coeff_row(I, (uint32_t)sy0 & 31u, ty);
    int64_t src_index = (int64_t)sy * (int64_t)P->stride + (int64_t)sx * (int64_t)P->bytes_per_pixel;
    for (int yp = 0; yp < I; ++yp) {
        if (sy + yp >= P->source_rect[1] && sy + yp < P->source_rect[1] + P->source_rect[3]) {
            v4 xsum = {{0.0f, 0.0f, 0.0f, 0.0f}};
            for (int xp = 0; xp < I; ++xp) {
                v4 pixel;
                if (sx + xp >= P->source_rect[0] && sx + xp < P->source_rect[0] + P->source_rect[2]) {
                    int64_t off = src_index + (int64_t)P->bytes_per_pixel * xp;
                    if (off < 0 || (uint64_t)off + (uint64_t)P->bytes_per_pixel > in_len) { *oob = 1; pixel = bg; }   /* Rust: slice index panic */
                    else pixel = pix_to_float(input + off, count, scalar);
                } else {
                    pixel = bg;
                }
                for (int ch = 0; ch < 4; ++ch) xsum.v[ch] += pixel.v[ch] * coeffs_x[xp];
            }
            for (int ch = 0; ch < 4; ++ch) sum.v[ch] += xsum.v[ch] * coeffs_y[yp];
        } else {
            for (int ch = 0; ch < 4; ++ch) sum.v[ch] += bg.v[ch] * coeffs_y[yp];
        }
        src_index += P->stride;
    }
    for (int ch = 0; ch < 4; ++ch) sum.v[ch] = rs_min(sum.v[ch], P->pixel_value_limit);
    return sum;
}

/* cpu_undistort.rs:255-260 */
static void remap_colorrange(v4* px, int is_y) {
    float s = is_y ? 0.85882352f : 0.87843137f;
    for (int ch = 0; ch < 4; ++ch) px->v[ch] *= s;
    px->v[0] += 16.0f;
    px->v[1] += 16.0f;
}

/* ------------------------------------------------------------------------------------------
 * undistort_image_cpu — cpu_undistort.rs:233-633 (main loop 519-633)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* in; size_t in_len;
    uint8_t* out; size_t out_len;
    warp_ctx W;
    int count, scalar, I;
    size_t rows;
    int tid, nthreads;
    int oob;
    size_t* next_row;   /* shared cursor: chunks of ROW_CHUNK rows are claimed dynamically (rayon splits work adaptively too) */
} job_t;

#define ROW_CHUNK 4

static void process_row(job_t* J, size_t y) {
    const gf_kernel_params* P = J->W.P;
    const size_t ostride = (size_t)P->output_stride;
    const size_t bpp = (size_t)P->bytes_per_pixel;
    size_t row_off = y * ostride;
    size_t row_len = J->out_len - row_off; if (row_len > ostride) row_len = ostride;
    size_t npix = row_len / bpp;   /* a trailing partial chunk cannot hold a pixel */
    uint8_t* row = J->out + row_off;

    v4 bg; for (int ch = 0; ch < 4; ++ch) bg.v[ch] = P->background[ch] * P->max_pixel_value;   /* :523 */
    const float factor = rs_max(1.0f - P->lens_correction_amount, 0.001f);                       /* :526 */
    const v2 out_c = {(float)P->output_width / 2.0f, (float)P->output_height / 2.0f};            /* :527 */
    const v2 out_f = {P->f[0] / P->fov / factor, P->f[1] / P->fov / factor};                     /* :528 */
    const int fill_bg = (P->flags & 4) == 4, fix_range = (P->flags & 1) == 1, is_y = P->plane_index == 0;

    for (size_t x = 0; x < npix; ++x) {
        float opx = map_coord((float)x, (float)P->output_rect[0], (float)(P->output_rect[0] + P->output_rect[2]), 0.0f, (float)P->output_width);
        float opy = map_coord((float)y, (float)P->output_rect[1], (float)(P->output_rect[1] + P->output_rect[3]), 0.0f, (float)P->output_height);
        if (opx >= 0.0f && opy >= 0.0f && rs_f32_as_i32(opx) < P->output_width && rs_f32_as_i32(opy) < P->output_height) {   /* :551 */
            v4 pixel = bg;
            uint8_t* pix_out = row + x * bpp;
            if (fill_bg) { pix_from_float(pix_out, bg, J->count, J->scalar); continue; }          /* :558-561 */
            v2 position = {(float)x, (float)y};
            v2 uv;
            if (undistort_coord(position, &J->W, out_c, out_f, &uv)) {                            /* :565 */
                v4 jac = {{1.0f, 0.0f, 0.0f, 1.0f}};
                if (J->I > 8) {                                                                    /* :567-572 */
                    const float eps = 0.01f;
                    v2 px = {position.x + eps, position.y}, py = {position.x, position.y + eps}, rx = {0.0f, 0.0f}, ry = {0.0f, 0.0f};
                    if (!undistort_coord(px, &J->W, out_c, out_f, &rx)) { rx.x = 0.0f; rx.y = 0.0f; }    /* unwrap_or_default() */
                    if (!undistort_coord(py, &J->W, out_c, out_f, &ry)) { ry.x = 0.0f; ry.y = 0.0f; }
                    v2 xyx = {rx.x - uv.x, rx.y - uv.y}, xyy = {ry.x - uv.x, ry.y - uv.y};
                    jac.v[0] = xyx.x / eps; jac.v[1] = xyy.x / eps; jac.v[2] = xyx.y / eps; jac.v[3] = xyy.y / eps;
                }
                float width_f = (float)P->width, height_f = (float)P->height;
                if (P->background_mode == 3) {                                                     /* :576-613 */
                    float widthf = width_f - 1.0f, heightf = height_f - 1.0f;
                    float feather = rs_max(P->background_margin_feather * heightf, 0.0001f);
                    v2 pt2 = uv;
                    float alpha = 1.0f;
                    if ((uv.x > widthf - feather) || (uv.x < feather) || (uv.y > heightf - feather) || (uv.y < feather)) {
                        alpha = rs_max(rs_min(rs_min(rs_min(rs_min(widthf - uv.x, heightf - uv.y), uv.x), uv.y) / feather, 1.0f), 0.0f);
                        pt2.x = pt2.x / width_f; pt2.y = pt2.y / height_f;
                        pt2.x = ((pt2.x - 0.5f) * (1.0f - P->background_margin)) + 0.5f;
                        pt2.y = ((pt2.y - 0.5f) * (1.0f - P->background_margin)) + 0.5f;
                        pt2.x = pt2.x * width_f; pt2.y = pt2.y * height_f;
                    }
                    v2 frame_size = {(float)P->width, (float)P->height};
                    if (P->input_rotation != 0.0f) {
                        float rotation = P->input_rotation * (RS_PI_F / 180.0f);
                        v2 size = frame_size;
                        frame_size = rotate_point(size, rotation, (v2){0.0f, 0.0f}, (v2){0.0f, 0.0f});
                        frame_size.x = rs_round(fabsf(frame_size.x)); frame_size.y = rs_round(fabsf(frame_size.y));
                    }
                    float sx0 = (float)P->source_rect[0], sx1 = (float)(P->source_rect[0] + P->source_rect[2]);
                    float sy0 = (float)P->source_rect[1], sy1 = (float)(P->source_rect[1] + P->source_rect[3]);
                    uv.x  = map_coord(uv.x,  0.0f, frame_size.x, sx0, sx1); uv.y  = map_coord(uv.y,  0.0f, frame_size.y, sy0, sy1);
                    pt2.x = map_coord(pt2.x, 0.0f, frame_size.x, sx0, sx1); pt2.y = map_coord(pt2.y, 0.0f, frame_size.y, sy0, sy1);
                    v4 c1 = sample_input_at(J->I, uv,  jac, J->in, J->in_len, P, bg, J->count, J->scalar, &J->oob);
                    v4 c2 = sample_input_at(J->I, pt2, jac, J->in, J->in_len, P, bg, J->count, J->scalar, &J->oob);
                    for (int ch = 0; ch < 4; ++ch) pixel.v[ch] = c1.v[ch] * alpha + c2.v[ch] * (1.0f - alpha);
                    if (fix_range) remap_colorrange(&pixel, is_y);
                    pix_from_float(pix_out, pixel, J->count, J->scalar);
                    continue;
                }
                pixel = sample_input_at(J->I, uv, jac, J->in, J->in_len, P, bg, J->count, J->scalar, &J->oob);   /* :615 */
            }
            if (fix_range) remap_colorrange(&pixel, is_y);                                         /* :619-621 */
            pix_from_float(pix_out, pixel, J->count, J->scalar);                                   /* :622 */
        }
    }
}

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    /* rows are claimed in chunks of ROW_CHUNK from a shared cursor (rayon par_chunks_mut is row-granular, :543) */
    for (;;) {
        size_t base = __atomic_fetch_add(J->next_row, (size_t)ROW_CHUNK, __ATOMIC_RELAXED);
        if (base >= J->rows) break;
        size_t end = base + ROW_CHUNK; if (end > J->rows) end = J->rows;
        for (size_t y = base; y < end; ++y) process_row(J, y);
    }
    return NULL;
}

/* usable hardware threads: online CPUs, capped by the cgroup v2 CPU quota (cpu.max) when the container has one —
 * 128 runnable threads on a 16-CPU quota only get throttled. */
int gf_oracle_online_cpus(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r");
    if (f) {
        char q[64]; long period = 0;
        if (fscanf(f, "%63s %ld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) {
            long quota = atol(q);
            long cpus = (quota + period - 1) / period;
            if (cpus >= 1 && cpus < n) n = cpus;
        }
        fclose(f);
    }
    return (int)n;
}
const char* gf_oracle_describe(void) {
    return "gyroflow CPU oracle (C restatement of cpu_undistort.rs @ b5e8828; libm transcendentals; parity unpinned)";
}

int gf_oracle_undistort_image(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len,
                              const gf_kernel_params* P, int pixel_type, int distortion_model, int digital_lens,
                              const float* matrices, size_t matrix_rows, const float* mesh, size_t mesh_len, int threads) {
    int count, scalar;
    if (!in || !out || !P || !matrices) return GF_ERR_BAD_PARAMS;
    if (!pix_layout(pixel_type, &count, &scalar)) return GF_ERR_BAD_PARAMS;
    static const int scalar_bytes[4] = {1, 2, 4, 2};
    if (P->bytes_per_pixel != count * scalar_bytes[scalar]) return GF_ERR_BAD_PARAMS;          /* assert_eq! :541 */
    if (P->output_stride <= 0 || P->stride <= 0) return GF_ERR_BAD_STRIDE;                      /* :534-537 */
    if (P->matrix_count < 1 || (size_t)P->matrix_count > matrix_rows) return GF_ERR_BAD_PARAMS; /* matrices[idx] would panic */
    const int I = P->interpolation;
    if (!(I == 2 || I == 4 || I == 8 || (I >= 10 && I <= 13))) return GF_ERR_UNSUPPORTED_COMBO;  /* Interpolation enum, stabilization/mod.rs:24-34 */
    if (distortion_model <= GF_LENS_NONE || distortion_model >= GF_LENS_COUNT) return GF_ERR_BAD_PARAMS;
    if (digital_lens < 0 || digital_lens >= GF_LENS_COUNT) return GF_ERR_BAD_PARAMS;

    double* mesh64 = NULL;
    if (mesh && mesh_len) {
        mesh64 = (double*)malloc(mesh_len * sizeof(double));
        if (!mesh64) return GF_ERR_BAD_PARAMS;
        for (size_t i = 0; i < mesh_len; ++i) mesh64[i] = (double)mesh[i];                       /* :539 */
    }

    int nthreads = threads > 0 ? threads : gf_oracle_online_cpus();
    size_t rows = (out_len + (size_t)P->output_stride - 1) / (size_t)P->output_stride;
    if ((size_t)nthreads > (rows + ROW_CHUNK - 1) / ROW_CHUNK) nthreads = (int)((rows + ROW_CHUNK - 1) / ROW_CHUNK);
    if (nthreads < 1) nthreads = 1;

    size_t next_row = 0;
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; ++t) {
        job_t* J = &jobs[t];
        J->in = in; J->in_len = in_len; J->out = out; J->out_len = out_len;
        J->W.P = P; J->W.matrices = matrices; J->W.matrix_rows = matrix_rows;
        J->W.model = distortion_model; J->W.digital = digital_lens;
        J->W.r_limit_sq = P->r_limit * P->r_limit;                                                /* :521 */
        J->W.mesh = mesh64; J->W.mesh_len = mesh64 ? mesh_len : 0;
        J->count = count; J->scalar = scalar; J->I = I; J->rows = rows; J->tid = t; J->nthreads = nthreads; J->next_row = &next_row;
    }
    if (nthreads == 1) worker(&jobs[0]);
    else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    int oob = 0; for (int t = 0; t < nthreads; ++t) oob |= jobs[t].oob;
    free(jobs); free(th); free(mesh64);
    return oob ? GF_ERR_BUFFER_TOO_SMALL : GF_OK;
}

int gf_oracle_rotate_and_distort(float x, float y, size_t idx, const gf_kernel_params* P, const float* matrices,
                                 int distortion_model, int digital_lens, const float* mesh, size_t mesh_len, float* ou, float* ov) {
    double mesh64[GF_MESH_MAX_LEN];
    if (mesh_len > GF_MESH_MAX_LEN) mesh_len = GF_MESH_MAX_LEN;
    for (size_t i = 0; i < mesh_len; ++i) mesh64[i] = (double)mesh[i];
    warp_ctx W = { P, matrices, (size_t)P->matrix_count, distortion_model, digital_lens, P->r_limit * P->r_limit, mesh64, mesh ? mesh_len : 0 };
    v2 o = {0, 0};
    int r = rotate_and_distort((v2){x, y}, idx, &W, &o);
    *ou = o.x; *ov = o.y; return r;
}

int gf_oracle_undistort_coord(float x, float y, const gf_kernel_params* P, const float* matrices,
                              int distortion_model, int digital_lens, const float* mesh, size_t mesh_len, float* ou, float* ov) {
    double mesh64[GF_MESH_MAX_LEN];
    if (mesh_len > GF_MESH_MAX_LEN) mesh_len = GF_MESH_MAX_LEN;
    for (size_t i = 0; i < mesh_len; ++i) mesh64[i] = (double)mesh[i];
    warp_ctx W = { P, matrices, (size_t)P->matrix_count, distortion_model, digital_lens, P->r_limit * P->r_limit, mesh64, mesh ? mesh_len : 0 };
    const float factor = rs_max(1.0f - P->lens_correction_amount, 0.001f);
    const v2 out_c = {(float)P->output_width / 2.0f, (float)P->output_height / 2.0f};
    const v2 out_f = {P->f[0] / P->fov / factor, P->f[1] / P->fov / factor};
    v2 o = {0, 0};
    int r = undistort_coord((v2){x, y}, &W, out_c, out_f, &o);
    *ou = o.x; *ov = o.y; return r;
}

/* ==========================================================================================
 * Adaptive-zoom companion — zooming/fov_iterative.rs, cpu_undistort.rs:636-858, frame_transform.rs:352-438
 * (f64 quaternion algebra as nalgebra 0.34.2; quat_at_timestamp gyro_source/mod.rs:857-879)
 * ======================================================================================== */
typedef struct { double w, i, j, k; } quat;
static quat q_mul(quat a, quat b) {
    quat r = { a.w*b.w - a.i*b.i - a.j*b.j - a.k*b.k, a.w*b.i + a.i*b.w + a.j*b.k - a.k*b.j,
               a.w*b.j - a.i*b.k + a.j*b.w + a.k*b.i, a.w*b.k + a.i*b.j - a.j*b.i + a.k*b.w };
    return r;
}
static quat q_slerp(quat a, quat b, double t) {
    double d = a.w*b.w + a.i*b.i + a.j*b.j + a.k*b.k;
    if (d < 0.0) { b.w = -b.w; b.i = -b.i; b.j = -b.j; b.k = -b.k; d = -d; }
    if (d >= 1.0) return a;
    double hang = acos(d), s = sqrt(1.0 - d*d);
    if (fabs(s) < 1e-14) return a;
    double ta = sin((1.0 - t) * hang) / s, tb = sin(t * hang) / s;
    quat r = { a.w*ta + b.w*tb, a.i*ta + b.i*tb, a.j*ta + b.j*tb, a.k*ta + b.k*tb };
    return r;
}
static quat track_get(const gf_quat_track* t, size_t i) { quat q = { t->quats[4*i], t->quats[4*i+1], t->quats[4*i+2], t->quats[4*i+3] }; return q; }
static quat quat_at_timestamp(const gf_quat_track* t, double duration_ms, double timestamp_ms) {     /* gyro_source/mod.rs:857-879 */
    quat id = {1.0, 0.0, 0.0, 0.0};
    if (t->n < 2 || duration_ms <= 0.0) return id;
    int64_t first_ts = t->ts_us[0], last_ts = t->ts_us[t->n - 1];
    int64_t lookup = (int64_t)llround(timestamp_ms * 1000.0);
    if (lookup > last_ts) lookup = last_ts;
    if (lookup < first_ts) lookup = first_ts;
    size_t lo = 0, hi = t->n;
    while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (t->ts_us[mid] <= lookup) lo = mid; else hi = mid; }
    if (t->ts_us[lo] == lookup || lo + 1 >= t->n) return track_get(t, lo);
    double fract = (double)(lookup - t->ts_us[lo]) / (double)(t->ts_us[lo + 1] - t->ts_us[lo]);
    return q_slerp(track_get(t, lo), track_get(t, lo + 1), fract);
}
static void mat3_mul_d(const double* a, const double* b, double* o) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[r*3+c] = a[r*3]*b[c] + a[r*3+1]*b[3+c] + a[r*3+2]*b[6+c];
}

/* frame_transform.rs:52-58 */
static double pts_get_fov(const gf_compute_params* cp, size_t frame, int use_fovs) {
    double fov_scale = cp->fov_scale + ((cp->fov_overview && use_fovs) ? 1.0 : 0.0);
    double fov = 1.0;
    if (use_fovs) {
        double f = 1.0;
        if (frame < cp->n_fovs) f = cp->fovs[frame]; else if (cp->n_fovs > 1) f = cp->fovs[cp->n_fovs - 1];
        fov = f * fov_scale;
    }
    fov = fmax(fov, 0.001);
    return fov * (double)cp->width / (double)(cp->output_width > 1 ? cp->output_width : 1);
}

/* CatmullRom<Vector3<f64>>::interpolate — gyro_source/splines.rs:22-83 (search_lower_cp :57-73, catmull_rom :75-80).
 * pos[n] ascending, val[n][3].  Returns 0 for None. */
static int catmull_rom3(const double* pos, const double* val, size_t n, double t, double out[3]) {
    if (n < 2 || t != t) return 0;
    size_t lo = 0, hi = n;                                           /* binary_search_by: first index with pos >= t */
    while (lo < hi) { size_t mid = lo + (hi - lo) / 2; if (pos[mid] < t) lo = mid + 1; else hi = mid; }
    size_t lower;
    if (lo < n && pos[lo] == t) { if (lo == n - 1) return 0; lower = lo; }          /* Ok(i) */
    else { if (lo >= n || lo == 0) return 0; lower = lo - 1; }                     /* Err(i) */
    if (lower + 1 >= n) return 0;
    const double k = (t - pos[lower]) / (pos[lower + 1] - pos[lower]);              /* normalize */
    for (int c = 0; c < 3; ++c) {
        const double a = val[3 * lower + c], b = val[3 * (lower + 1) + c];
        const double x = lower == 0 ? a * 2.0 - b : val[3 * (lower - 1) + c];
        const double y = lower + 2 >= n ? b * 2.0 - a : val[3 * (lower + 2) + c];
        out[c] = ((((a * 3.0 - x) - b * 3.0) + y) * 0.5) * k * k * k + ((b - x) * 0.5) * k + a + (((b * 4.0 + a * -5.0 + x + x) - y) * 0.5) * k * k;
    }
    return 1;
}
/* the `shifts` of at_timestamp_for_points — frame_transform.rs:412-434: (sx, sy, ra, ox, oy) per element of points_iter (ONE element,
 * the point (0, 0), when rolling-shutter correction is off).  Returns the number of entries written (0 = None). */
static size_t shifts_for_points(const gf_compute_params* cp, const float* pts, size_t n, size_t frame, float* shifts /* n x 5 */) {
    if (!cp->camera_stab || frame >= cp->n_camera_stab) return 0;
    if (cp->suppress_rotation && cp->frame_readout_time == 0.0) return 0;                                   /* :432-434 */
    const gf_camera_stab* is = &cp->camera_stab[frame];
    double frt = fabs(cp->frame_readout_time);
    const size_t cnt = frt > 0.0 ? n : 1;
    const double scale_x = (double)cp->width  / (double)is->crop_area[2] / (double)is->pixel_pitch[0];
    const double scale_y = (double)cp->height / (double)is->crop_area[3] / (double)is->pixel_pitch[1];
    for (size_t p = 0; p < cnt; ++p) {
        const double py = frt > 0.0 ? (double)pts[2 * p + 1] : 0.0;
        const double out_min = (double)is->crop_area[1], out_max = (double)is->crop_area[1] + (double)is->crop_area[3];
        const double y = (py - 0.0) * (out_max - out_min) / ((double)cp->height - 0.0) + out_min;         /* map_coord, f64 */
        double s[3] = {0.0, 0.0, 0.0}, o[3] = {0.0, 0.0, 0.0};
        if (!catmull_rom3(is->ibis_pos, is->ibis_xyz, is->n_ibis, y + is->offset, s)) { s[0] = s[1] = s[2] = 0.0; }
        if (!catmull_rom3(is->ois_pos, is->ois_xyz, is->n_ois, y + is->offset, o)) { o[0] = o[1] = o[2] = 0.0; }
        const double ra = s[2] / 1000.0;
        float* d = shifts + 5 * p;
        d[0] = (float)(s[0] * scale_x); d[1] = (float)(s[1] * scale_y); d[2] = (float)(ra * (M_PI / 180.0));
        d[3] = (float)(o[0] * scale_x); d[4] = (float)(o[1] * scale_y);
    }
    return cnt;
}

/* at_timestamp_for_points — frame_transform.rs:352-438: per-point K_new * R (f64), use_fovs = false */
static void rotations_for_points(const gf_compute_params* cp, const float* pts, size_t n, double timestamp_ms, size_t frame, int use_fovs,
                                 double* rot /* n x 9 */, double* fov_out) {
    double fov = pts_get_fov(cp, frame, use_fovs);
    const double* K = cp->camera_matrix;
    double hr = cp->input_horizontal_stretch > 0.01 ? cp->input_horizontal_stretch : 1.0;
    double new_k[9]; memcpy(new_k, K, sizeof(new_k));
    new_k[0] = new_k[0] * (1.0 / hr) / fov; new_k[4] = new_k[4] * (1.0 / hr) / fov;
    new_k[2] = (double)cp->output_width / 2.0; new_k[5] = (double)cp->output_height / 2.0;
    double frt = fabs(cp->frame_readout_time);                       /* get_frame_readout_time(can_invert = false) :376 */
    if (cp->readout_inverted) frt *= -1.0;
    double row_readout_time = frt / (double)(cp->readout_horizontal ? cp->width : cp->height);
    double start_ts = timestamp_ms - frt / 2.0;
    double a = cp->video_rotation * (M_PI / 180.0);
    double rz[9] = { cos(a), -sin(a), 0.0, sin(a), cos(a), 0.0, 0.0, 0.0, 1.0 };
    quat q1 = quat_at_timestamp(&cp->org, cp->duration_ms, timestamp_ms - cp->gyro_offset_ms);
    q1.i = -q1.i; q1.j = -q1.j; q1.k = -q1.k;
    quat sq1 = quat_at_timestamp(&cp->smoothed, cp->duration_ms, timestamp_ms - cp->gyro_offset_ms);
    quat q0 = q_mul(sq1, q1);
    size_t cnt = fabs(frt) > 0.0 ? n : 1;                            /* :389 */
    for (size_t p = 0; p < cnt; ++p) {
        double px = fabs(frt) > 0.0 ? (double)pts[2*p] : 0.0, py = fabs(frt) > 0.0 ? (double)pts[2*p+1] : 0.0;
        double quat_time = fabs(frt) > 0.0 ? start_ts + row_readout_time * (cp->readout_horizontal ? px : py) : start_ts;
        quat q = q_mul(q0, quat_at_timestamp(&cp->org, cp->duration_ms, quat_time - cp->gyro_offset_ms));
        double ww = q.w*q.w, ii = q.i*q.i, jj = q.j*q.j, kk = q.k*q.k;
        double ij = q.i*q.j*2.0, wk = q.w*q.k*2.0, wj = q.w*q.j*2.0, ik = q.i*q.k*2.0, jk = q.j*q.k*2.0, wi = q.w*q.i*2.0;
        double rq[9] = { ww+ii-jj-kk, ij-wk, wj+ik, wk+ij, ww-ii+jj-kk, jk-wi, ik-wj, wi+jk, ww-ii-jj+kk };
        double r[9]; mat3_mul_d(rz, rq, r);
        r[1] *= -1.0; r[2] *= -1.0; r[3] *= -1.0; r[6] *= -1.0;      /* :402-403 */
        if (cp->suppress_rotation) { for (int t = 0; t < 9; ++t) r[t] = (t % 4 == 0) ? 1.0 : 0.0; }
        mat3_mul_d(new_k, r, rot + 9 * p);
    }
    for (size_t p = cnt; p < n; ++p) memcpy(rot + 9 * p, rot, 9 * sizeof(double));   /* rot_per_point.get(index).unwrap_or(&rr) with rr = rotations[0] */
    *fov_out = fov;
}

/* R(o) of cpu_undistort.rs:794-815 */
typedef struct { const gf_kernel_params* kp; int model, digital; float out_cx, out_cy, out_fx, out_fy, fov; } lc_ctx;
static v2 lc_r_of(const lc_ctx* L, v2 o) {
    v2 q = o;
    if (L->digital != GF_LENS_NONE) {
        v2 uz = { (q.x - L->out_cx) * L->fov + L->out_cx, (q.y - L->out_cy) * L->fov + L->out_cy }, d;
        if (lens_undistort(L->digital, uz, L->kp, &d)) { q.x = (d.x - L->out_cx) / L->fov + L->out_cx; q.y = (d.y - L->out_cy) / L->fov + L->out_cy; }
    }
    v2 nn = { (q.x - L->out_cx) / L->out_fx, (q.y - L->out_cy) / L->out_fy }, d;
    if (lens_undistort(L->model, nn, L->kp, &d)) nn = d;
    float lrc = L->kp->light_refraction_coefficient;
    if (lrc != 1.0f && lrc > 0.0f) {
        float r = sqrtf(nn.x * nn.x + nn.y * nn.y);
        if (r != 0.0f) {
            float sin_theta_d = (r / sqrtf(1.0f + r * r)) / lrc;
            float r_d = sin_theta_d / sqrtf(1.0f - sin_theta_d * sin_theta_d);
            float s = r_d / r;
            nn.x = nn.x * s; nn.y = nn.y * s;
        }
    }
    v2 res = { (nn.x * L->out_fx) + L->out_cx, (nn.y * L->out_fy) + L->out_cy };
    return res;
}

/* undistort_points — cpu_undistort.rs:652-858 (mesh = None, shift_per_point = None) */
static void undistort_points(const gf_compute_params* cp, int model, int digital, const float* distorted, size_t n,
                             const double* rot_per_point, double lens_correction_amount, double fov,
                             const float* shifts, size_t n_shifts, size_t frame, float* out) {
    const double* K = cp->camera_matrix;
    const float fx = (float)K[0], fy = (float)K[4], cx = (float)K[2], cy = (float)K[5];
    gf_kernel_params kp; memset(&kp, 0, sizeof(kp));                /* :671-683 */
    kp.width = cp->width; kp.height = cp->height; kp.output_width = cp->output_width; kp.output_height = cp->output_height;
    kp.f[0] = fx; kp.f[1] = fy; kp.c[0] = cx; kp.c[1] = cy;
    for (int i = 0; i < 12; ++i) kp.k[i] = (float)cp->distortion_coeffs[i];
    for (int i = 0; i < 16 && i < cp->n_digital_lens_params; ++i) kp.digital_lens_params[i] = (float)cp->digital_lens_params[i];
    kp.light_refraction_coefficient = (float)cp->light_refraction_coefficient;
    const int lc = lens_correction_amount < 1.0;
    lc_ctx L; memset(&L, 0, sizeof(L));
    float amount = 0.0f, factor = 0.0f;
    if (lc) {                                                        /* :686-694 */
        L.kp = &kp; L.model = model; L.digital = digital;
        L.out_cx = (float)cp->output_width / 2.0f; L.out_cy = (float)cp->output_height / 2.0f;
        amount = (float)lens_correction_amount;
        factor = rs_max(1.0f - amount, 0.001f);
        L.out_fx = fx / (float)fov / factor; L.out_fy = fy / (float)fov / factor;
        L.fov = (float)fov;
    }
    for (size_t idx = 0; idx < n; ++idx) {
        float x = distorted[2*idx], y = distorted[2*idx+1];
        if (cp->input_horizontal_stretch > 0.001) x *= (float)cp->input_horizontal_stretch;     /* :702-703 */
        if (cp->input_vertical_stretch   > 0.001) y *= (float)cp->input_vertical_stretch;
        if (digital != GF_LENS_NONE) { v2 t; if (lens_undistort(digital, (v2){x, y}, &kp, &t)) { x = t.x; y = t.y; } }   /* :705-710 */
        if (cp->distorting_mesh && frame < cp->n_distorting_mesh && cp->distorting_mesh[frame].data && cp->distorting_mesh[frame].len > 9) {   /* :712-746 */
            const double* md = cp->distorting_mesh[frame].data; const size_t mlen = cp->distorting_mesh[frame].len;
            const float fw = (float)cp->width, fh = (float)cp->height;
            const float ox = (float)md[5], oy = (float)md[6], cw = (float)md[7], chh = (float)md[8];
            const size_t o = md[0] > 0.0 ? (size_t)md[0] : 0;
            if (md[0] > 0.0 && o < mlen && md[o] > 0.0) {            /* FocalPlaneDistortion :714-733 (added on this path) */
                const double stblz_grid = md[4] / 8.0;
                x = map_coord(x, 0.0f, fw, ox, ox + cw);
                y = map_coord(y, 0.0f, fh, oy, oy + chh);
                double q = floor((double)y / stblz_grid);
                q = q != q ? 0.0 : (q < 0.0 ? 0.0 : (q > 7.0 ? 7.0 : q));
                const size_t fi = (size_t)q;
                const double delta = (double)y - stblz_grid * (double)fi;
                x += (float)(md[o + 4 + fi * 2 + 0] * delta);
                y += (float)(md[o + 4 + fi * 2 + 1] * delta);
                for (size_t j = 0; j < fi; ++j) { x += (float)(md[o + 4 + j * 2 + 0] * stblz_grid); y += (float)(md[o + 4 + j * 2 + 1] * stblz_grid); }
                x = map_coord(x, ox, ox + cw, 0.0f, fw);
                y = map_coord(y, oy, oy + chh, 0.0f, fh);
            }
            if (md[0] > 10.0) {                                      /* :735-745 */
                x = map_coord(x, 0.0f, fw, ox, ox + cw);
                y = map_coord(y, 0.0f, fh, oy, oy + chh);
                double nx, ny;
                gf_oracle_interpolate_mesh((double)x, (double)y, md, &nx, &ny);
                x = map_coord((float)nx, ox, ox + cw, 0.0f, fw);
                y = map_coord((float)ny, oy, oy + chh, 0.0f, fh);
            }
        }
        if (shifts && idx < n_shifts) {                              /* :748-757 (sic: y is rotated with the UPDATED x) */
            const float* sh = shifts + 5 * idx;
            const float cos_a = cosf(sh[2]), sin_a = sinf(sh[2]);
            x = x - cx - sh[3] + sh[0];
            y = y - cy - sh[4] + sh[1];
            x = cos_a * x - sin_a * y + cx;
            y = sin_a * x + cos_a * y + cy;
        }
        v2 pw = { (x - cx) / fx, (y - cy) / fy };                    /* :762 */
        float rot[9]; for (int t = 0; t < 9; ++t) rot[t] = (float)rot_per_point[9*idx + t];      /* :764 */
        v2 pt;
        if (lens_undistort(model, pw, &kp, &pt)) {
            if (kp.light_refraction_coefficient != 1.0f && kp.light_refraction_coefficient > 0.0f) {   /* :767-776 */
                float r = sqrtf(pt.x * pt.x + pt.y * pt.y);
                if (r != 0.0f) {
                    float sin_theta_d = (r / sqrtf(1.0f + r * r)) / kp.light_refraction_coefficient;
                    float r_d = sin_theta_d / sqrtf(1.0f - sin_theta_d * sin_theta_d);
                    float f2 = r_d / r;
                    pt.x *= f2; pt.y *= f2;
                }
            }
            /* pr = rot * (pt.0, pt.1, 1): nalgebra gemv accumulates column by column */
            float pr0 = rot[0] * pt.x + rot[1] * pt.y + rot[2] * 1.0f;
            float pr1 = rot[3] * pt.x + rot[4] * pt.y + rot[5] * 1.0f;
            float pr2 = rot[6] * pt.x + rot[7] * pt.y + rot[8] * 1.0f;
            pt.x = pr0 / pr2; pt.y = pr1 / pr2;                      /* :780 */
            if (lc) {                                                /* :782-852 */
                v2 nn = { (pt.x - L.out_cx) / L.out_fx, (pt.y - L.out_cy) / L.out_fy };
                v2 d = lens_distort(model, nn.x, nn.y, 1.0f, &kp);
                v2 p2 = { (d.x * L.out_fx) + L.out_cx, (d.y * L.out_fy) + L.out_cy };
                if (digital != GF_LENS_NONE) {
                    v2 uz = { (p2.x - L.out_cx) * L.fov + L.out_cx, (p2.y - L.out_cy) * L.fov + L.out_cy };
                    v2 dd = lens_distort(digital, uz.x, uz.y, 1.0f, &kp);
                    p2.x = (dd.x - L.out_cx) / L.fov + L.out_cx; p2.y = (dd.y - L.out_cy) / L.fov + L.out_cy;
                }
                v2 o = pt;
                if (isfinite(p2.x) && isfinite(p2.y)) { o.x = p2.x * factor + pt.x * amount; o.y = p2.y * factor + pt.y * amount; }
                for (int it = 0; it < 10; ++it) {
                    v2 r = lc_r_of(&L, o);
                    float g0 = amount * o.x + factor * r.x - pt.x, g1 = amount * o.y + factor * r.y - pt.y;
                    if (fabsf(g0) < 0.02f && fabsf(g1) < 0.02f) break;
                    const float eps = 1.0f;
                    v2 rx = lc_r_of(&L, (v2){o.x + eps, o.y}), ry = lc_r_of(&L, (v2){o.x, o.y + eps});
                    float j11 = amount + factor * (rx.x - r.x) / eps, j21 = factor * (rx.y - r.y) / eps;
                    float j12 = factor * (ry.x - r.x) / eps,          j22 = amount + factor * (ry.y - r.y) / eps;
                    float det = j11 * j22 - j12 * j21;
                    if (!isfinite(det) || fabsf(det) < 1e-9f) break;
                    float dx = ( j22 * g0 - j12 * g1) / det, dy = (-j21 * g0 + j11 * g1) / det;
                    if (!isfinite(dx) || !isfinite(dy)) break;
                    o.x = o.x - dx; o.y = o.y - dy;
                }
                pt = o;
            }
            out[2*idx] = pt.x; out[2*idx+1] = pt.y;
        } else {
            out[2*idx] = -1000000.0f; out[2*idx+1] = -1000000.0f;    /* :855 */
        }
    }
}

void gf_oracle_undistort_points_rs_ex(const gf_compute_params* cp, int model, int digital, const float* distorted, size_t n,
                                      double timestamp_ms, size_t frame, double lens_correction_amount, int use_fovs, float* out) {
    if (n == 0) return;
    double* rot = (double*)malloc(n * 9 * sizeof(double));
    double fov;
    rotations_for_points(cp, distorted, n, timestamp_ms, frame, use_fovs, rot, &fov);
    float* shifts = (float*)malloc(n * 5 * sizeof(float));
    const size_t n_shifts = shifts_for_points(cp, distorted, n, frame, shifts);
    undistort_points(cp, model, digital, distorted, n, rot, lens_correction_amount, fov, n_shifts ? shifts : NULL, n_shifts, frame, out);
    free(shifts);
    free(rot);
}
void gf_oracle_undistort_points_rs(const gf_compute_params* cp, int model, int digital, const float* distorted, size_t n,
                                   double timestamp_ms, size_t frame, double lens_correction_amount, float* out) {
    gf_oracle_undistort_points_rs_ex(cp, model, digital, distorted, n, timestamp_ms, frame, lens_correction_amount, 0, out);
}

/* ---- ST maps — stmap.rs:86-136 -------------------------------------------------------------------
 * The two closures of generate_stmaps and the RGB encoding of parallel_exr (:131-135); the size computation of :58-77 is
 * restated by the test around gf_oracle_undistort_points_rs_ex. */
void gf_oracle_stmap_undistort(const gf_kernel_params* P, const float* matrices, int model, int digital, float* out_rgb) {
    const int w = P->width, h = P->height;
    warp_ctx W = { P, matrices, (size_t)P->matrix_count, model, digital, P->r_limit * P->r_limit, NULL, 0 };
    const int horizontal = (P->flags & 16) == 16;
    for (int yi = 0; yi < h; ++yi) for (int xi = 0; xi < w; ++xi) {
        float x = (float)xi, y = (float)yi;
        v2 coords = {0.0f, 0.0f};
        int32_t s0 = horizontal ? rs_f32_as_i32(rs_round(x)) : rs_f32_as_i32(rs_round(y));                  /* :91-95 */
        int32_t lim = horizontal ? P->width : P->height;
        if (s0 > lim) s0 = lim; if (s0 < 0) s0 = 0;
        size_t sy = (size_t)s0;
        if (P->matrix_count > 1) {                                                                            /* :96-105 */
            v2 pt;
            if (rotate_and_distort((v2){x, y}, (size_t)P->matrix_count / 2, &W, &pt)) {
                int32_t s1 = rs_f32_as_i32(rs_round(horizontal ? pt.x : pt.y));
                if (s1 > lim) s1 = lim; if (s1 < 0) s1 = 0;
                sy = (size_t)s1;
            }
        }
        size_t idx = sy < (size_t)P->matrix_count - 1 ? sy : (size_t)P->matrix_count - 1;                    /* :108 */
        v2 r;
        if (rotate_and_distort((v2){x, y}, idx, &W, &r)) coords = r;                                          /* :109, :123-126 */
        float* o = out_rgb + ((size_t)yi * (size_t)w + (size_t)xi) * 3;
        o[0] = coords.x / (float)w; o[1] = 1.0f - (coords.y / (float)h); o[2] = 0.0f;                        /* :131-135 */
    }
}
void gf_oracle_stmap_distort(const gf_compute_params* cp, int model, int digital, double timestamp_ms, size_t frame, float* out_rgb) {
    const int w = cp->width, h = cp->height;
    const size_t n = (size_t)w * (size_t)h;
    float* pts = (float*)malloc(n * 2 * sizeof(float)); float* und = (float*)malloc(n * 2 * sizeof(float));
    for (int yi = 0; yi < h; ++yi) for (int xi = 0; xi < w; ++xi) { pts[2 * ((size_t)yi * w + xi)] = (float)xi; pts[2 * ((size_t)yi * w + xi) + 1] = (float)yi; }
    /* :113-115: one point at a time in the reference; per-point rotations make the batch identical */
    gf_oracle_undistort_points_rs_ex(cp, model, digital, pts, n, timestamp_ms, frame, 1.0, 1, und);
    for (size_t i = 0; i < n; ++i) { out_rgb[3 * i] = und[2 * i] / (float)w; out_rgb[3 * i + 1] = 1.0f - (und[2 * i + 1] / (float)h); out_rgb[3 * i + 2] = 0.0f; }
    free(pts); free(und);
}

/* fov_iterative.rs:136-151 */
static int nearest_edge(const float* poly, size_t n, float cx, float cy, float inv_aspect, float* w, float* h) {
    int idx = -1;
    for (size_t i = 0; i < n; ++i) {
        float ap0 = fabsf(poly[2*i] - cx), ap1 = fabsf(poly[2*i+1] - cy);
        if (ap0 < *w && ap1 < *h) {
            if (ap1 > ap0 * inv_aspect) { *w = ap1 / inv_aspect; *h = ap1; }
            else                        { *w = ap0; *h = ap0 * inv_aspect; }
            idx = (int)i;
        }
    }
    return idx;
}
/* fov_iterative.rs:154-175 */
static size_t points_around_rect(float w, float h, size_t w_div, size_t h_div, float margin, float* out) {
    w -= margin * 2.0f; h -= margin * 2.0f;
    size_t wcnt = (w_div > 2 ? w_div : 2) - 1, hcnt = (h_div > 2 ? h_div : 2) - 1;
    float wstep = w / (float)wcnt, hstep = h / (float)hcnt;
    size_t k = 0;
    for (size_t i = 0; i < wcnt; ++i) { out[2*k] = (float)i * wstep;          out[2*k+1] = 0.0f; ++k; }
    for (size_t i = 0; i < hcnt; ++i) { out[2*k] = w;                         out[2*k+1] = (float)i * hstep; ++k; }
    for (size_t i = 0; i < wcnt; ++i) { out[2*k] = (float)(wcnt - i) * wstep; out[2*k+1] = h; ++k; }
    for (size_t i = 0; i < hcnt; ++i) { out[2*k] = 0.0f;                      out[2*k+1] = (float)(hcnt - i) * hstep; ++k; }
    for (size_t i = 0; i < k; ++i) { out[2*i] += margin; out[2*i+1] += margin; }
    return k;
}

double gf_oracle_find_fov(const gf_compute_params* cp, int model, int digital, int org_ow, int org_oh, float margin,
                          double timestamp_ms, size_t frame) {
    /* FovIterative::new :78-89 */
    float ratio = (float)cp->width / (float)(org_ow > 1 ? org_ow : 1);
    float in_w = (float)cp->width, in_h = (float)cp->height;
    float out_w = (float)org_ow * ratio, out_h = (float)org_oh * ratio;
    float inv_aspect = out_h / out_w;
    float rect[2 * 124], poly[2 * 124], relevant[6], distorted[2 * 64];
    size_t len = points_around_rect(in_w, in_h, 31, 31, margin, rect);         /* :38 */
    float cx = in_w / 2.0f, cy = in_h / 2.0f;                                   /* :40 */
    const double lca = cp->lens_correction_amount;
    gf_oracle_undistort_points_rs(cp, model, digital, rect, len, timestamp_ms, frame, lca, poly);   /* :98 */
    for (size_t i = 0; i < len; ++i) { poly[2*i] -= (float)cp->adaptive_zoom_center_offset[0] * in_w; poly[2*i+1] -= (float)cp->adaptive_zoom_center_offset[1] * in_h; }
    float w = 1000000.0f, h = 1000000.0f * inv_aspect;                          /* :107 */
    size_t plen = len;
    for (int it = 1; it < 5; ++it) {                                            /* :110-131 */
        int idx = nearest_edge(poly, plen, cx, cy, inv_aspect, &w, &h);
        if (idx < 0) break;
        if (len == 0) continue;
        /* NB (reference behaviour kept): idx indexes the *current* polygon but is applied to `rect` */
        /*     and `idx.overflowing_sub(1).0 % len` wraps through usize::MAX: idx = 0 -> (2^64 - 1) % len (= 15 for len = 120), not len - 1 */
        size_t i0 = ((size_t)idx - (size_t)1) % len, i1 = (size_t)idx, i2 = ((size_t)idx + 1) % len;
        relevant[0] = rect[2*i0]; relevant[1] = rect[2*i0+1]; relevant[2] = rect[2*i1]; relevant[3] = rect[2*i1+1]; relevant[4] = rect[2*i2]; relevant[5] = rect[2*i2+1];
        /* interpolate_points(&relevant, 30) :180-189 */
        const size_t steps = 30, d = steps + 1, new_len = d * 3 - steps;
        for (size_t i = 0; i < new_len; ++i) {
            size_t idx1 = i / d, idx2 = idx1 + 1 < 2 ? idx1 + 1 : 2;
            float f = (float)(i % d) / (float)d;
            distorted[2*i]   = relevant[2*idx1]   + f * (relevant[2*idx2]   - relevant[2*idx1]);
            distorted[2*i+1] = relevant[2*idx1+1] + f * (relevant[2*idx2+1] - relevant[2*idx1+1]);
        }
        gf_oracle_undistort_points_rs(cp, model, digital, distorted, new_len, timestamp_ms, frame, lca, poly);
        plen = new_len;
        for (size_t i = 0; i < plen; ++i) { poly[2*i] -= (float)cp->adaptive_zoom_center_offset[0] * in_w; poly[2*i+1] -= (float)cp->adaptive_zoom_center_offset[1] * in_h; }
        nearest_edge(poly, plen, cx, cy, inv_aspect, &w, &h);                   /* :127 (index discarded: re-evaluated at loop top) */
    }
    return (double)(w * 2.0f / out_w);                                          /* :133 */
}

/* zoom_dynamic.rs:177-200 */
static void envelope_follower(const double* a, size_t n, double alpha, double* out) {
    if (n == 0) return;
    double* rev = (double*)malloc(n * sizeof(double));
    double q = a[n - 1];
    for (size_t r = 0; r < n; ++r) { double x = a[n - 1 - r]; q = fmin(x, x * alpha + q * (1.0 - alpha)); rev[r] = q; }
    q = rev[n - 1];
    for (size_t r = 0; r < n; ++r) { double x = rev[n - 1 - r]; q = fmin(x, x * alpha + q * (1.0 - alpha)); out[r] = q; }
    free(rev);
}
void gf_oracle_zoom_dynamic(const double* fov_minimal, size_t n, double window_s, double fps, int method, double* out) {
    if (n == 0) return;
    if (method == 1) {                                               /* :69-75 */
        double a1 = 1.0 - exp(-(1.0 / fps) / window_s), a2 = 1.0 - exp(-(1.0 / fps) / 0.2);
        double* tmp = (double*)malloc(n * sizeof(double));
        envelope_follower(fov_minimal, n, a1, tmp);
        envelope_follower(tmp, n, a2, out);
        free(tmp);
        return;
    }
    /* gaussian: :58-67, helpers :80-126 */
    size_t frames = (size_t)floor(window_s * fps); if (frames % 2 == 0) frames += 1;
    size_t half = frames / 2, padn = n + 2 * half;
    double* pad = (double*)malloc(padn * sizeof(double));
    for (size_t i = 0; i < padn; ++i) pad[i] = i < half ? fov_minimal[0] : (i >= half + n ? fov_minimal[n - 1] : fov_minimal[i - half]);
    double* mn = (double*)malloc(n * sizeof(double));
    for (size_t i = 0; i + frames <= padn; ++i) { double m = pad[i]; for (size_t j = 1; j < frames; ++j) m = fmin(m, pad[i + j]); mn[i] = m; }
    for (size_t i = 0; i < padn; ++i) pad[i] = i < half ? mn[0] : (i >= half + n ? mn[n - 1] : mn[i - half]);
    double* g = (double*)malloc(frames * sizeof(double));
    double std = (double)frames / 6.0, sig2 = 2.0 * std * std, sum = 0.0;
    for (size_t i = 0; i < frames; ++i) { long x = (long)i - (long)half; g[i] = exp(-(double)(x * x) / sig2); sum += g[i]; }
    for (size_t i = 0; i < frames; ++i) g[i] /= sum;
    for (size_t i = 0; i + frames <= padn; ++i) { double s = 0.0; for (size_t j = 0; j < frames; ++j) s += pad[i + j] * g[j]; out[i] = s; }
    free(pad); free(mn); free(g);
}
