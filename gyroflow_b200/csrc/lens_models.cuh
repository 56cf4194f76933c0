// lens_models.cuh — the lens-model "plugins" of the warp, one device function pair per model.
//
// Behavioural source: src/core/stabilization/distortion_models/<model>.rs (gyroflow @ b5e8828).
// Each model M provides
//     bool Lens<M>::undistort(px, py, P, noop, &ox, &oy)   <- DistortionModel::undistort_point (Option -> bool)
//     void Lens<M>::distort(x, y, z, P, noop, &ox, &oy)    <- DistortionModel::distort_point
// `noop` is the model's own "all coefficients are zero -> identity" test (fisheye/sony: k0..k3 == 0, generic
// polynomial: k0..k11 == 0, gopro: k1 == 0), evaluated once per launch on the host instead of once per pixel.
// Operation order inside every expression follows the Rust source so that, with -fmad=false,
// results are bit-identical to the CPU path.  `P` is the 368-byte KernelParams living in the
// kernel's constant parameter bank, so k[]/f[]/c[] reads are uniform-register loads.
#pragma once
#include "gf_math.cuh"
#include "../../include/gyroflow_cuda.h"

namespace gf {

#define GF_DEV __device__ __forceinline__

template <int M> struct Lens;   // primary template intentionally undefined

// ---- opencv_fisheye.rs ---------------------------------------------------------------------
template <> struct Lens<GF_LENS_OPENCV_FISHEYE> {
    static GF_DEV bool undistort(float px, float py, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :12-70
        const float k0 = P.k[0], k1 = P.k[1], k2 = P.k[2], k3 = P.k[3];
        if (noop) { ox = px; oy = py; return true; }
        const float EPS = 1e-6f;
        const float PI_F = 3.14159274101257324f;
        float theta_d = sqrtf(px * px + py * py);
        theta_d = rs_min(rs_max(theta_d, -PI_F), PI_F);
        bool converged = false;
        float theta = theta_d, scale = 0.0f;
        if (fabsf(theta_d) > EPS) {
            theta = 0.0f;
            for (int i = 0; i < 10; ++i) {
                const float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta6 * theta2;
                const float k0_theta2 = k0 * theta2, k1_theta4 = k1 * theta4, k2_theta6 = k2 * theta6, k3_theta8 = k3 * theta8;
                float theta_fix = (theta * (1.0f + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d)
                                / (1.0f + 3.0f * k0_theta2 + 5.0f * k1_theta4 + 7.0f * k2_theta6 + 9.0f * k3_theta8);
                theta_fix = rs_min(rs_max(theta_fix, -0.9f), 0.9f);
                theta = theta - theta_fix;
                if (fabsf(theta_fix) < EPS) { converged = true; break; }
            }
            scale = gf_tanf(theta) / theta_d;
        } else {
            converged = true;
        }
        const bool theta_flipped = (theta_d < 0.0f && theta > 0.0f) || (theta_d > 0.0f && theta < 0.0f);
        if (converged && !theta_flipped) { ox = px * scale; oy = py * scale; return true; }
        return false;
    }
    static GF_DEV void distort(float x, float y, float z, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :72-93
        const float k0 = P.k[0], k1 = P.k[1], k2 = P.k[2], k3 = P.k[3];
        x = x / z; y = y / z;
        if (noop) { ox = x; oy = y; return; }
        const float r = sqrtf(x * x + y * y);
        const float theta = gf_atanf(r);
        const float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
        const float theta_d = theta * (1.0f + k0 * theta2 + k1 * theta4 + k2 * theta6 + k3 * theta8);
        const float scale = r == 0.0f ? 1.0f : theta_d / r;
        ox = x * scale; oy = y * scale;
    }
};

// ---- opencv_standard.rs --------------------------------------------------------------------
template <> struct Lens<GF_LENS_OPENCV_STANDARD> {
    static GF_DEV bool undistort(float px, float py, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :12-30
        const float* k = P.k;
        float x = px, y = py;
        const float x0 = px, y0 = py;
        for (int i = 0; i < 20; ++i) {
            const float r2 = x * x + y * y;
            const float icdist = (1.0f + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1.0f + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0.0f) return false;
            const float delta_x = 2.0f * k[2] * x * y + k[3] * (r2 + 2.0f * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const float delta_y = k[2] * (r2 + 2.0f * y * y) + 2.0f * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - delta_x) * icdist;
            y = (y0 - delta_y) * icdist;
        }
        ox = x; oy = y; return true;
    }
    static GF_DEV void distort(float x, float y, float z, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :32-48
        const float* k = P.k;
        x = x / z; y = y / z;
        const float r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
        const float a1 = 2.0f * x * y, a2 = r2 + 2.0f * x * x, a3 = r2 + 2.0f * y * y;
        const float cdist = 1.0f + k[0] * r2 + k[1] * r4 + k[4] * r6;
        const float icdist2 = 1.0f / (1.0f + k[5] * r2 + k[6] * r4 + k[7] * r6);
        ox = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
        oy = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
    }
};

#define GF_NEWTON_EPS 0.00001f
// ---- poly3.rs ------------------------------------------------------------------------------
template <> struct Lens<GF_LENS_POLY3> {
    static GF_DEV bool undistort(float px, float py, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :14-52
        const float inv_k1 = 1.0f / P.k[0];
        const float rd = sqrtf(px * px + py * py);
        if (rd == 0.0f) return false;
        const float rd_div_k1 = rd * inv_k1;
        float ru = rd;
        for (int i = 0; i < 10; ++i) {
            const float fru = ru * ru * ru + ru * inv_k1 - rd_div_k1;
            if (fru >= -GF_NEWTON_EPS && fru < GF_NEWTON_EPS) break;
            if (i > 5) return false;
            ru = ru - (fru / (3.0f * ru * ru + inv_k1));
        }
        if (ru < 0.0f) return false;
        ru = ru / rd;
        ox = px * ru; oy = py * ru; return true;
    }
    static GF_DEV void distort(float x, float y, float z, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :54-63
        x = x / z; y = y / z;
        const float poly2 = P.k[0] * (x * x + y * y) + 1.0f;
        ox = x * poly2; oy = y * poly2;
    }
};
// ---- poly5.rs ------------------------------------------------------------------------------
template <> struct Lens<GF_LENS_POLY5> {
    static GF_DEV bool undistort(float px, float py, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :14-41
        const float k0 = P.k[0], k1 = P.k[1];
        const float rd = sqrtf(px * px + py * py);
        if (rd == 0.0f) return false;
        float ru = rd;
        for (int i = 0; i < 10; ++i) {
            const float ru2 = ru * ru;
            const float fru = ru * (1.0f + k0 * ru2 + k1 * ru2 * ru2) - rd;
            if (fru >= -GF_NEWTON_EPS && fru < GF_NEWTON_EPS) break;
            if (i > 5) return false;
            ru = ru - (fru / (1.0f + 3.0f * k0 * ru2 + 5.0f * k1 * ru2 * ru2));
        }
        if (ru < 0.0f) return false;
        ru = ru / rd;
        ox = px * ru; oy = py * ru; return true;
    }
    static GF_DEV void distort(float x, float y, float z, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :43-53
        x = x / z; y = y / z;
        const float ru2 = x * x + y * y;
        const float poly4 = 1.0f + P.k[0] * ru2 + P.k[1] * ru2 * ru2;
        ox = x * poly4; oy = y * poly4;
    }
};
// ---- ptlens.rs -----------------------------------------------------------------------------
template <> struct Lens<GF_LENS_PTLENS> {
    static GF_DEV bool undistort(float px, float py, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :14-40
        const float k0 = P.k[0], k1 = P.k[1], k2 = P.k[2];
        const float rd = sqrtf(px * px + py * py);
        if (rd == 0.0f) return false;
        float ru = rd;
        for (int i = 0; i < 10; ++i) {
            const float fru = ru * (k0 * ru * ru * ru + k1 * ru * ru + k2 * ru + 1.0f) - rd;
            if (fru >= -GF_NEWTON_EPS && fru < GF_NEWTON_EPS) break;
            if (i > 5) return false;
            ru = ru - (fru / (4.0f * k0 * ru * ru * ru + 3.0f * k1 * ru * ru + 2.0f * k2 * ru + 1.0f));
        }
        if (ru < 0.0f) return false;
        ru = ru / rd;
        ox = px * ru; oy = py * ru; return true;
    }
    static GF_DEV void distort(float x, float y, float z, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :42-53
        x = x / z; y = y / z;
        const float ru2 = x * x + y * y;
        const float r = sqrtf(ru2);
        const float poly3 = P.k[0] * ru2 * r + P.k[1] * ru2 + P.k[2] * r + 1.0f;
        ox = x * poly3; oy = y * poly3;
    }
};
// ---- insta360.rs ---------------------------------------------------------------------------
template <> struct Lens<GF_LENS_INSTA360> {
    static GF_DEV void distort(float x, float y, float z, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :27-48
        const float k1 = P.k[0], k2 = P.k[1], k3 = P.k[2], p1 = P.k[3], p2 = P.k[4], xi = P.k[5];
        const float len = sqrtf(x * x + y * y + z * z);
        x = (x / len) / ((z / len) + xi);
        y = (y / len) / ((z / len) + xi);
        const float r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
        ox = x * (1.0f + k1 * r2 + k2 * r4 + k3 * r6) + 2.0f * p1 * x * y + p2 * (r2 + 2.0f * x * x);
        oy = y * (1.0f + k1 * r2 + k2 * r4 + k3 * r6) + 2.0f * p2 * x * y + p1 * (r2 + 2.0f * y * y);
    }
    static GF_DEV bool undistort(float ptx, float pty, const gf_kernel_params& P, bool noop, float& ox, float& oy) {       // :10-25
        float px = ptx, py = pty;
        for (int i = 0; i < 200; ++i) {
            float dx, dy;
            distort(px, py, 1.0f, P, noop, dx, dy);
            dx = dx - ptx; dy = dy - pty;
            if (fabsf(dx) < 1e-6f && fabsf(dy) < 1e-6f) break;
            px -= dx; py -= dy;
        }
        ox = px; oy = py; return true;
    }
};
// ---- sony.rs -------------------------------------------------------------------------------
template <> struct Lens<GF_LENS_SONY> {
    static GF_DEV bool undistort(float px, float py, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :10-63
        const float* k = P.k;
        if (noop) { ox = px; oy = py; return true; }
        const float EPS = 1e-6f;
        const float theta_d = sqrtf(px * px + py * py);
        bool converged = false;
        float theta = theta_d, scale = 0.0f;
        if (fabsf(theta_d) > EPS) {
            theta = 0.0f;
            for (int i = 0; i < 10; ++i) {
                const float theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta2 * theta3;
                const float k0 = k[0], k1t = k[1] * theta, k2t = k[2] * theta2, k3t = k[3] * theta3, k4t = k[4] * theta4, k5t = k[5] * theta5;
                const float theta_fix = (theta * (k0 + k1t + k2t + k3t + k4t + k5t) - theta_d)
                                      / (k0 + 2.0f * k1t + 3.0f * k2t + 4.0f * k3t + 5.0f * k4t + 6.0f * k5t);
                theta = theta - theta_fix;
                if (fabsf(theta_fix) < EPS) { converged = true; break; }
            }
            scale = gf_tanf(theta) / theta_d;
        } else {
            converged = true;
        }
        const bool theta_flipped = (theta_d < 0.0f && theta > 0.0f) || (theta_d > 0.0f && theta < 0.0f);
        if (converged && !theta_flipped) { ox = px * scale; oy = py * scale; return true; }
        return false;
    }
    static GF_DEV void distort(float x, float y, float z, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :65-89
        const float* k = P.k;
        x = x / z; y = y / z;
        if (noop) { ox = x; oy = y; return; }
        const float r = sqrtf(x * x + y * y);
        const float theta = gf_atanf(r);
        const float theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta2 * theta3, theta6 = theta3 * theta3;
        const float theta_d = theta * k[0] + theta2 * k[1] + theta3 * k[2] + theta4 * k[3] + theta5 * k[4] + theta6 * k[5];
        const float scale = r == 0.0f ? 1.0f : theta_d / r;
        ox = x * scale; oy = y * scale;
    }
};
// ---- generic_polynomial.rs -----------------------------------------------------------------
template <> struct Lens<GF_LENS_GENERIC_POLYNOMIAL> {
    static GF_DEV bool all_zero(const float* k) {
        bool z = true;
        #pragma unroll
        for (int i = 0; i < 12; ++i) z = z && (k[i] == 0.0f);
        return z;
    }
    static GF_DEV bool undistort(float px, float py, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :18-81
        const float* k = P.k;
        if (noop) { ox = px; oy = py; return true; }
        const float EPS = 1e-6f;
        const float theta_d = sqrtf(px * px + py * py);
        bool converged = false;
        float theta = theta_d, scale = 0.0f;
        if (fabsf(theta_d) > EPS) {
            theta = 0.0f;
            for (int i = 0; i < 10; ++i) {
                const float theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta2 * theta3,
                            theta6 = theta3 * theta3, theta7 = theta3 * theta4, theta8 = theta4 * theta4, theta9 = theta4 * theta5,
                            theta10 = theta5 * theta5, theta11 = theta5 * theta6;
                const float k0 = k[0], k1t = k[1] * theta, k2t = k[2] * theta2, k3t = k[3] * theta3, k4t = k[4] * theta4, k5t = k[5] * theta5,
                            k6t = k[6] * theta6, k7t = k[7] * theta7, k8t = k[8] * theta8, k9t = k[9] * theta9, k10t = k[10] * theta10, k11t = k[11] * theta11;
                const float theta_fix = (theta * (k0 + k1t + k2t + k3t + k4t + k5t + k6t + k7t + k8t + k9t + k10t + k11t) - theta_d)
                                      / (k0 + 2.0f * k1t + 3.0f * k2t + 4.0f * k3t + 5.0f * k4t + 6.0f * k5t + 7.0f * k6t + 8.0f * k7t + 9.0f * k8t + 10.0f * k9t + 11.0f * k10t + 12.0f * k11t);
                theta = theta - theta_fix;
                if (fabsf(theta_fix) < EPS) { converged = true; break; }
            }
            scale = gf_tanf(theta) / theta_d;
        } else {
            converged = true;
        }
        const bool theta_flipped = (theta_d < 0.0f && theta > 0.0f) || (theta_d > 0.0f && theta < 0.0f);
        if (converged && !theta_flipped) { ox = px * scale; oy = py * scale; return true; }
        return false;
    }
    static GF_DEV void distort(float x, float y, float z, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :83-122
        const float* k = P.k;
        x = x / z; y = y / z;
        if (noop) { ox = x; oy = y; return; }
        const float r = sqrtf(x * x + y * y);
        const float theta = gf_atanf(r);
        const float theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta2 * theta3,
                    theta6 = theta3 * theta3, theta7 = theta3 * theta4, theta8 = theta4 * theta4, theta9 = theta4 * theta5,
                    theta10 = theta5 * theta5, theta11 = theta5 * theta6, theta12 = theta6 * theta6;
        const float theta_d = theta * k[0] + theta2 * k[1] + theta3 * k[2] + theta4 * k[3] + theta5 * k[4] + theta6 * k[5]
                            + theta7 * k[6] + theta8 * k[7] + theta9 * k[8] + theta10 * k[9] + theta11 * k[10] + theta12 * k[11];
        const float scale = r == 0.0f ? 1.0f : theta_d / r;
        ox = x * scale; oy = y * scale;
    }
};
// ---- gopro.rs ------------------------------------------------------------------------------
template <> struct Lens<GF_LENS_GOPRO> {
    static GF_DEV float poly_eval(float p, const float* k) {    // :19-21
        return k[0] + p * (k[1] + p * (k[2] + p * (k[3] + p * (k[4] + p * (k[5] + p * k[6])))));
    }
    static GF_DEV float poly_deriv(float p, const float* k) {   // :22-24
        return k[1] + p * (2.0f * k[2] + p * (3.0f * k[3] + p * (4.0f * k[4] + p * (5.0f * k[5] + p * (6.0f * k[6])))));
    }
    static GF_DEV float poly_invert(float theta, const float* k) {   // :26-36
        float p = (theta - k[0]) / k[1];
        for (int i = 0; i < 10; ++i) {
            const float d = poly_deriv(p, k);
            if (fabsf(d) < 1e-12f) break;
            const float fix = (poly_eval(p, k) - theta) / d;
            p -= fix;
            if (fabsf(fix) < 1e-7f) break;
        }
        return p;
    }
    static GF_DEV bool undistort(float ptx, float pty, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :42-57
        const float* k = P.k;
        if (noop) { ox = ptx; oy = pty; return true; }
        const float r_norm = sqrtf(ptx * ptx + pty * pty);
        if (r_norm < 1e-9f) { ox = ptx; oy = pty; return true; }
        const float p = r_norm / k[1];
        const float theta = poly_eval(p, k);
        const float TMAX = 1.5533f;
        const float tt = gf_tanf(TMAX);
        const float rr = theta < TMAX ? gf_tanf(theta) : tt + (theta - TMAX) * (1.0f + tt * tt);
        const float scale = rr / r_norm;
        ox = ptx * scale; oy = pty * scale; return true;
    }
    static GF_DEV void distort(float x, float y, float z, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // :61-74
        const float* k = P.k;
        const float posx = x / z, posy = y / z;
        if (noop) { ox = posx; oy = posy; return; }
        const float r = sqrtf(posx * posx + posy * posy);
        const float TMAX = 1.5533f;
        const float tt = gf_tanf(TMAX);
        const float theta = r < tt ? gf_atanf(r) : TMAX + (r - tt) / (1.0f + tt * tt);
        const float p = poly_invert(theta, k);
        const float r_norm = k[1] * p;
        const float scale = r < 1e-9f ? 1.0f : r_norm / r;
        ox = posx * scale; oy = posy * scale;
    }
};

// ================================= digital lenses ==========================================
struct SuperviewFn {     // gopro_superview.rs:12-19
    static GF_DEV void map(float& ux, float& uy, const float*) {
        const float x2 = ux * ux, y2 = uy * uy;
        const float nx = ux * (1.2100393f + x2 * (-1.2758402f + x2 * 1.7751845f));
        const float ny = uy * (0.9364505f + (0.4465308f - 0.7683315f * y2) * y2 + (-0.3574087f + 1.1584653f * y2 + 0.3529348f * x2) * x2);
        ux = nx; uy = ny;
    }
};
struct Superview6Fn {    // gopro6_superview.rs:12-17
    static GF_DEV void map(float& ux, float& uy, const float*) {
        ux *= 1.0f - 0.48f * fabsf(ux);
        ux *= 0.943396f * (1.0f + 0.157895f * fabsf(ux));
        uy *= 0.943396f * (1.0f + 0.060000f * fabsf(uy * 2.0f));
    }
};
struct HyperviewFn {     // gopro_hyperview.rs:10-17
    static GF_DEV void map(float& ux, float& uy, const float*) {
        const float x2 = ux * ux, y2 = uy * uy;
        const float nx = ux * (1.5805143f + x2 * (-8.1668825f + x2 * (74.5198746f + x2 * (-451.5002441f + x2 * (1551.2922363f + x2 * (-2735.5422363f + x2 * 1923.1572266f))))) + y2 * -0.1086027f);
        const float ny = uy * (1.0238225f + y2 * -0.1025671f + x2 * (-0.2639930f + x2 * 0.2979266f));
        ux = nx; uy = ny;
    }
};
struct GoproMapFn {      // gopro_warp.rs:22-39
    static GF_DEV void map(float& ux, float& uy, const float* p) {
        const float x = rs_clamp(ux, -0.5f, 0.5f), y = rs_clamp(uy, -0.5f, 0.5f);
        const float x2 = x * x, y2 = y * y;
        const float poly_x = p[0] + x2 * (p[1] + x2 * (p[2] + x2 * (p[3] + x2 * (p[4] + x2 * (p[5] + x2 * p[6])))));
        const float nx = x * (poly_x + p[7] * y2) + (ux - x);
        const float ny = y * (p[8] + p[9] * y2 + p[10] * y2 * y2 + x2 * (p[11] + p[12] * y2 + p[13] * x2)) + (uy - y);
        ux = nx; uy = ny;
    }
};

// the *_view.rs family: normalise, warp, x-scale, de-normalise
template <typename Fn, int XSCALE_KIND /*0 none, 1 superview 1.333333333, 2 hyperview 1.555555555*/>
struct ViewLens {
    static GF_DEV float xs() { return XSCALE_KIND == 1 ? 1.333333333f : 1.555555555f; }
    static GF_DEV bool undistort(float ux, float uy, const gf_kernel_params& P, bool noop, float& ox, float& oy) {
        const float cw = (float)P.output_width, chh = (float)P.output_height;
        ux = (ux / cw) - 0.5f; uy = (uy / chh) - 0.5f;
        Fn::map(ux, uy, nullptr);
        if (XSCALE_KIND != 0) ux = ux / xs();
        ox = (ux + 0.5f) * cw; oy = (uy + 0.5f) * chh;
        return true;
    }
    static GF_DEV void distort(float x, float y, float, const gf_kernel_params& P, bool noop, float& ox, float& oy) {
        const float sw = (float)P.width, sh = (float)P.height;
        x = (x / sw) - 0.5f; y = (y / sh) - 0.5f;
        if (XSCALE_KIND != 0) x = x * xs();
        float ppx = x, ppy = y;
        for (int i = 0; i < 12; ++i) {
            float dx = ppx, dy = ppy;
            Fn::map(dx, dy, nullptr);
            dx = dx - x; dy = dy - y;
            if (fabsf(dx) < 1e-6f && fabsf(dy) < 1e-6f) break;
            ppx -= dx; ppy -= dy;
        }
        ox = (ppx + 0.5f) * sw; oy = (ppy + 0.5f) * sh;
    }
};
template <> struct Lens<GF_LENS_GOPRO_SUPERVIEW>  : ViewLens<SuperviewFn, 1>  {};   // gopro_superview.rs:23-57
template <> struct Lens<GF_LENS_GOPRO6_SUPERVIEW> : ViewLens<Superview6Fn, 0> {};   // gopro6_superview.rs:21-51
template <> struct Lens<GF_LENS_GOPRO_HYPERVIEW>  : ViewLens<HyperviewFn, 2>  {};   // gopro_hyperview.rs:21-55

template <> struct Lens<GF_LENS_GOPRO_WARP> {
    static GF_DEV bool undistort(float ux, float uy, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // gopro_warp.rs:43-56
        const float* p = P.digital_lens_params;
        const float factor = p[14] != 0.0f ? p[14] : 1.0f;
        const float cw = (float)P.output_width, chh = (float)P.output_height;
        ux = (ux / cw) - 0.5f; uy = (uy / chh) - 0.5f;
        GoproMapFn::map(ux, uy, p);
        ux = ux / factor;
        ox = (ux + 0.5f) * cw; oy = (uy + 0.5f) * chh;
        return true;
    }
    static GF_DEV void distort(float x, float y, float, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // gopro_warp.rs:60-94
        const float* p = P.digital_lens_params;
        const float factor = p[14] != 0.0f ? p[14] : 1.0f;
        const float sw = (float)P.width, sh = (float)P.height;
        x = (x / sw) - 0.5f; y = (y / sh) - 0.5f;
        const float tx = x * factor, ty = y;
        float ppx = x, ppy = y;
        for (int i = 0; i < 12; ++i) {
            float dx = ppx, dy = ppy;
            GoproMapFn::map(dx, dy, p);
            dx = dx - tx; dy = dy - ty;
            if (fabsf(dx) < 1e-6f && fabsf(dy) < 1e-6f) break;
            ppx -= dx; ppy -= dy;
        }
        float rx = ppx, ry = ppy;
        GoproMapFn::map(rx, ry, p);
        if (fabsf(rx - tx) > 0.02f || fabsf(ry - ty) > 0.02f) { ox = -99999.0f; oy = -99999.0f; return; }
        ox = (ppx + 0.5f) * sw; oy = (ppy + 0.5f) * sh;
    }
};
template <> struct Lens<GF_LENS_DIGITAL_STRETCH> {
    static GF_DEV bool undistort(float ux, float uy, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // digital_stretch.rs:12-15
        ox = ux / P.digital_lens_params[0]; oy = uy / P.digital_lens_params[1]; return true;
    }
    static GF_DEV void distort(float x, float y, float, const gf_kernel_params& P, bool noop, float& ox, float& oy) {   // digital_stretch.rs:19-22
        ox = x * P.digital_lens_params[0]; oy = y * P.digital_lens_params[1];
    }
};
// "no digital lens" (Option::None): identity, never called when DIGITAL == GF_LENS_NONE
template <> struct Lens<GF_LENS_NONE> {
    static GF_DEV bool undistort(float ux, float uy, const gf_kernel_params&, bool, float& ox, float& oy) { ox = ux; oy = uy; return true; }
    static GF_DEV void distort(float x, float y, float, const gf_kernel_params&, bool, float& ox, float& oy) { ox = x; oy = y; }
};

} // namespace gf
