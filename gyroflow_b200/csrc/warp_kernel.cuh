// warp_kernel.cuh — the fused undistort -> rotate (per-scanline) -> redistort -> sample kernel.
//
// One thread = one output pixel; x is the fastest thread index so that stores of a warp cover one
// contiguous 32*bpp-byte span and the 2x2 (or IxI) source taps of neighbouring lanes land in the same
// L1 lines.  Everything per-frame-uniform (the 368-byte KernelParams, derived constants) rides in the
// kernel's __grid_constant__ parameter block, i.e. the constant bank: no loads, uniform registers.
// The per-scanline matrices (rows x 14 f32) are read through the read-only L1 path: neighbouring
// pixels resolve to the same or adjacent rows, so a warp touches one or two 56-byte rows.
//
// Behavioural source (bit-exact target): src/core/stabilization/cpu_undistort.rs:133-228 (rotate_and_distort),
// :329-419 (sample_input_at), :421-517 (undistort_coord), :519-633 (main loop).  Compile with -fmad=false.
#pragma once
#include "lens_models.cuh"
#include "gf_coeffs_tables.h"
#include <cuda_fp16.h>

namespace gf {

struct WarpArgs {
    gf_kernel_params p;             // verbatim KernelParams
    const uint8_t* src;
    uint8_t*       dst;
    const float*   matrices;        // device, rows x 14, 8-byte aligned
    const float*   mesh;            // device f32 (nullptr when mesh_len == 0)
    unsigned long long src_len, dst_len;
    int   mesh_len;
    int   out_rows;                 // ceil(dst_len / output_stride): rows the reference iterates (par_chunks_mut)
    int   out_cols;                 // floor(output_stride / bpp): pixels per full row (chunks_mut)
    int   src_vec_ok, dst_vec_ok;   // base pointer and stride allow whole-pixel vector access
    // derived on the host with the same IEEE float ops as cpu_undistort.rs:521-528
    float r_limit_sq;
    float out_c[2], out_f[2];
    float bg[4];
};

// ------------------------------------------------------------------------------------------
// Pixel formats (pixel_formats.rs): COUNT channels of a SCALAR type.
// ------------------------------------------------------------------------------------------
enum { SC_U8 = 0, SC_U16 = 1, SC_F32 = 2, SC_F16 = 3 };

template <int COUNT_, int SCALAR_> struct Pix {
    static constexpr int COUNT = COUNT_;
    static constexpr int SCALAR = SCALAR_;
    static constexpr int SBYTES = SCALAR_ == SC_U8 ? 1 : (SCALAR_ == SC_F32 ? 4 : 2);
    static constexpr int BYTES = COUNT_ * SBYTES;
    static constexpr bool VEC = (BYTES == 1 || BYTES == 2 || BYTES == 4 || BYTES == 8 || BYTES == 16);

    static GF_DEV float scalar_to_float(uint32_t raw) {
        if (SCALAR == SC_F32) return __uint_as_float(raw);
        if (SCALAR == SC_F16) return __half2float(__ushort_as_half((unsigned short)raw));
        return (float)raw;                                  // u8 / u16 widen exactly
    }
    static GF_DEV uint32_t float_to_scalar(float v) {       // PixelType::from_float: Rust `as` casts
        if (SCALAR == SC_F32) return __float_as_uint(v);
        if (SCALAR == SC_F16) return (uint32_t)__half_as_ushort(__float2half_rn(v));
        int i = __float2int_rz(v);                          // trunc, saturating, NaN -> 0
        const int hi = SCALAR == SC_U8 ? 255 : 65535;
        i = i < 0 ? 0 : (i > hi ? hi : i);
        return (uint32_t)i;
    }
    // to_float
    static GF_DEV void load(const uint8_t* __restrict__ p, bool vec_ok, float (&v)[COUNT]) {
        if (VEC && vec_ok) {
            if (BYTES == 1) { v[0] = scalar_to_float(__ldg(p)); }
            else if (BYTES == 2) {
                const uint32_t w = __ldg(reinterpret_cast<const unsigned short*>(p));
                if (COUNT == 1) v[0] = scalar_to_float(w);
                else { v[0] = scalar_to_float(w & 0xffu); v[COUNT > 1 ? 1 : 0] = scalar_to_float(w >> 8); }
            } else if (BYTES == 4) {
                const uint32_t w = __ldg(reinterpret_cast<const unsigned int*>(p));
                if (COUNT == 1) v[0] = scalar_to_float(w);
                else if (COUNT == 2) { v[0] = scalar_to_float(w & 0xffffu); v[COUNT > 1 ? 1 : 0] = scalar_to_float(w >> 16); }
                else {
                    #pragma unroll
                    for (int i = 0; i < COUNT; ++i) v[i] = scalar_to_float((w >> (8 * i)) & 0xffu);
                }
            } else if (BYTES == 8) {
                const uint2 w = __ldg(reinterpret_cast<const uint2*>(p));
                #pragma unroll
                for (int i = 0; i < COUNT; ++i) { const uint32_t q = (i < 2) ? w.x : w.y; v[i] = scalar_to_float((q >> (16 * (i & 1))) & 0xffffu); }
            } else {
                const uint4 w = __ldg(reinterpret_cast<const uint4*>(p));
                const uint32_t q[4] = {w.x, w.y, w.z, w.w};
                #pragma unroll
                for (int i = 0; i < COUNT; ++i) v[i] = scalar_to_float(q[i & 3]);
            }
        } else {
            #pragma unroll
            for (int i = 0; i < COUNT; ++i) {
                uint32_t raw = 0;
                #pragma unroll
                for (int b = 0; b < SBYTES; ++b) raw |= (uint32_t)__ldg(p + i * SBYTES + b) << (8 * b);
                v[i] = scalar_to_float(raw);
            }
        }
    }
    static GF_DEV void store(uint8_t* __restrict__ p, bool vec_ok, const float (&v)[COUNT]) {
        uint32_t s[COUNT];
        #pragma unroll
        for (int i = 0; i < COUNT; ++i) s[i] = float_to_scalar(v[i]);
        if (VEC && vec_ok) {
            if (BYTES == 1) { *p = (uint8_t)s[0]; }
            else if (BYTES == 2) {
                const uint32_t w = COUNT == 1 ? s[0] : (s[0] | (s[COUNT > 1 ? 1 : 0] << 8));
                *reinterpret_cast<unsigned short*>(p) = (unsigned short)w;
            } else if (BYTES == 4) {
                uint32_t w;
                if (COUNT == 1) w = s[0];
                else if (COUNT == 2) w = s[0] | (s[COUNT > 1 ? 1 : 0] << 16);
                else { w = 0; _Pragma("unroll") for (int i = 0; i < COUNT; ++i) w |= s[i] << (8 * i); }
                *reinterpret_cast<unsigned int*>(p) = w;
            } else if (BYTES == 8) {
                uint2 w = make_uint2(0u, 0u);
                #pragma unroll
                for (int i = 0; i < COUNT; ++i) { if (i < 2) w.x |= s[i] << (16 * (i & 1)); else w.y |= s[i] << (16 * (i & 1)); }
                *reinterpret_cast<uint2*>(p) = w;
            } else {
                *reinterpret_cast<uint4*>(p) = make_uint4(s[0], s[COUNT > 1 ? 1 : 0], s[COUNT > 2 ? 2 : 0], s[COUNT > 3 ? 3 : 0]);
            }
        } else {
            #pragma unroll
            for (int i = 0; i < COUNT; ++i) {
                #pragma unroll
                for (int b = 0; b < SBYTES; ++b) p[i * SBYTES + b] = (uint8_t)(s[i] >> (8 * b));
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// Mesh correction (f64) — gyro_source/splines.rs:100-176, sony.rs:557-563.  The reference widens the
// f32 mesh to f64 once per frame (cpu_undistort.rs:539); here each element is widened on load.
// ------------------------------------------------------------------------------------------
#define GF_MAX_GRID 9
struct MeshView {
    const float* __restrict__ m;
    GF_DEV double operator[](uint32_t i) const { return (double)__ldg(m + i); }
};

static __device__ __noinline__ double mesh_bivariate(const MeshView mesh, uint32_t n_x, uint32_t n_y, double size_x, double size_y,
                                              uint32_t mesh_offset, double x, double y) {
    double a[GF_MAX_GRID], b[GF_MAX_GRID], c[GF_MAX_GRID], d[GF_MAX_GRID], alpha[GF_MAX_GRID], mu[GF_MAX_GRID], z[GF_MAX_GRID];
    uint32_t i = as_usize_small(((double)n_x - 1.0) * x / size_x);
    if (i > n_x - 2) i = n_x - 2;
    const double dx = x - size_x * (double)i / (double)(n_x - 1);
    const double dx2 = dx * dx;
    const uint32_t grid = GF_MAX_GRID, raw_mesh_len = n_x * n_y * 2, block = grid * 4;
    const uint32_t offs = 9 + raw_mesh_len + (mesh_offset * n_y * block) + i;
    for (uint32_t j = 0; j < n_y; ++j) {
        const uint32_t rb = offs + j * block;
        a[j] = mesh[rb] + mesh[rb + grid] * dx + mesh[rb + grid * 2] * dx2 + mesh[rb + grid * 3] * dx2 * dx;   // intermediate_values[j]; cubic_spline_coefficients copies it into a[]
    }
    // cubic_spline_coefficients(intermediate, step 1, offset 0, size_y, n_y) — splines.rs:100-124
    const uint32_t n = n_y;
    const double h = size_y / (double)(n - 1);
    const double inv_h = 1.0 / h;
    const double three_inv_h = 3.0 * inv_h;
    const double h_over_3 = h / 3.0;
    const double inv_3h = 1.0 / (3.0 * h);
    for (uint32_t q = 1; q + 1 < n; ++q) alpha[q] = three_inv_h * (a[q + 1] - 2.0 * a[q] + a[q - 1]);
    mu[0] = 0.0; z[0] = 0.0;
    for (uint32_t q = 1; q + 1 < n; ++q) {
        mu[q] = 1.0 / (4.0 - mu[q - 1]);
        z[q] = (alpha[q] * inv_h - z[q - 1]) * mu[q];
    }
    c[n - 1] = 0.0;
    for (int q = (int)n - 2; q >= 0; --q) {
        c[q] = z[q] - mu[q] * c[q + 1];
        b[q] = (a[q + 1] - a[q]) * inv_h - h_over_3 * (c[q + 1] + 2.0 * c[q]);
        d[q] = (c[q + 1] - c[q]) * inv_3h;
    }
    // cubic_spline_interpolate — splines.rs:126-139
    if (y <= 0.0) return a[0] + b[0] * y;
    if (y >= size_y) {
        const double slope = b[n - 2] + 2.0 * c[n - 2] * h + 3.0 * d[n - 2] * h * h;
        return a[n - 1] + slope * (y - size_y);
    }
    uint32_t k = as_usize_small(((double)n - 1.0) * y / size_y);
    if (k > n - 2) k = n - 2;
    const double dy = y - size_y * (double)k / (double)(n - 1);
    return a[k] + b[k] * dy + c[k] * dy * dy + d[k] * dy * dy * dy;
}

// ------------------------------------------------------------------------------------------
// rotate_and_distort — cpu_undistort.rs:133-228
// ------------------------------------------------------------------------------------------
template <int LENS, int DIGITAL>
GF_DEV bool rotate_and_distort(float px, float py, uint32_t idx, const WarpArgs& A, float& ou, float& ov) {
    const gf_kernel_params& P = A.p;
    const float2* __restrict__ mp = reinterpret_cast<const float2*>(A.matrices + (size_t)idx * GF_MATRIX_STRIDE);
    const float2 m01 = __ldg(mp + 0), m23 = __ldg(mp + 1), m45 = __ldg(mp + 2), m67 = __ldg(mp + 3), m8_9 = __ldg(mp + 4);
    const float _x = (px * m01.x) + (py * m01.y) + m23.x + P.translation3d[0];
    const float _y = (px * m23.y) + (py * m45.x) + m45.y + P.translation3d[1];
    float       _w = (px * m67.x) + (py * m67.y) + m8_9.x + P.translation3d[2];
    if (!(_w > 0.0f)) return false;
    if (A.r_limit_sq > 0.0f && (_x * _x + _y * _y) > A.r_limit_sq * _w) return false;                // :139 (sic: * _w)

    if (P.light_refraction_coefficient != 1.0f && P.light_refraction_coefficient > 0.0f) {            // :143-152
        if (_w != 0.0f) {
            const float r = sqrtf(_x * _x + _y * _y) / _w;
            const float sin_theta_d = (r / sqrtf(1.0f + r * r)) * P.light_refraction_coefficient;
            const float r_d = sin_theta_d / sqrtf(1.0f - sin_theta_d * sin_theta_d);
            if (r_d != 0.0f) _w *= r / r_d;
        }
    }

    float ux, uy;
    Lens<LENS>::distort(_x, _y, _w, P, ux, uy);                                                        // :154
    ux = ux * P.f[0]; uy = uy * P.f[1];                                                                // :155

    const float2 m10_11 = __ldg(mp + 5), m12_13 = __ldg(mp + 6);
    if (m8_9.y != 0.0f || m10_11.x != 0.0f || m10_11.y != 0.0f || m12_13.x != 0.0f || m12_13.y != 0.0f) {   // :157-165
        const float ang_rad = m10_11.y;
        const float cos_a = gf_cosf(-ang_rad), sin_a = gf_sinf(-ang_rad);
        const float tx = cos_a * ux - sin_a * uy - m8_9.y   + m12_13.x;
        const float ty = sin_a * ux + cos_a * uy - m10_11.x + m12_13.y;
        ux = tx; uy = ty;
    }

    ux = ux + P.c[0]; uy = uy + P.c[1];                                                                // :167

    if (A.mesh_len > 0) {
        const MeshView mesh{A.mesh};
        const double mesh0 = mesh[0];
        if (mesh0 > 10.0) {                                                                            // :169-185
            const float origin_x = (float)mesh[5], origin_y = (float)mesh[6];
            const float crop_w = (float)mesh[7], crop_h = (float)mesh[8];
            if ((P.flags & 128) == 128) uy = (float)P.height - uy;
            ux = map_coord(ux, 0.0f, (float)P.width,  origin_x, origin_x + crop_w);
            uy = map_coord(uy, 0.0f, (float)P.height, origin_y, origin_y + crop_h);
            const uint32_t n_x = as_usize_small(mesh[1]), n_y = as_usize_small(mesh[2]);
            const double sx = mesh[3], sy = mesh[4];
            const double nx = mesh_bivariate(mesh, n_x, n_y, sx, sy, 0, (double)ux, (double)uy);
            const double ny = mesh_bivariate(mesh, n_x, n_y, sx, sy, 1, (double)ux, (double)uy);
            ux = map_coord((float)nx, origin_x, origin_x + crop_w, 0.0f, (float)P.width);
            uy = map_coord((float)ny, origin_y, origin_y + crop_h, 0.0f, (float)P.height);
            if ((P.flags & 128) == 128) uy = (float)P.height - uy;
        }
        // FocalPlaneDistortion :188-214 (a missing FPD block means "none"; the reference would index out of bounds)
        const uint32_t o = as_usize_small(mesh0);
        if (mesh0 > 0.0 && o < (uint32_t)A.mesh_len && mesh[o] > 0.0) {
            const double mesh_size_y = mesh[4];
            const float origin_x = (float)mesh[5], origin_y = (float)mesh[6];
            const float crop_w = (float)mesh[7], crop_h = (float)mesh[8];
            const double stblz_grid = mesh_size_y / 8.0;
            if ((P.flags & 128) == 128) uy = (float)P.height - uy;
            ux = map_coord(ux, 0.0f, (float)P.width,  origin_x, origin_x + crop_w);
            uy = map_coord(uy, 0.0f, (float)P.height, origin_y, origin_y + crop_h);
            const uint32_t idx2 = as_usize_small(fmin(fmax(floor((double)uy / stblz_grid), 0.0), 7.0));
            const double delta = (double)uy - stblz_grid * (double)idx2;
            ux -= (float)(mesh[o + 4 + idx2 * 2 + 0] * delta);
            uy -= (float)(mesh[o + 4 + idx2 * 2 + 1] * delta);
            for (uint32_t j = 0; j < idx2; ++j) {
                ux -= (float)(mesh[o + 4 + j * 2 + 0] * stblz_grid);
                uy -= (float)(mesh[o + 4 + j * 2 + 1] * stblz_grid);
            }
            ux = map_coord(ux, origin_x, origin_x + crop_w, 0.0f, (float)P.width);
            uy = map_coord(uy, origin_y, origin_y + crop_h, 0.0f, (float)P.height);
            if ((P.flags & 128) == 128) uy = (float)P.height - uy;
        }
    }

    if (DIGITAL != GF_LENS_NONE && (P.flags & 2) == 2) {                                               // :216-220
        float dx, dy;
        Lens<DIGITAL>::distort(ux, uy, 1.0f, P, dx, dy);
        ux = dx; uy = dy;
    }

    if (P.input_horizontal_stretch > 0.001f) ux /= P.input_horizontal_stretch;                         // :222-223
    if (P.input_vertical_stretch   > 0.001f) uy /= P.input_vertical_stretch;

    ou = ux; ov = uy;
    return true;
}

GF_DEV void rotate_point(float px, float py, float angle, float ox, float oy, float o2x, float o2y, float& rx, float& ry) {   // :262-265
    const float ca = gf_cosf(angle), sa = gf_sinf(angle);
    rx = ca * (px - ox) - sa * (py - oy) + o2x;
    ry = sa * (px - ox) + ca * (py - oy) + o2y;
}

// undistort_coord — cpu_undistort.rs:421-517
template <int LENS, int DIGITAL>
GF_DEV bool undistort_coord(float ox_, float oy_, const WarpArgs& A, float& ru, float& rv) {
    const gf_kernel_params& P = A.p;
    float opx = map_coord(ox_, (float)P.output_rect[0], (float)(P.output_rect[0] + P.output_rect[2]), 0.0f, (float)P.output_width);
    float opy = map_coord(oy_, (float)P.output_rect[1], (float)(P.output_rect[1] + P.output_rect[3]), 0.0f, (float)P.output_height);
    opx += P.translation2d[0];
    opy += P.translation2d[1];

    if (P.lens_correction_amount < 1.0f) {                                                             // :429-460
        float nx = opx, ny = opy;
        const float ocx = A.out_c[0], ocy = A.out_c[1], ofx = A.out_f[0], ofy = A.out_f[1];
        if (DIGITAL != GF_LENS_NONE && (P.flags & 2) == 2) {
            const float uzx = (nx - ocx) * P.fov + ocx, uzy = (ny - ocy) * P.fov + ocy;
            float tx, ty;
            if (Lens<DIGITAL>::undistort(uzx, uzy, P, tx, ty)) {
                nx = (tx - ocx) / P.fov + ocx;
                ny = (ty - ocy) / P.fov + ocy;
            }
        }
        nx = (nx - ocx) / ofx; ny = (ny - ocy) / ofy;
        { float tx, ty; if (Lens<LENS>::undistort(nx, ny, P, tx, ty)) { nx = tx; ny = ty; } }
        if (P.light_refraction_coefficient != 1.0f && P.light_refraction_coefficient > 0.0f) {
            const float r = sqrtf(nx * nx + ny * ny);
            if (r != 0.0f) {
                const float sin_theta_d = (r / sqrtf(1.0f + r * r)) / P.light_refraction_coefficient;
                const float r_d = sin_theta_d / sqrtf(1.0f - sin_theta_d * sin_theta_d);
                const float factor = r_d / r;
                nx *= factor; ny *= factor;
            }
        }
        nx = (nx * ofx) + ocx; ny = (ny * ofy) + ocy;
        const float ia = 1.0f - P.lens_correction_amount;
        opx = nx * ia + (opx * P.lens_correction_amount);
        opy = ny * ia + (opy * P.lens_correction_amount);
    }

    // rolling-shutter row :465-482
    const bool hrs = (P.flags & 16) == 16;
    const int lim = hrs ? P.width : P.height;
    int sy = as_i32(rs_round(hrs ? opx : opy));
    sy = max(min(sy, lim), 0);
    if (P.matrix_count > 1) {
        float tu, tv;
        if (rotate_and_distort<LENS, DIGITAL>(opx, opy, (uint32_t)P.matrix_count / 2u, A, tu, tv)) {
            int v = as_i32(rs_round(hrs ? tu : tv));
            sy = max(min(v, lim), 0);
        }
    }
    const uint32_t idx = min((uint32_t)sy, (uint32_t)(P.matrix_count - 1));

    float u, v;
    if (!rotate_and_distort<LENS, DIGITAL>(opx, opy, idx, A, u, v)) return false;                      // :483

    float fsx = (float)P.width, fsy = (float)P.height;
    if (P.input_rotation != 0.0f) {                                                                    // :485-491
        const float rotation = P.input_rotation * (3.14159274101257324f / 180.0f);
        const float sx = fsx, sy2 = fsy;
        rotate_point(sx, sy2, rotation, 0.0f, 0.0f, 0.0f, 0.0f, fsx, fsy);
        fsx = rs_round(fabsf(fsx)); fsy = rs_round(fabsf(fsy));
        float nu, nv;
        rotate_point(u, v, rotation, sx / 2.0f, sy2 / 2.0f, fsx / 2.0f, fsy / 2.0f, nu, nv);
        u = nu; v = nv;
    }

    const float width_f = (float)P.width, height_f = (float)P.height;
    if (P.background_mode == 1) {                                                                      // edge repeat :495-499
        u = rs_min(rs_max(u, 3.0f), width_f - 3.0f);
        v = rs_min(rs_max(v, 3.0f), height_f - 3.0f);
    } else if (P.background_mode == 2) {                                                               // edge mirror :500-509
        const float rx = rs_round(u), ry = rs_round(v);
        const float width3 = width_f - 3.0f, height3 = height_f - 3.0f;
        if (rx > width3)  u = width3  - (rx - width3);
        if (rx < 3.0f)    u = 3.0f + width_f - (width3 + rx);
        if (ry > height3) v = height3 - (ry - height3);
        if (ry < 3.0f)    v = 3.0f + height_f - (height3 + ry);
    }
    if (P.background_mode != 3) {                                                                      // :510-515
        u = map_coord(u, 0.0f, fsx, (float)P.source_rect[0], (float)(P.source_rect[0] + P.source_rect[2]));
        v = map_coord(v, 0.0f, fsy, (float)P.source_rect[1], (float)(P.source_rect[1] + P.source_rect[3]));
    }
    ru = u; rv = v;
    return true;
}

// ------------------------------------------------------------------------------------------
// sample_input_at, separable branch (I = 2, 4, 8) — cpu_undistort.rs:370-418
// ------------------------------------------------------------------------------------------
template <int I> GF_DEV void coeff_row(uint32_t frac, float (&c)[I]) {
    if (I == 2) { c[0] = 1.0f - (float)frac / 32.0f; c[1] = (float)frac / 32.0f; }     // == COEFFS[frac*2 ..], exact dyadics
    else if (I == 4) { _Pragma("unroll") for (int i = 0; i < I; ++i) c[i] = GF_COEFFS_BICUBIC_DEV[(frac << 2) + i]; }
    else { _Pragma("unroll") for (int i = 0; i < I; ++i) c[i] = GF_COEFFS_LANCZOS4_DEV[(frac << 3) + i]; }
}

template <int I, class PIX>
GF_DEV void sample_input_at(float uvx, float uvy, const WarpArgs& A, float (&sum)[PIX::COUNT]) {
    const gf_kernel_params& P = A.p;
    constexpr int C = PIX::COUNT;
    const float offset = I == 2 ? 0.0f : (I == 4 ? 1.0f : 3.0f);
    const float u = uvx - offset, v = uvy - offset;
    const int sx0 = as_i32(rs_round(u * 32.0f));
    const int sy0 = as_i32(rs_round(v * 32.0f));
    const int sx = sx0 >> 5, sy = sy0 >> 5;
    float cx[I], cy[I];
    coeff_row<I>((uint32_t)sx0 & 31u, cx);
    coeff_row<I>((uint32_t)sy0 & 31u, cy);
    const int rx0 = P.source_rect[0], ry0 = P.source_rect[1], rx1 = rx0 + P.source_rect[2], ry1 = ry0 + P.source_rect[3];
    #pragma unroll
    for (int ch = 0; ch < C; ++ch) sum[ch] = 0.0f;
    const long long row_base = (long long)sy * (long long)P.stride + (long long)sx * (long long)PIX::BYTES;
    #pragma unroll
    for (int yp = 0; yp < I; ++yp) {
        if (sy + yp >= ry0 && sy + yp < ry1) {
            float xsum[C];
            #pragma unroll
            for (int ch = 0; ch < C; ++ch) xsum[ch] = 0.0f;
            #pragma unroll
            for (int xp = 0; xp < I; ++xp) {
                float px[C];
                if (sx + xp >= rx0 && sx + xp < rx1) {
                    PIX::load(A.src + (row_base + (long long)yp * P.stride + (long long)xp * PIX::BYTES), A.src_vec_ok != 0, px);
                } else {
                    #pragma unroll
                    for (int ch = 0; ch < C; ++ch) px[ch] = A.bg[ch];
                }
                #pragma unroll
                for (int ch = 0; ch < C; ++ch) xsum[ch] += px[ch] * cx[xp];
            }
            #pragma unroll
            for (int ch = 0; ch < C; ++ch) sum[ch] += xsum[ch] * cy[yp];
        } else {
            #pragma unroll
            for (int ch = 0; ch < C; ++ch) sum[ch] += A.bg[ch] * cy[yp];
        }
    }
    #pragma unroll
    for (int ch = 0; ch < C; ++ch) sum[ch] = rs_min(sum[ch], P.pixel_value_limit);
}

template <int C> GF_DEV void remap_colorrange(float (&px)[C], bool is_y) {      // cpu_undistort.rs:255-260
    const float s = is_y ? 0.85882352f : 0.87843137f;
    #pragma unroll
    for (int ch = 0; ch < C; ++ch) px[ch] *= s;
    px[0] += 16.0f;
    if (C > 1) px[C > 1 ? 1 : 0] += 16.0f;
}

// ------------------------------------------------------------------------------------------
// The kernel — main loop of undistort_image_cpu, cpu_undistort.rs:543-625
// ------------------------------------------------------------------------------------------
#define GF_BLOCK_X 32
#define GF_BLOCK_Y 8

template <int LENS, int DIGITAL, class PIX, int I>
__global__ void __launch_bounds__(GF_BLOCK_X * GF_BLOCK_Y)
warp_kernel(const __grid_constant__ WarpArgs A) {
    const gf_kernel_params& P = A.p;
    constexpr int C = PIX::COUNT;
    const int x = blockIdx.x * GF_BLOCK_X + threadIdx.x;
    const int y = blockIdx.y * GF_BLOCK_Y + threadIdx.y;
    if (x >= A.out_cols || y >= A.out_rows) return;
    const unsigned long long off = (unsigned long long)y * (unsigned long long)P.output_stride + (unsigned long long)x * PIX::BYTES;
    if (off + PIX::BYTES > A.dst_len) return;                     // trailing partial row (chunks_mut of a short last row)

    const float opx = map_coord((float)x, (float)P.output_rect[0], (float)(P.output_rect[0] + P.output_rect[2]), 0.0f, (float)P.output_width);
    const float opy = map_coord((float)y, (float)P.output_rect[1], (float)(P.output_rect[1] + P.output_rect[3]), 0.0f, (float)P.output_height);
    if (!(opx >= 0.0f && opy >= 0.0f && as_i32(opx) < P.output_width && as_i32(opy) < P.output_height)) return;   // :551

    uint8_t* const out = A.dst + off;
    float pixel[C];
    #pragma unroll
    for (int ch = 0; ch < C; ++ch) pixel[ch] = A.bg[ch];
    if ((P.flags & 4) == 4) { PIX::store(out, A.dst_vec_ok != 0, pixel); return; }                              // fill_bg :558-561

    const bool fix_range = (P.flags & 1) == 1, is_y = P.plane_index == 0;
    float u, v;
    if (undistort_coord<LENS, DIGITAL>((float)x, (float)y, A, u, v)) {                                          // :565
        if (P.background_mode == 3) {                                                                            // :576-613
            const float width_f = (float)P.width, height_f = (float)P.height;
            const float widthf = width_f - 1.0f, heightf = height_f - 1.0f;
            const float feather = rs_max(P.background_margin_feather * heightf, 0.0001f);
            float p2x = u, p2y = v, alpha = 1.0f;
            if ((u > widthf - feather) || (u < feather) || (v > heightf - feather) || (v < feather)) {
                alpha = rs_max(rs_min(rs_min(rs_min(rs_min(widthf - u, heightf - v), u), v) / feather, 1.0f), 0.0f);
                p2x = p2x / width_f; p2y = p2y / height_f;
                p2x = ((p2x - 0.5f) * (1.0f - P.background_margin)) + 0.5f;
                p2y = ((p2y - 0.5f) * (1.0f - P.background_margin)) + 0.5f;
                p2x = p2x * width_f; p2y = p2y * height_f;
            }
            float fsx = width_f, fsy = height_f;
            if (P.input_rotation != 0.0f) {
                const float rotation = P.input_rotation * (3.14159274101257324f / 180.0f);
                rotate_point(width_f, height_f, rotation, 0.0f, 0.0f, 0.0f, 0.0f, fsx, fsy);
                fsx = rs_round(fabsf(fsx)); fsy = rs_round(fabsf(fsy));
            }
            const float sx0 = (float)P.source_rect[0], sx1 = (float)(P.source_rect[0] + P.source_rect[2]);
            const float sy0 = (float)P.source_rect[1], sy1 = (float)(P.source_rect[1] + P.source_rect[3]);
            u   = map_coord(u,   0.0f, fsx, sx0, sx1); v   = map_coord(v,   0.0f, fsy, sy0, sy1);
            p2x = map_coord(p2x, 0.0f, fsx, sx0, sx1); p2y = map_coord(p2y, 0.0f, fsy, sy0, sy1);
            float c1[C], c2[C];
            sample_input_at<I, PIX>(u, v, A, c1);
            sample_input_at<I, PIX>(p2x, p2y, A, c2);
            #pragma unroll
            for (int ch = 0; ch < C; ++ch) pixel[ch] = c1[ch] * alpha + c2[ch] * (1.0f - alpha);
            if (fix_range) remap_colorrange<C>(pixel, is_y);
            PIX::store(out, A.dst_vec_ok != 0, pixel);
            return;
        }
        sample_input_at<I, PIX>(u, v, A, pixel);                                                                 // :615
    }
    if (fix_range) remap_colorrange<C>(pixel, is_y);                                                             // :619-621
    PIX::store(out, A.dst_vec_ok != 0, pixel);                                                                   // :622
}

} // namespace gf
