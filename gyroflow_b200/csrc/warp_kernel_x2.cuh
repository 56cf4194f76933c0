// warp_kernel_x2.cuh — the warp with two output pixels per thread on Blackwell's packed f32x2 pipe.
//
// Same arithmetic, same rounding, same results as warp_kernel.cuh (the scalar kernel remains the general
// implementation and the fallback for every feature this file does not cover).  What changes is the schedule:
// a thread owns the vertically adjacent pixels (x, y) and (x, y + 1); every FP32 multiply/add of the
// undistort -> rotate -> redistort chain is issued once for both (FFMA2/FMUL2/FADD2, see f32x2.cuh), the range
// selection of atanf becomes a shared-memory table lookup so both lanes run straight-line code, and division /
// square root use the compiler's own MUFU + FFMA refinement sequences on pairs.  Only the "lean" feature set
// (see F_GENERAL_ONLY in warp_kernel.cuh) is compiled here.
//
// Behavioural source: src/core/stabilization/cpu_undistort.rs:133-228, :421-517, :543-625 (as warp_kernel.cuh).
#pragma once
#include "warp_kernel.cuh"
#include "f32x2.cuh"

namespace gf {

using p2::f2;

// ------------------------------------------------------------------------------------------
// packed lens models: Lens2<M>::distort(x, y, z) for two pixels.  kHas = a packed implementation exists.
// ------------------------------------------------------------------------------------------
template <int M> struct Lens2 { static constexpr bool kHas = false; };

// opencv_fisheye.rs:72-93 (k != 0: the lean kernel is only chosen when F_LENS_NOOP is clear)
template <> struct Lens2<GF_LENS_OPENCV_FISHEYE> {
    static constexpr bool kHas = true;
    static GF_DEV void distort(f2 x, f2 y, f2 z, const gf_kernel_params& P, const p2::AtanRow* tab, f2& ox, f2& oy) {
        using namespace p2;
        const bool ok = in_window(x) && in_window(y) && in_window(z);
        x = div_exact(x, z, ok); y = div_exact(y, z, ok);
        const f2 r = sqrt_exact(add(mul(x, x), mul(y, y)));
        const f2 theta = atanf2(r, tab);
        const f2 theta2 = mul(theta, theta), theta4 = mul(theta2, theta2), theta6 = mul(theta4, theta2), theta8 = mul(theta4, theta4);
        f2 s = add(bc(1.0f), mul(bc(P.k[0]), theta2));
        s = add(s, mul(bc(P.k[1]), theta4));
        s = add(s, mul(bc(P.k[2]), theta6));
        s = add(s, mul(bc(P.k[3]), theta8));
        const f2 theta_d = mul(theta, s);
        const f2 q = div_exact(theta_d, r, in_window(theta_d) && in_window(r));
        const f2 scale = mk(r.x == 0.0f ? 1.0f : q.x, r.y == 0.0f ? 1.0f : q.y);
        ox = mul(x, scale); oy = mul(y, scale);
    }
};

// ------------------------------------------------------------------------------------------
// rotate_and_distort for two pixels — cpu_undistort.rs:133-228, lean feature set
// (no translation3d, r_limit, refraction, mesh, digital lens, input stretch).
// ------------------------------------------------------------------------------------------
struct MatRow { float2 m01, m23, m45, m67, m89, m1011, m1213; };
GF_DEV MatRow load_row(const float* __restrict__ matrices, uint32_t idx) {
    const float2* __restrict__ mp = reinterpret_cast<const float2*>(matrices + (size_t)idx * GF_MATRIX_STRIDE);
    MatRow r;
    r.m01 = __ldg(mp + 0); r.m23 = __ldg(mp + 1); r.m45 = __ldg(mp + 2); r.m67 = __ldg(mp + 3);
    r.m89 = __ldg(mp + 4); r.m1011 = __ldg(mp + 5); r.m1213 = __ldg(mp + 6);
    return r;
}
GF_DEV bool row_has_ibis(const MatRow& r) {      // :157 — any of m[9..13] != 0.0
    return ((__float_as_uint(r.m89.y) | __float_as_uint(r.m1011.x) | __float_as_uint(r.m1011.y) |
             __float_as_uint(r.m1213.x) | __float_as_uint(r.m1213.y)) << 1) != 0u;
}
GF_DEV void apply_ibis(const MatRow& r, float& ux, float& uy) {     // :158-164
    const float ang_rad = r.m1011.y;
    const float cos_a = gf_cosf(-ang_rad), sin_a = gf_sinf(-ang_rad);
    const float tx = cos_a * ux - sin_a * uy - r.m89.y   + r.m1213.x;
    const float ty = sin_a * ux + cos_a * uy - r.m1011.x + r.m1213.y;
    ux = tx; uy = ty;
}

template <int LENS>
GF_DEV void rotate_and_distort_x2(f2 px, f2 py, const MatRow& ra, const MatRow& rb, const WarpArgs& A, const p2::AtanRow* tab,
                                  f2& ou, f2& ov, bool& oka, bool& okb) {
    using namespace p2;
    const gf_kernel_params& P = A.p;
    const f2 _x = add(add(mul(px, mk(ra.m01.x, rb.m01.x)), mul(py, mk(ra.m01.y, rb.m01.y))), mk(ra.m23.x, rb.m23.x));
    const f2 _y = add(add(mul(px, mk(ra.m23.y, rb.m23.y)), mul(py, mk(ra.m45.x, rb.m45.x))), mk(ra.m45.y, rb.m45.y));
    const f2 _w = add(add(mul(px, mk(ra.m67.x, rb.m67.x)), mul(py, mk(ra.m67.y, rb.m67.y))), mk(ra.m89.x, rb.m89.x));
    oka = _w.x > 0.0f; okb = _w.y > 0.0f;                                                              // :138
    f2 ux, uy;
    Lens2<LENS>::distort(_x, _y, _w, P, tab, ux, uy);                                                  // :154
    ux = mul(ux, bc(P.f[0])); uy = mul(uy, bc(P.f[1]));                                                // :155
    if (row_has_ibis(ra)) apply_ibis(ra, ux.x, uy.x);                                                  // :157-165
    if (row_has_ibis(rb)) apply_ibis(rb, ux.y, uy.y);
    ou = add(ux, bc(P.c[0])); ov = add(uy, bc(P.c[1]));                                                // :167
}

// map_coord with a uniform divisor on a pair (see div_uniform in warp_kernel.cuh)
GF_DEV f2 map_apply_x2(f2 x, const MapC& m) {
    using namespace p2;
    const f2 a = mul(sub(x, bc(m.in_min)), bc(m.mul));
    const float a0 = fabsf(a.x), a1 = fabsf(a.y);
    f2 q;
    if (m.fast_div && a0 < 0x1p60f && a0 > 0x1p-80f && a1 < 0x1p60f && a1 > 0x1p-80f) {
        const f2 q0 = mul(a, bc(m.rcp));
        const f2 r0 = fma(bc(-m.div), q0, a);
        q = fma(r0, bc(m.rcp), q0);
    } else {
        q = mk(a.x / m.div, a.y / m.div);
    }
    return add(q, bc(m.add));
}

// sampling + conversion + store of one pixel, lean feature set (no fix_range, background mode 0)
template <class PIX>
GF_DEV void shade_lean(bool ok, float u, float v, const WarpArgs& A, uint8_t* __restrict__ out) {
    constexpr int C = PIX::COUNT;
    float pixel[C];
    if (ok) {
        if (PIX::SCALAR == SC_U8) {
            const int sx0 = as_i32(rs_round(u * 32.0f)), sy0 = as_i32(rs_round(v * 32.0f));
            const int sx = sx0 >> 5, sy = sy0 >> 5;
            if (sx >= A.src_rect[0] && sx + 2 <= A.src_rect[2] && sy >= A.src_rect[1] && sy + 2 <= A.src_rect[3]) {
                uint32_t N[C], s[C];
                sample_u8_bilinear<PIX>(sx0, sy0, A, N);
                #pragma unroll
                for (int ch = 0; ch < C; ++ch) s[ch] = (uint32_t)min((int)(N[ch] >> 10), A.u8_limit);
                PIX::store_scalars(out, true, s);
                return;
            }
            sample_generic<2, PIX>(sx0, sy0, A, pixel);
        } else {
            sample_input_at<2, PIX, false>(u, v, A, pixel);
        }
    } else {
        #pragma unroll
        for (int ch = 0; ch < C; ++ch) pixel[ch] = A.bg[ch];
    }
    PIX::store(out, true, pixel);
}

#define GF_X2_ROWS_PER_BLOCK (2 * GF_BLOCK_Y)

template <int LENS, class PIX, int MINB>
__global__ void __launch_bounds__(GF_BLOCK_X * GF_BLOCK_Y, MINB)
warp_kernel_x2(const __grid_constant__ WarpArgs A) {
    using namespace p2;
    __shared__ AtanRow atan_tab[ATAN_ROWS];
    atan_table_init(atan_tab, threadIdx.y * GF_BLOCK_X + threadIdx.x, GF_BLOCK_X * GF_BLOCK_Y);
    __syncthreads();

    const gf_kernel_params& P = A.p;
    const int x = blockIdx.x * GF_BLOCK_X + threadIdx.x;
    const int y0 = (blockIdx.y * GF_BLOCK_Y + threadIdx.y) * 2;
    if (x >= A.out_cols || y0 >= A.out_rows) return;
    const unsigned long long ostride = (unsigned long long)P.output_stride;
    const unsigned long long off_a = (unsigned long long)y0 * ostride + (unsigned long long)x * PIX::BYTES;
    const unsigned long long off_b = off_a + ostride;
    // lane validity: row exists, pixel fits in the buffer (short last row), bounds test of :551
    const float opx = map_apply_int((float)x, A.omap_x);
    const float opy_a = map_apply_int((float)y0, A.omap_y);
    const float opy_b = map_apply_int((float)(y0 + 1), A.omap_y);
    const bool in_x = opx >= 0.0f && as_i32(opx) < P.output_width;
    const bool wr_a = in_x && off_a + PIX::BYTES <= A.dst_len && opy_a >= 0.0f && as_i32(opy_a) < P.output_height;
    const bool wr_b = in_x && (y0 + 1) < A.out_rows && off_b + PIX::BYTES <= A.dst_len && opy_b >= 0.0f && as_i32(opy_b) < P.output_height;
    if (!wr_a && !wr_b) return;

    // undistort_coord, :421-517
    const f2 px = bc(opx + P.translation2d[0]);
    const f2 py = mk(opy_a + P.translation2d[1], opy_b + P.translation2d[1]);
    const int lim = A.rs_lim;
    int sy_a = max(min(as_i32(rs_round(py.x)), lim), 0);
    int sy_b = max(min(as_i32(rs_round(py.y)), lim), 0);
    if (A.feat & F_RS) {                                                                                // :470-479
        const MatRow mid = load_row(A.matrices, (uint32_t)P.matrix_count / 2u);
        f2 tu, tv; bool oa, ob;
        rotate_and_distort_x2<LENS>(px, py, mid, mid, A, atan_tab, tu, tv, oa, ob);
        if (oa) sy_a = max(min(as_i32(rs_round(tv.x)), lim), 0);
        if (ob) sy_b = max(min(as_i32(rs_round(tv.y)), lim), 0);
    }
    const uint32_t last = (uint32_t)(P.matrix_count - 1);
    const MatRow ra = load_row(A.matrices, min((uint32_t)sy_a, last));                                  // :482
    const MatRow rb = load_row(A.matrices, min((uint32_t)sy_b, last));
    f2 u, v; bool ok_a, ok_b;
    rotate_and_distort_x2<LENS>(px, py, ra, rb, A, atan_tab, u, v, ok_a, ok_b);                         // :483
    u = map_apply_x2(u, A.smap_x);                                                                      // :510-515
    v = map_apply_x2(v, A.smap_y);

    if (wr_a) shade_lean<PIX>(ok_a, u.x, v.x, A, A.dst + off_a);                                        // :615-622
    if (wr_b) shade_lean<PIX>(ok_b, u.y, v.y, A, A.dst + off_b);
}

} // namespace gf
