// warp_kernel_tile.cuh — packed two-pixel kernel with the rolling-shutter row search amortised over a warp tile.
//
// The reference evaluates the full rotate -> distort chain twice per pixel when matrix_count > 1: once with the middle
// row's matrix only to learn which scanline `sy = clamp(round(pt.y))` the pixel came from (cpu_undistort.rs:470-479), then
// again with that row's matrix (:482-483).  Only the integer `sy` of the first evaluation is used, and pt.y is a smooth
// function of the output position.  Each warp therefore owns a 32 x 8 pixel region and
//   A. evaluates the first pass EXACTLY (same packed arithmetic as warp_kernel_x2) on a coarse 8 x 8 grid of it
//      (columns 0,4,9,13,18,22,27,31 of the region — both ends included — times all 8 rows);
//   B. walks the region in four 8 x 8 tiles: every pixel interpolates pt.y between its two bracketing coarse samples and
//      takes sy from the estimate iff the estimate is provably decisive: both samples valid (w > 0 is monotone in x, so
//      everything between two valid samples is valid), the local slope change is tiny (bounds the interpolation error far
//      below the margin) and the estimate is at least GF_ROW_MARGIN away from a rounding boundary.  If ANY pixel of the
//      tile is not decisive the whole tile falls back to the exact first pass (a warp-uniform branch).
// Either way sy is the value the reference computes; the second pass and the sampling are unchanged.
// tests/test_parity_gpu.py checks the result bit-for-bit against the oracle, incl. full 4K/8K frames and a stress sweep
// over many frames/rotations.
#pragma once
#include "warp_kernel_x2.cuh"

namespace gf {

#define GF_ROW_MARGIN      0.015625f    // 1/64 px: distance the interpolated pt.y must keep from a x.5 boundary
#define GF_SLOPE_TOL       0.0009765625f // 1/1024 px/px: max slope change between adjacent coarse intervals (interp. error < 1e-3 px)
#define GF_TILE_REGION_W   32
#define GF_TILE_REGION_H   8

GF_DEV int coarse_col(int i) { return (9 * i) >> 1; }                 // 0,4,9,13,18,22,27,31
GF_DEV int coarse_index(int c) { return min((((2 * c + 1) * 57) >> 9), 6); }   // largest i <= 6 with coarse_col(i) <= c

template <int LENS, class PIX, int MINB>
__global__ void __launch_bounds__(GF_BLOCK_X * GF_BLOCK_Y, MINB)
warp_kernel_tile(const __grid_constant__ WarpArgs A) {
    using namespace p2;
    __shared__ AtanRow atan_tab[ATAN_ROWS];
    atan_table_init(atan_tab, threadIdx.y * GF_BLOCK_X + threadIdx.x, GF_BLOCK_X * GF_BLOCK_Y);
    __syncthreads();

    const gf_kernel_params& P = A.p;
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int x0 = blockIdx.x * (2 * GF_TILE_REGION_W) + (warp & 1) * GF_TILE_REGION_W;
    const int y0 = blockIdx.y * (4 * GF_TILE_REGION_H) + (warp >> 1) * GF_TILE_REGION_H;
    if (x0 >= A.out_cols || y0 >= A.out_rows) return;                 // warp-uniform
    const int l7 = lane & 7, rp = lane >> 3;
    const int ya = y0 + 2 * rp;                                        // this lane's two rows: ya, ya + 1
    const float opy_a = map_apply_int((float)ya, A.omap_y);
    const float opy_b = map_apply_int((float)(ya + 1), A.omap_y);
    const f2 py = mk(opy_a + P.translation2d[1], opy_b + P.translation2d[1]);
    const int lim = A.rs_lim;
    const bool rs = (A.feat & F_RS) != 0;
    const unsigned FULL = 0xffffffffu;

    // ---- A. exact first pass on the coarse grid ------------------------------------------------------------
    float cy_a = 0.0f, cy_b = 0.0f; int cvalid = 0;
    if (rs) {
        const MatRow mid = load_row(A.matrices, (uint32_t)P.matrix_count / 2u);
        const float opx_c = map_apply_int((float)(x0 + coarse_col(l7)), A.omap_x);
        f2 tu, tv; bool oa, ob;
        rotate_and_distort_x2<LENS>(bc(opx_c + P.translation2d[0]), py, mid, mid, A, atan_tab, tu, tv, oa, ob);
        cy_a = tv.x; cy_b = tv.y; cvalid = (oa ? 1 : 0) | (ob ? 2 : 0);
    }

    const unsigned long long ostride = (unsigned long long)P.output_stride;
    const bool row_a_ok = ya < A.out_rows && opy_a >= 0.0f && as_i32(opy_a) < P.output_height;
    const bool row_b_ok = (ya + 1) < A.out_rows && opy_b >= 0.0f && as_i32(opy_b) < P.output_height;
    const uint32_t last = (uint32_t)(P.matrix_count - 1);

    // ---- B. four 8 x 8 tiles --------------------------------------------------------------------------------
    #pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int c = 8 * it + l7;
        const int x = x0 + c;
        if (x0 + 8 * it >= A.out_cols) break;                          // warp-uniform: the rest of the region is outside the row
        const float opx = map_apply_int((float)x, A.omap_x);
        const bool in_x = x < A.out_cols && opx >= 0.0f && as_i32(opx) < P.output_width;
        const unsigned long long off_a = (unsigned long long)ya * ostride + (unsigned long long)x * PIX::BYTES;
        const unsigned long long off_b = off_a + ostride;
        const bool wr_a = in_x && row_a_ok && off_a + PIX::BYTES <= A.dst_len;
        const bool wr_b = in_x && row_b_ok && off_b + PIX::BYTES <= A.dst_len;
        const f2 px = bc(opx + P.translation2d[0]);

        int sy_a = max(min(as_i32(rs_round(py.x)), lim), 0);           // :465-469 (used when the first pass returns None)
        int sy_b = max(min(as_i32(rs_round(py.y)), lim), 0);
        if (rs) {
            // bracketing coarse samples of this lane's rows live in lanes rp*8 + ia, rp*8 + ia + 1
            const int ia = coarse_index(c), ca = coarse_col(ia), cb = coarse_col(ia + 1);
            const int ip = ia > 0 ? ia - 1 : ia + 2;                   // a third sample for the slope-change test
            const int base = rp * 8;
            const float va_a = __shfl_sync(FULL, cy_a, base + ia), vb_a = __shfl_sync(FULL, cy_a, base + ia + 1), vp_a = __shfl_sync(FULL, cy_a, base + ip);
            const float va_b = __shfl_sync(FULL, cy_b, base + ia), vb_b = __shfl_sync(FULL, cy_b, base + ia + 1), vp_b = __shfl_sync(FULL, cy_b, base + ip);
            const int vmask = __shfl_sync(FULL, cvalid, base + ia) & __shfl_sync(FULL, cvalid, base + ia + 1) & __shfl_sync(FULL, cvalid, base + ip);
            const float inv_h = 1.0f / (float)(cb - ca);
            const float t = (float)(c - ca) * inv_h;
            const float inv_hp = 1.0f / (float)(coarse_col(ip) - ca);  // signed: negative when the third sample is on the left
            const float sl_a = (vb_a - va_a) * inv_h, sl_b = (vb_b - va_b) * inv_h;
            const float est_a = va_a + (vb_a - va_a) * t, est_b = va_b + (vb_b - va_b) * t;
            const float fa = est_a + 0.5f, fb = est_b + 0.5f;
            const float ka = floorf(fa), kb = floorf(fb);
            const float da = fa - ka, db = fb - kb;
            const bool dec_a = (vmask & 1) && fabsf((vp_a - va_a) * inv_hp - sl_a) < GF_SLOPE_TOL && da >= GF_ROW_MARGIN && da <= 1.0f - GF_ROW_MARGIN && fabsf(est_a) < 1e6f;
            const bool dec_b = (vmask & 2) && fabsf((vp_b - va_b) * inv_hp - sl_b) < GF_SLOPE_TOL && db >= GF_ROW_MARGIN && db <= 1.0f - GF_ROW_MARGIN && fabsf(est_b) < 1e6f;
            const bool undecided = (wr_a && !dec_a) || (wr_b && !dec_b);
            if (__any_sync(FULL, undecided)) {                         // exact first pass for this tile (:470-479)
                const MatRow mid = load_row(A.matrices, (uint32_t)P.matrix_count / 2u);
                f2 tu, tv; bool oa, ob;
                rotate_and_distort_x2<LENS>(px, py, mid, mid, A, atan_tab, tu, tv, oa, ob);
                if (oa) sy_a = max(min(as_i32(rs_round(tv.x)), lim), 0);
                if (ob) sy_b = max(min(as_i32(rs_round(tv.y)), lim), 0);
            } else {
                sy_a = max(min((int)ka, lim), 0);
                sy_b = max(min((int)kb, lim), 0);
            }
        }
        if (!wr_a && !wr_b) continue;
        const MatRow ra = load_row(A.matrices, min((uint32_t)sy_a, last));                              // :482
        const MatRow rb = load_row(A.matrices, min((uint32_t)sy_b, last));
        f2 u, v; bool ok_a, ok_b;
        rotate_and_distort_x2<LENS>(px, py, ra, rb, A, atan_tab, u, v, ok_a, ok_b);                     // :483
        u = map_apply_x2(u, A.smap_x);                                                                  // :510-515
        v = map_apply_x2(v, A.smap_y);
        if (wr_a) shade_lean<PIX>(ok_a, u.x, v.x, A, A.dst + off_a);                                    // :615-622
        if (wr_b) shade_lean<PIX>(ok_b, u.y, v.y, A, A.dst + off_b);
    }
}

} // namespace gf
