// selftest.cu — device-side check of the packed f32x2 primitives against the scalar operations they replace.
// gf_cuda_selftest(device, n, out[4]): n pseudo-random operand pairs per primitive; out = mismatch counts for
// {div_exact_checked vs '/', sqrt_exact vs sqrtf, atanf2 vs gf_atanf, div_uniform vs '/'}.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/gyroflow_cuda.h"
#include "warp_kernel_x2.cuh"

using namespace gf;

namespace {
inline uint32_t __float_as_uint_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
__device__ __forceinline__ uint32_t mix(uint64_t v) {
    v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
    return (uint32_t)v;
}
// operand generator: a mix of fully random bit patterns, "coordinate-like" magnitudes and values near 1
__device__ __forceinline__ float gen(uint64_t i, uint32_t salt) {
    const uint32_t h = mix(i * 0x9E3779B97F4A7C15ULL + salt);
    const uint32_t sel = mix(i + 77u * salt) & 7u;
    if (sel < 3) return __uint_as_float(h);                                             // anything, incl. NaN/inf/denormals
    if (sel < 6) return __uint_as_float((h & 0x807fffffu) | ((100u + (mix(i ^ salt) % 56u)) << 23));   // 2^-27 .. 2^28
    return __uint_as_float((h & 0x007fffffu) | 0x3f800000u) - ((h >> 31) ? 0.0f : 1.0f);        // [1,2) or [0,1)
}
__device__ __forceinline__ bool same(float a, float b) { return __float_as_uint(a) == __float_as_uint(b) || (a != a && b != b); }

__global__ void selftest_kernel(unsigned long long n, unsigned long long seed, unsigned long long* out, uint32_t* dbg) {
    __shared__ p2::AtanRow tab[p2::ATAN_ROWS];
    p2::atan_table_init(tab, threadIdx.x, blockDim.x);
    __syncthreads();
    unsigned long long bad[4] = {0, 0, 0, 0};
    if (blockIdx.x == 0 && threadIdx.x < p2::ATAN_ROWS) {       // the literal table in global memory == the computed rows
        const p2::AtanRow& g = p2::GF_ATAN_TAB[threadIdx.x]; const p2::AtanRow& c = tab[threadIdx.x];
        if (!(same(g.A, c.A) && same(g.B, c.B) && same(g.C, c.C) && same(g.D, c.D) && same(g.hi, c.hi) && same(g.lo, c.lo))) bad[2]++;
    }
    MapC m; m.in_min = 0.0f; m.mul = 1.0f; m.add = 0.0f; m.identity = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = i + seed;
        const float a0 = gen(k, 1), a1 = gen(k, 2), b0 = gen(k, 3), b1 = gen(k, 4);
        const p2::f2 q = p2::div_exact_checked(p2::mk(a0, a1), p2::mk(b0, b1));
        if (!same(q.x, a0 / b0)) bad[0]++;
        if (!same(q.y, a1 / b1)) bad[0]++;
        const float s0 = fabsf(a0), s1 = fabsf(b1);
        const p2::f2 r = p2::sqrt_exact(p2::mk(s0, s1));
        if (!same(r.x, sqrtf(s0))) bad[1]++;
        if (!same(r.y, sqrtf(s1))) bad[1]++;
        const p2::f2 t = p2::atanf2(p2::mk(a1, b0), tab);
        if (!same(t.x, gf_atanf(a1))) { bad[2]++; if (dbg) { const unsigned slot = atomicAdd(&dbg[0], 1u); if (slot < 16) { dbg[1 + slot * 4] = __float_as_uint(a1); dbg[2 + slot * 4] = __float_as_uint(t.x); dbg[3 + slot * 4] = __float_as_uint(gf_atanf(a1)); dbg[4 + slot * 4] = __float_as_uint(b0); } } }
        if (!same(t.y, gf_atanf(b0))) bad[2]++;
        // uniform divisor: typical frame sizes and random ones
        const float dv = (k & 1) ? (float)(16 + (mix(k) % 16368)) : fabsf(gen(k, 5));
        m.div = dv; m.rcp = 1.0f / dv;
        const float ad = fabsf(dv);
        m.fast_div = (ad >= 0x1p-40f && ad <= 0x1p40f) ? 1 : 0;
        if (!same(div_uniform(a0, m), a0 / dv)) bad[3]++;
        if (m.fast_div) {
            bool flagged = false;                        // the kernel sends flagged pairs to the scalar code; unflagged ones must be exact
            const p2::f2 mq = map_apply_x2(p2::mk(a0, b1), m, flagged);
            { const float n0 = fabsf(a0), n1 = fabsf(b1);      // the host-side guarantees of map_apply_x2, restated for random operands
              if (!((n0 == 0.0f || n0 > 0x1p-74f) && (n1 == 0.0f || n1 > 0x1p-74f) && dv <= 0x1p20f)) flagged = true; }
            if (!flagged && !same(mq.x, ((a0 - 0.0f) * 1.0f) / dv + 0.0f)) bad[3]++;
            if (!flagged && !same(mq.y, ((b1 - 0.0f) * 1.0f) / dv + 0.0f)) bad[3]++;
        }
        if (round_away_i32(a0) != as_i32(rs_round(a0))) bad[3]++;
        if (round_away_i32(b1 * 32.0f) != as_i32(rs_round(b1 * 32.0f))) bad[3]++;
        {   // conversion-free rounding: exact inside its window, a safe sentinel outside (contract at round_half_away_w)
            int wa, wb;
            round_half_away_w(p2::mul(p2::mk(a0, b0), p2::bc(64.0f)), wa, wb);
            const float t[2] = {a0 * 32.0f, b0 * 32.0f}; const int r[2] = {wa >> 1, wb >> 1};
            for (int j = 0; j < 2; ++j) {
                const int e = as_i32(rs_round(t[j]));
                const bool fine = (t[j] > -0.25f && t[j] < 4194304.0f) ? (r[j] == e)
                                : (t[j] >= 4194304.0f) ? (r[j] >= 4194304) : (r[j] <= 0 && (e >= 0 || r[j] < 0));   // t <= -1/4 or NaN
                if (!fine) bad[3]++;
            }
            const int lim = (int)(mix(k ^ 0x55u) & 0x3fffffu);
            int ra, rb;
            round_away_clamped_x2(p2::mk(a1, b1), lim, ra, rb);
            if (ra != max(min(as_i32(rs_round(a1)), lim), 0)) bad[3]++;
            if (rb != max(min(as_i32(rs_round(b1)), lim), 0)) bad[3]++;
        }
        // straight-line atanf on the range the kernel admits (r = sqrt(a), a in [2^-56, 2^48))
        const float c0 = fabsf(a1), c1 = fabsf(b0);
        if (in_window_r2(c0) && in_window_r2(c1)) {
            const p2::f2 rr = p2::sqrt_seq(p2::mk(c0, c1));
            if (!same(rr.x, sqrtf(c0)) || !same(rr.y, sqrtf(c1))) bad[1]++;
            const p2::f2 tc = p2::atanf2_core(rr, p2::GF_ATAN_TAB);
            if (!same(tc.x, gf_atanf(rr.x))) bad[2]++;
            if (!same(tc.y, gf_atanf(rr.y))) bad[2]++;
        }
        // unguarded division inside the windows the kernel checks
        if (p2::in_window(b0) && p2::in_window(b1) && zero_or_in_window(a0) && zero_or_in_window(a1)) {
            const p2::f2 qs = p2::div_seq(p2::mk(a0, a1), p2::mk(b0, b1));
            // a == -0 comes out as +0 (the residual fma loses the sign); no consumer in the kernel sees the sign of a zero
            const float e0 = a0 / b0, e1 = a1 / b1;
            if (!(same(qs.x, e0) || (qs.x == 0.0f && e0 == 0.0f)) || !(same(qs.y, e1) || (qs.y == 0.0f && e1 == 0.0f))) bad[0]++;
        }
    }
    for (int j = 0; j < 4; ++j) if (bad[j]) atomicAdd(&out[j], bad[j]);
}
// Exhaustive sweeps (opt-in, GF_RUN_EXHAUSTIVE=1 in the tests): every float of the packed atanf's admitted range [2^-28, 2^24) and of
// the packed square root's window [2^-56, 2^48), two consecutive floats per thread (lane .x / lane .y), against the scalar functions.
__global__ void sweep_kernel(uint32_t lo, uint32_t hi, int which, unsigned long long* out) {
    unsigned long long bad = 0;
    const uint32_t step = 2u * gridDim.x * blockDim.x;
    for (uint64_t u = (uint64_t)lo + 2u * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x); u + 1 < hi; u += step) {
        const float a = __uint_as_float((uint32_t)u), b = __uint_as_float((uint32_t)u + 1u);
        if (which == 0) {
            const p2::f2 t = p2::atanf2_core(p2::mk(a, b), p2::GF_ATAN_TAB);
            if (!same(t.x, gf_atanf(a))) bad++;
            if (!same(t.y, gf_atanf(b))) bad++;
        } else {
            const p2::f2 r = p2::sqrt_seq(p2::mk(a, b));
            if (!same(r.x, sqrtf(a))) bad++;
            if (!same(r.y, sqrtf(b))) bad++;
        }
    }
    if (bad) atomicAdd(out, bad);
}
} // namespace

// out2[0]: mismatches of atanf2_core over all floats in [2^-28, 2^24); out2[1]: of sqrt_seq over all floats in [2^-56, 2^48).
extern "C" GF_API int gf_cuda_selftest_exhaustive(int device, unsigned long long* out2) {
    if (!out2) return GF_ERR_BAD_PARAMS;
    if (cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return GF_ERR_CUDA; }
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, 2 * sizeof(unsigned long long)) != cudaSuccess) return GF_ERR_CUDA;
    cudaMemset(d, 0, 2 * sizeof(unsigned long long));
    sweep_kernel<<<148 * 16, 256>>>(__float_as_uint_host(0x1p-28f), __float_as_uint_host(0x1p24f), 0, d);
    sweep_kernel<<<148 * 16, 256>>>(__float_as_uint_host(0x1p-56f), __float_as_uint_host(0x1p48f), 1, d + 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out2, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return GF_ERR_CUDA; }
    return GF_OK;
}

extern "C" GF_API int gf_cuda_selftest(int device, unsigned long long n, unsigned long long seed, unsigned long long* out4) {
    if (!out4) return GF_ERR_BAD_PARAMS;
    if (cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return GF_ERR_CUDA; }
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, 4 * sizeof(unsigned long long)) != cudaSuccess) return GF_ERR_CUDA;
    cudaMemset(d, 0, 4 * sizeof(unsigned long long));
    uint32_t* dbg = nullptr;
    if (getenv("GF_SELFTEST_DEBUG")) { cudaMalloc(&dbg, 65 * 4); cudaMemset(dbg, 0, 65 * 4); }
    selftest_kernel<<<148 * 8, 256>>>(n, seed, d, dbg);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out4, d, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (dbg) {
        uint32_t h[65]; cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        for (unsigned i = 0; i < (h[0] < 16 ? h[0] : 16); ++i) {
            float x, g, w, o; memcpy(&x, &h[1 + i * 4], 4); memcpy(&g, &h[2 + i * 4], 4); memcpy(&w, &h[3 + i * 4], 4); memcpy(&o, &h[4 + i * 4], 4);
            printf("atanf2 mismatch: x=%a (%08x) got=%a want=%a other-lane=%a (%08x)\n", x, h[1 + i * 4], g, w, o, h[4 + i * 4]);
        }
        cudaFree(dbg);
    }
    cudaFree(d);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return GF_ERR_CUDA; }
    return GF_OK;
}
