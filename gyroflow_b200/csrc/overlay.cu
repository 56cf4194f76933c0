// overlay.cu — the preview overlays of the reference's GPU kernels: draw_pixel + draw_safe_area
// (src/core/gpu/opencl_undistort.cl:109-154; colour / alpha tables :109-120; buffer produced by gpu/drawing.rs).
// The CPU path — the parity target of the warp itself — draws none of this (cpu_undistort.rs:234-251, :607, :617 are commented
// out), so the stage is OFF unless the caller switches it on (gf_cuda_set_overlays); it exists for the live-preview caller
// (SURVEY §8 f4).  Both uses of draw_pixel are separate passes around the warp kernel:
//   * input stage  (isInput, drawing entries with stage bit 0): the .cl kernel draws onto every source tap it reads (:338, :374), a
//     function of the tap's position only — identical to drawing onto the input image first.  Done on the device copy of the input
//     (the HOST staging buffer, or a context-owned copy of a DEVICE input: the caller's buffer is never modified);
//   * output stage (entries with stage bit 1) + draw_safe_area: applied to the final pixel (:644-645, :654-655) — a pass over the
//     output buffer after the warp.
// DATA_CONVERT is convert_<T>_sat (round toward zero, saturating) = the `as` cast of PixelType::from_float.  Arithmetic unfused
// (-fmad=false), like the oracle's restatement (gf_oracle_draw_overlays).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../include/gyroflow_cuda.h"
#include "c_abi_internal.h"

namespace {

__device__ __forceinline__ float ovl_load(const uint8_t* p, int scalar) {          // PixelType::to_float, one channel
    switch (scalar) {
    case 0:  return (float)p[0];
    case 1:  return (float)(*reinterpret_cast<const unsigned short*>(p));
    case 2:  return *reinterpret_cast<const float*>(p);
    default: return __half2float(__ushort_as_half(*reinterpret_cast<const unsigned short*>(p)));
    }
}
__device__ __forceinline__ void ovl_store(uint8_t* p, int scalar, float v) {        // PixelType::from_float: truncate, saturate, NaN -> 0
    switch (scalar) {
    case 0:  { int i = __float2int_rz(v); i = i < 0 ? 0 : (i > 255 ? 255 : i); p[0] = (uint8_t)i; } break;
    case 1:  { int i = __float2int_rz(v); i = i < 0 ? 0 : (i > 65535 ? 65535 : i); *reinterpret_cast<unsigned short*>(p) = (unsigned short)i; } break;
    case 2:  *reinterpret_cast<float*>(p) = v; break;
    default: *reinterpret_cast<unsigned short*>(p) = __half_as_ushort(__float2half_rn(v)); break;
    }
}

__constant__ float OVL_COLORS[9][4] = { {0, 0, 0, 0}, {255, 0, 0, 255}, {0, 255, 0, 255}, {0, 0, 255, 255}, {254, 251, 71, 255},
                                        {200, 200, 0, 255}, {255, 0, 255, 255}, {0, 128, 255, 255}, {0, 200, 200, 255} };
__constant__ float OVL_ALPHAS[4] = { 1.0f, 0.75f, 0.50f, 0.25f };

struct OverlayArgs {
    uint8_t* buf; unsigned long long len;
    int width, height, stride, count, scalar, sbytes;
    int draw, draw_width, is_input;          // flags & 8, max(params.width, params.output_width)
    float canvas_scale;
    float safe[4];
    const uint8_t* drawing; unsigned long long drawing_len;
};

__global__ void overlay_kernel(const OverlayArgs A) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= A.width || y >= A.height) return;
    const int bpp = A.count * A.sbytes;
    const unsigned long long off = (unsigned long long)y * (unsigned long long)A.stride + (unsigned long long)x * (unsigned long long)bpp;
    if (off + (unsigned long long)bpp > A.len) return;
    uint8_t* px = A.buf + off;
    if (A.draw && A.drawing_len) {                                                                   // draw_pixel :121-140
        const float fpos = roundf(floorf((float)y / A.canvas_scale) * ((float)A.draw_width / A.canvas_scale) + floorf((float)x / A.canvas_scale));
        // `(int)`: truncating, saturating, NaN -> 0 (as the oracle's restatement casts it); negative positions index nothing
        const int ipos = fpos != fpos ? 0 : __float2int_rz(fpos);
        const unsigned long long pos = ipos >= 0 ? (unsigned long long)ipos : ~0ull;
        if (pos < A.drawing_len) {
            const unsigned data = A.drawing[pos];
            const unsigned color = (data & 0xF8u) >> 3, alpha = (data & 0x06u) >> 1, stage = data & 1u;
            if (data > 0u && ((stage == 0u) == (A.is_input != 0)) && color < 9u) {
                const float af = OVL_ALPHAS[alpha], ia = 1.0f - af;
                for (int c = 0; c < A.count; ++c) {
                    const float v = ovl_load(px + c * A.sbytes, A.scalar);
                    ovl_store(px + c * A.sbytes, A.scalar, OVL_COLORS[color][c] * af + v * ia);
                }
            }
        }
    }
    if (!A.is_input) {                                                                               // draw_safe_area :141-154
        const float fx = (float)x, fy = (float)y;
        const bool safe = fx >= A.safe[0] && fx <= A.safe[2] && fy >= A.safe[1] && fy <= A.safe[3];
        if (!safe) {
            const bool border = fx >= A.safe[0] - 5.0f && fx <= A.safe[2] + 5.0f && fy >= A.safe[1] - 5.0f && fy <= A.safe[3] + 5.0f;
            for (int c = 0; c < A.count; ++c) {
                const float f = c < 3 ? 0.5f : 1.0f;
                float v = ovl_load(px + c * A.sbytes, A.scalar) * f;
                ovl_store(px + c * A.sbytes, A.scalar, v);                      // converted back to the pixel type between the two multiplications, like the .cl
                if (border) { v = ovl_load(px + c * A.sbytes, A.scalar) * f; ovl_store(px + c * A.sbytes, A.scalar, v); }
            }
        }
    }
}

} // namespace

int gf_internal_draw_overlays(void* cu_stream, uint8_t* buf_dev, size_t len, int width, int height, int stride, const gf_kernel_params* p,
                              int count, int scalar, int is_input, const uint8_t* drawing_dev, size_t drawing_len) {
    OverlayArgs A;
    A.buf = buf_dev; A.len = len; A.width = width; A.height = height; A.stride = stride; A.count = count; A.scalar = scalar;
    A.sbytes = scalar == 0 ? 1 : (scalar == 2 ? 4 : 2);
    A.draw = (p->flags & GF_FLAG_DRAWING_ENABLED) && drawing_dev && drawing_len ? 1 : 0;
    A.draw_width = p->width > p->output_width ? p->width : p->output_width;
    A.is_input = is_input; A.canvas_scale = p->canvas_scale;
    for (int i = 0; i < 4; ++i) A.safe[i] = p->safe_area_rect[i];
    A.drawing = drawing_dev; A.drawing_len = drawing_len;
    if (width <= 0 || height <= 0) return GF_OK;
    const dim3 block(32, 8), grid((unsigned)(width + 31) / 32, (unsigned)(height + 7) / 8);
    overlay_kernel<<<grid, block, 0, (cudaStream_t)cu_stream>>>(A);
    return cudaGetLastError() == cudaSuccess ? GF_OK : GF_ERR_CUDA;
}
