// Micro-benchmark: issue/throughput of scalar FP32 (FMUL/FADD) vs packed f32x2 (FMUL2/FADD2, sm_100 only).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o bench_f32x2 tools/bench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
__global__ void k_scalar(float* out, float a, float b) {
    float v[2 * CH];
    for (int i = 0; i < 2 * CH; ++i) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 2 * CH; ++i) { v[i] = v[i] * a; v[i] = v[i] + b; }
    }
    float s = 0; for (int i = 0; i < 2 * CH; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b) {
    float2 v[CH]; const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < CH; ++i) v[i] = make_float2(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1);
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < CH; ++i) { v[i] = __fmul2_rn(v[i], a2); v[i] = __fadd2_rn(v[i], b2); }
    }
    float s = 0; for (int i = 0; i < CH; ++i) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        for (int mode = 0; mode < 2; ++mode) {
            cudaEventRecord(e0);
            if (mode == 0) k_scalar<<<148 * 8, 256>>>(d, 1.0000001f, 1e-7f); else k_packed<<<148 * 8, 256>>>(d, 1.0000001f, 1e-7f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 148.0 * 8 * 256 * ITERS * 2.0 * CH * 2;   // mul + add per element
            printf("%s: %.3f ms  %.2f TFLOP/s (non-fused)\n", mode ? "packed f32x2" : "scalar      ", ms, flops / ms / 1e9);
        }
    }
    return 0;
}
