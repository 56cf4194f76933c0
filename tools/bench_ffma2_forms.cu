// Micro-benchmark: does FFMA2's issue rate on sm_100a depend on where its operands come from (register pairs vs uniform registers)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o bench_ffma2_forms tools/bench_ffma2_forms.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
// mode 0: all three operands in registers (constants loaded from global memory, opaque to the compiler)
// mode 1: the constant operand comes from a kernel parameter (uniform register / constant bank)
template <int MODE>
__global__ void k(float* out, const float2* cptr, float2 nz, float2 one, float2 a2, float2 b2) {
    float2 v[CH];
    float2 cn = nz, co = one, ca = a2, cb = b2;
    if (MODE == 0) { cn = cptr[threadIdx.x & 1]; co = cptr[2 + (threadIdx.x & 1)]; ca = cptr[4 + (threadIdx.x & 1)]; cb = cptr[6 + (threadIdx.x & 1)]; }
    for (int i = 0; i < CH; ++i) v[i] = make_float2(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1);
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < CH; ++i) { v[i] = __ffma2_rn(v[i], ca, cn); v[i] = __ffma2_rn(v[i], co, cb); }
    }
    float s = 0; for (int i = 0; i < CH; ++i) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// scalar reference: FMUL + FADD with a register / uniform constant
template <int MODE>
__global__ void ks(float* out, const float* cptr, float a, float b) {
    float v[2 * CH]; float ca = a, cb = b;
    if (MODE == 0) { ca = cptr[threadIdx.x & 1]; cb = cptr[2 + (threadIdx.x & 1)]; }
    for (int i = 0; i < 2 * CH; ++i) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 2 * CH; ++i) { v[i] = v[i] * ca; v[i] = v[i] + cb; }
    }
    float s = 0; for (int i = 0; i < 2 * CH; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// issue-slot model: does an FFMA2 (or an ALU instruction) block the issue port for its second cycle?  An FP stream interleaved 1:1
// with an independent integer LOP3 stream: if the pair costs max(fp, int) the pipes overlap, if it costs the sum they do not.
template <int MODE>   // 0: FFMA2 + LOP3, 1: FMUL/FADD + LOP3, 2: LOP3 only
__global__ void kmix(float* out, float2 nz, float2 one, float2 a2, float2 b2, unsigned m) {
    float2 v[CH]; unsigned w[CH];
    for (int i = 0; i < CH; ++i) { v[i] = make_float2(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1); w[i] = threadIdx.x * 977u + i; }
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) { v[i] = __ffma2_rn(v[i], a2, nz); w[i] = (w[i] ^ m) | (w[i] >> 31); v[i] = __ffma2_rn(v[i], one, b2); w[i] = (w[i] & ~m) ^ (w[i] << 1); }
            if (MODE == 1) { v[i].x = v[i].x * a2.x; w[i] = (w[i] ^ m) | (w[i] >> 31); v[i].x = v[i].x + b2.x; w[i] = (w[i] & ~m) ^ (w[i] << 1); }
            if (MODE == 2) { w[i] = (w[i] ^ m) | (w[i] >> 31); w[i] = (w[i] & ~m) ^ (w[i] << 1); }
        }
    }
    float s = 0; for (int i = 0; i < CH; ++i) s += v[i].x + v[i].y + (float)w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    float2 h[8] = {{-0.f, -0.f}, {-0.f, -0.f}, {1, 1}, {1, 1}, {1.0000001f, 1.0000001f}, {1.0000001f, 1.0000001f}, {1e-7f, 1e-7f}, {1e-7f, 1e-7f}};
    float2* c2; cudaMalloc(&c2, sizeof(h)); cudaMemcpy(c2, h, sizeof(h), cudaMemcpyHostToDevice);
    float hs[4] = {1.0000001f, 1.0000001f, 1e-7f, 1e-7f}; float* c1; cudaMalloc(&c1, sizeof(hs)); cudaMemcpy(c1, hs, sizeof(hs), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) for (int mode = 0; mode < 4; ++mode) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(d, c2, h[0], h[2], h[4], h[6]);
        if (mode == 1) k<1><<<148 * 8, 256>>>(d, c2, h[0], h[2], h[4], h[6]);
        if (mode == 2) ks<0><<<148 * 8, 256>>>(d, c1, hs[0], hs[2]);
        if (mode == 3) ks<1><<<148 * 8, 256>>>(d, c1, hs[0], hs[2]);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warp_instr = 148.0 * 8 * 8 * ITERS * 2.0 * CH * (mode < 2 ? 1 : 2);      // per launch
        const double cyc = ms * 1e-3 * 1.965e9 * 148 * 4;                                       // SMSP-cycles at 1.965 GHz
        const char* nm[4] = {"FFMA2 R,R,R       ", "FFMA2 const operand", "FMUL/FADD R,R     ", "FMUL/FADD const   "};
        printf("%s: %.3f ms  %.2f cycles per warp-instruction per SMSP\n", nm[mode], ms, cyc / warp_instr);
    }
    for (int mode = 0; mode < 3; ++mode) {
        cudaEventRecord(e0);
        if (mode == 0) kmix<0><<<148 * 8, 256>>>(d, h[0], h[2], h[4], h[6], 0x5a5a5a5au);
        if (mode == 1) kmix<1><<<148 * 8, 256>>>(d, h[0], h[2], h[4], h[6], 0x5a5a5a5au);
        if (mode == 2) kmix<2><<<148 * 8, 256>>>(d, h[0], h[2], h[4], h[6], 0x5a5a5a5au);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double groups = 148.0 * 8 * 8 * ITERS * 2.0 * CH;      // (fp op + int op) groups per launch
        const double cyc = ms * 1e-3 * 1.965e9 * 148 * 4;
        const char* nm[3] = {"FFMA2 + int ops interleaved", "FMUL/FADD + int ops        ", "int ops alone              "};
        printf("%s: %.3f ms  %.2f cycles per (fp, int) group per SMSP\n", nm[mode], ms, cyc / groups);
    }
    return 0;
}
