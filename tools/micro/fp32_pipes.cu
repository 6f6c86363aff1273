// Microbenchmark (development aid): FP32 pipe rates on B200 -- scalar FFMA, packed FFMA2 (fma.rn.f32x2, new on
// sm_100), and each mixed with independent ALU / LDS work, to see what shares an issue slot with what.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp32_pipes tools/micro/fp32_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int *iout) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i * 1e-6f;
    __syncthreads();
    float2 a0 = make_float2(threadIdx.x * 1e-3f, 1.f), a1 = make_float2(2.f, 3.f), a2 = make_float2(4.f, 5.f),
           a3 = make_float2(6.f, 7.f);
    const float2 m = make_float2(0.999f + blockIdx.x * 1e-9f, 0.998f);
    // MODE >= 6: the addend is a run-time register too (3-register FFMA, no immediate form)
    const float2 b = MODE >= 6 ? make_float2(1e-3f + blockIdx.x * 1e-9f, 2e-3f + threadIdx.x * 1e-9f) : make_float2(1e-3f, 2e-3f);
    int i0 = threadIdx.x, i1 = blockIdx.x, i2 = 3, i3 = 7;
    float l0 = 0.f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2 || MODE == 4 || MODE == 6) { // scalar: 8 FFMA
                a0.x = fmaf(a0.x, m.x, b.x); a0.y = fmaf(a0.y, m.y, b.y); a1.x = fmaf(a1.x, m.x, b.x); a1.y = fmaf(a1.y, m.y, b.y);
                a2.x = fmaf(a2.x, m.x, b.x); a2.y = fmaf(a2.y, m.y, b.y); a3.x = fmaf(a3.x, m.x, b.x); a3.y = fmaf(a3.y, m.y, b.y);
            } else { // packed: 4 FFMA2 = the same 8 FMAs
                a0 = __ffma2_rn(a0, m, b); a1 = __ffma2_rn(a1, m, b); a2 = __ffma2_rn(a2, m, b); a3 = __ffma2_rn(a3, m, b);
            }
            if (MODE == 2 || MODE == 3) { // + 4 independent integer ALU ops (LOP3/IADD3)
                i0 = (i0 ^ i1) + 3; i1 = (i1 & i2) + i0; i2 = (i2 | i3) ^ i0; i3 = i3 + i1;
            }
            if (MODE == 4 || MODE == 5) { // + 2 shared-memory loads
                l0 += sm[(i0 + u * 33 + it) & 1023]; l0 += sm[(i0 + u * 65 + it + 512) & 1023];
            }
        }
    }
    float r = a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + l0;
    if (r == 123.456f) out[0] = r;
    if ((i0 ^ i1 ^ i2 ^ i3) == 0x12345678) iout[0] = i0;
}

template <int MODE>
void run(const char *name, int nSM, float *d, int *di) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<nSM * 8, 256>>>(d, di);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    double fma = 2.0 * 8 * 8 * ITERS * (double)nSM * 8 * 256;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s fp32\n", name, best, fma / (best * 1e-3) * 1e-12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float *d; int *di; cudaMalloc(&d, 64); cudaMalloc(&di, 64);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("FFMA", p.multiProcessorCount, d, di);
    run<1>("FFMA2", p.multiProcessorCount, d, di);
    run<2>("FFMA  + 4 int ALU / 8 FMA", p.multiProcessorCount, d, di);
    run<3>("FFMA2 + 4 int ALU / 8 FMA", p.multiProcessorCount, d, di);
    run<4>("FFMA  + 2 LDS / 8 FMA", p.multiProcessorCount, d, di);
    run<5>("FFMA2 + 2 LDS / 8 FMA", p.multiProcessorCount, d, di);
    run<6>("FFMA 3-register", p.multiProcessorCount, d, di);
    run<7>("FFMA2 3-register", p.multiProcessorCount, d, di);
    return 0;
}
