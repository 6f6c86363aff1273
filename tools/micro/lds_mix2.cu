// Microbenchmark 2 (development aid, round 2): cost of shared-memory operand delivery next to 128 FFMA per iteration,
// k_eval residency (4 warps per CTA, 5 CTAs per SM).  All loads/stores are volatile inline PTX (no hoisting, no merging).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/lds_mix2 tools/micro/lds_mix2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ float4 lds128(unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds64(unsigned a) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float lds32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts128(unsigned a, float4 v) { asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
// MODE: 0 none
//  1 K x LDS.128, lanes of a group of 8 share an address (4 addresses per warp)      2 K x LDS.128, one address per warp
//  3 K x LDS.128, one address per lane (stride 144 B, conflict-free)                  4 2K x LDS.64 (4 addresses per warp)
//  5 4K x LDS.32 (4 addresses per warp)                                               6 K x STS.128 per lane (values from the chains)
//  7 K x (LDS.128 per lane at the top, STS.128 per lane at the bottom, different addresses: accumulators of ANOTHER sink)
//  8 K x SHFL.IDX of 4 values (operand broadcast by shuffle)
template <int MODE, int K, int NF>
__global__ void __launch_bounds__(128, 5) k(float *out, int n) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 2560; i += 128) sm[i] = make_float4(1.f + i * 1e-7f, 1e-3f, 2e-3f, 3e-3f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float m0 = 0.999f + blockIdx.x * 1e-9f, m1 = 1e-4f + lane * 1e-9f;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + warp * 640 * 16;
    if (MODE == 1 || MODE == 4 || MODE == 5) base += (lane >> 3) * 144;
    if (MODE == 3 || MODE == 6 || MODE == 7) base += lane * 144;
    for (int it = 0; it < n; ++it) {
        // the address moves every iteration (8 positions): nothing is loop-invariant
        const unsigned off = base + (it & 7) * ((MODE == 3 || MODE == 6 || MODE == 7) ? 16 : 4 * 144);
        float4 v[K > 0 ? K : 1];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 7) v[j] = lds128(off + 16 * j);
            if (MODE == 4) { const float2 p = lds64(off + 16 * j), q = lds64(off + 16 * j + 8); v[j] = make_float4(p.x, p.y, q.x, q.y); }
            if (MODE == 5) v[j] = make_float4(lds32(off + 16 * j), lds32(off + 16 * j + 4), lds32(off + 16 * j + 8), lds32(off + 16 * j + 12));
            if (MODE == 8) v[j] = make_float4(__shfl_sync(0xffffffffu, a[0], (it + j) & 31), __shfl_sync(0xffffffffu, a[1], (it + j) & 31),
                                              __shfl_sync(0xffffffffu, a[2], (it + j) & 31), __shfl_sync(0xffffffffu, a[3], (it + j) & 31));
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            float mul = m0;
            if (K > 0 && MODE != 6) { const float4 w = v[(f / 4) % (K > 0 ? K : 1)]; mul = (f & 3) == 0 ? w.x : (f & 3) == 1 ? w.y : (f & 3) == 2 ? w.z : w.w; }
            a[f & 7] = fmaf(a[f & 7], mul, m1);
        }
        if (MODE == 6 || MODE == 7) {
#pragma unroll
            for (int j = 0; j < K; ++j) sts128(off + 16 * j + (MODE == 7 ? 8 * 16 : 0), make_float4(a[0], a[1], a[2], a[3 + (j & 3)]));
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    if (r == 123.456f) out[0] = r;
}

template <int MODE, int K, int NF>
void run(const char *name, int nSM, float *d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k<MODE, K, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2560 * 16);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE, K, NF><<<nSM * 5, 128, 2560 * 16>>>(d, ITERS);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cyc = best * 1e-3 * clk * 1e3 / (5.0 * ITERS);
    printf("%-66s %7.3f ms %6.1f cycles/warp-iteration, %5.1f beyond the %d FFMA\n", name, best, cyc, cyc - NF, NF);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float *d; cudaMalloc(&d, 64);
    const int n = p.multiProcessorCount;
    printf("%s, %d SMs, %d kHz; 20 warps per SM\n", p.name, n, p.clockRate);
    run<0, 0, 128>("128 FFMA", n, d);
    run<1, 9, 128>("+ 9 LDS.128, 4 addresses per warp (k_eval today)", n, d);
    run<2, 9, 128>("+ 9 LDS.128, 1 address per warp", n, d);
    run<3, 9, 128>("+ 9 LDS.128, 32 addresses (per lane, conflict-free)", n, d);
    run<4, 9, 128>("+ 18 LDS.64, 4 addresses per warp", n, d);
    run<5, 9, 128>("+ 36 LDS.32, 4 addresses per warp", n, d);
    run<1, 1, 128>("+ 1 LDS.128, 4 addresses per warp", n, d);
    run<2, 1, 128>("+ 1 LDS.128, 1 address per warp", n, d);
    run<3, 1, 128>("+ 1 LDS.128 per lane", n, d);
    run<3, 2, 128>("+ 2 LDS.128 per lane", n, d);
    run<6, 1, 128>("+ 1 STS.128 per lane", n, d);
    run<6, 2, 128>("+ 2 STS.128 per lane", n, d);
    run<7, 1, 128>("+ 1 LDS.128 + 1 STS.128 per lane (smem accumulators)", n, d);
    run<7, 2, 128>("+ 2 LDS.128 + 2 STS.128 per lane", n, d);
    run<8, 1, 128>("+ 4 SHFL.IDX", n, d);
    run<8, 9, 128>("+ 36 SHFL.IDX", n, d);
    return 0;
}
