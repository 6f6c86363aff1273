// Microbenchmark (development aid, round 2): what does an LDS.128 cost next to FP32 work on B200?
// Mixes shaped like k_eval's hot loop -- NF FFMA (3-register, independent chains) per iteration plus K shared-memory
// accesses of a given kind -- run with k_eval's residency (4 warps per CTA, 5 CTAs per SM).  Cycles per iteration per
// warp-scheduler slot are derived from the elapsed time; an iteration of pure FFMA costs NF issue cycles per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/lds_mix tools/micro/lds_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
// KIND: 0 none | 1 LDS.128, 4 addresses per warp (lanes of a sub-group share one: k_eval today) | 2 LDS.128 one address
// per lane (conflict-free stride) | 3 LDS.128 one address for the whole warp | 4 = 2 + STS.128 back (accumulators in
// shared memory) | 5 LDS.32 x 4 K, 4 addresses per warp | 6 LDS.64 x 2 K
template <int KIND, int K, int NF>
__global__ void __launch_bounds__(128, 5) k(float *out, int n) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 2304; i += 128) sm[i] = make_float4(i * 1e-6f, 1e-3f, 2e-3f, 3e-3f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    float m0 = 0.999f + blockIdx.x * 1e-9f, m1 = 1e-4f + lane * 1e-9f;
    int base = warp * 576;
    if (KIND == 1 || KIND == 5 || KIND == 6) base += (lane >> 3) * 9;
    if (KIND == 2 || KIND == 4) base += lane * 9;
    for (int it = 0; it < n; ++it) {
        const int off = base + ((it & 7) * 9 * (KIND == 2 || KIND == 4 ? 0 : 4));
        float4 v[K > 0 ? K : 1];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (KIND == 1 || KIND == 2 || KIND == 3 || KIND == 4) v[j] = sm[off + j];
            if (KIND == 5) {
                const float *f = reinterpret_cast<const float *>(&sm[off + j]);
                v[j] = make_float4(f[0], f[1], f[2], f[3]);
                asm volatile("" :: "f"(v[j].x), "f"(v[j].y), "f"(v[j].z), "f"(v[j].w));
            }
            if (KIND == 6) {
                const float2 *f = reinterpret_cast<const float2 *>(&sm[off + j]);
                const float2 p = f[0], q = f[1];
                v[j] = make_float4(p.x, p.y, q.x, q.y);
            }
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            float mul = m0;
            if (K > 0) { const float4 w = v[(f / 4) % K]; mul = (f & 3) == 0 ? w.x : (f & 3) == 1 ? w.y : (f & 3) == 2 ? w.z : w.w; }
            a[f & 7] = fmaf(a[f & 7], mul, m1);
        }
        if (KIND == 4) {
#pragma unroll
            for (int j = 0; j < K; ++j) sm[off + j] = make_float4(a[0], a[1], a[2], a[3]);
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    if (r == 123.456f) out[0] = r;
}

template <int KIND, int K, int NF>
void run(const char *name, int nSM, float *d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k<KIND, K, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2304 * 16);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<KIND, K, NF><<<nSM * 5, 128, 2304 * 16>>>(d, ITERS);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // 5 warps per scheduler, each ITERS iterations: cycles per iteration per scheduler slot = time * clock / (5 * ITERS)
    const double cyc = best * 1e-3 * clk * 1e3 / (5.0 * ITERS);
    printf("%-58s %7.3f ms  %6.1f cycles per warp-iteration (%d FFMA + %d smem ops) -> %5.1f beyond the FFMAs, %4.1f per op\n", name, best, cyc,
           NF, K * (KIND == 4 ? 2 : KIND == 5 ? 4 : KIND == 6 ? 2 : 1), cyc - NF, K ? (cyc - NF) / K : 0.0);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float *d; cudaMalloc(&d, 64);
    const int n = p.multiProcessorCount;
    printf("%s, %d SMs, %d kHz\n", p.name, n, p.clockRate);
    run<0, 0, 128>("128 FFMA", n, d);
    run<1, 4, 128>("128 FFMA + 4 LDS.128 (4 addresses / warp)", n, d);
    run<1, 9, 128>("128 FFMA + 9 LDS.128 (4 addresses / warp) = k_eval", n, d);
    run<3, 9, 128>("128 FFMA + 9 LDS.128 (1 address / warp)", n, d);
    run<2, 9, 128>("128 FFMA + 9 LDS.128 (32 addresses, conflict-free)", n, d);
    run<2, 2, 128>("128 FFMA + 2 LDS.128 (32 addresses)", n, d);
    run<4, 1, 128>("128 FFMA + 1 LDS.128 + 1 STS.128 (32 addresses)", n, d);
    run<4, 2, 128>("128 FFMA + 2 LDS.128 + 2 STS.128 (32 addresses)", n, d);
    run<1, 9, 256>("256 FFMA + 9 LDS.128 (4 addresses / warp) = two sinks per lane", n, d);
    run<0, 0, 256>("256 FFMA", n, d);
    run<1, 9, 64>("64 FFMA + 9 LDS.128 (4 addresses / warp)", n, d);
    run<0, 0, 64>("64 FFMA", n, d);
    return 0;
}
