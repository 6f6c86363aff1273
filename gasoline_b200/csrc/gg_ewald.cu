// gg_ewald.cu -- periodic Ewald correction (FP64), one thread per active sink particle.
//
// Replaces pkdEwaldInit (ewald.c:182-248, host side, once per call) and pkdBucketEwald (ewald.c:15-178).
// FP64 throughout: the gam[] recursion cancels for small alpha*r (ewald.c:98-136) and the whole correction is only
// ~154 terms per particle.  The root expansion uses COMPLETE multipoles (MEVAL meval.h:21-81); the trace
// combinations that MEVAL re-derives for every term depend on the moments only and are hoisted to the host.
#include <math.h>
#include <vector>
#include "gg_internal.h"

namespace {

struct Mom {
    double m, xx, yy, xy, xz, yz, zz;
    double O[10]; // xxx,xyy,xxy,yyy,xxz,yyz,xyz,xzz,yzz,zzz
    double H[15]; // xxxx,xyyy,xxxy,yyyy,xxxz,yyyz,xxyy,xxyz,xyyz,xxzz,xyzz,xzzz,yyzz,yzzz,zzzz
};

// The three contractions of the expansion with the displacement, via scaled monomials (see gg_tree_kernel.cu).
template <typename T>
__host__ __device__ inline void contract4(const double *H, T dx, T dy, T dz, T &hx, T &hy, T &hz) {
    T hxx = 0.5 * dx * dx, hyy = 0.5 * dy * dy, hzz = 0.5 * dz * dz, xy = dx * dy;
    T cxxx = (1.0 / 3.0) * hxx * dx, cyyy = (1.0 / 3.0) * hyy * dy, czzz = (1.0 / 3.0) * hzz * dz;
    T cxxy = hxx * dy, cxxz = hxx * dz, cxyy = hyy * dx, cyyz = hyy * dz, cxzz = hzz * dx, cyzz = hzz * dy,
      cxyz = xy * dz;
    hx = H[0] * cxxx + H[2] * cxxy + H[4] * cxxz + H[6] * cxyy + H[7] * cxyz + H[9] * cxzz + H[1] * cyyy +
         H[8] * cyyz + H[10] * cyzz + H[11] * czzz;
    hy = H[2] * cxxx + H[6] * cxxy + H[7] * cxxz + H[1] * cxyy + H[8] * cxyz + H[10] * cxzz + H[3] * cyyy +
         H[5] * cyyz + H[12] * cyzz + H[13] * czzz;
    hz = H[4] * cxxx + H[7] * cxxy + H[9] * cxxz + H[8] * cxyy + H[10] * cxyz + H[11] * cxzz + H[5] * cyyy +
         H[12] * cyyz + H[13] * cyzz + H[14] * czzz;
}
template <typename T>
__host__ __device__ inline void contract3(const double *O, T dx, T dy, T dz, T &ox, T &oy, T &oz) {
    T hxx = 0.5 * dx * dx, hyy = 0.5 * dy * dy, hzz = 0.5 * dz * dz, xy = dx * dy, xz = dx * dz, yz = dy * dz;
    ox = O[0] * hxx + O[2] * xy + O[4] * xz + O[1] * hyy + O[6] * yz + O[7] * hzz;
    oy = O[2] * hxx + O[1] * xy + O[6] * xz + O[3] * hyy + O[5] * yz + O[8] * hzz;
    oz = O[4] * hxx + O[6] * xy + O[7] * xz + O[5] * hyy + O[8] * yz + O[9] * hzz;
}

// MEVAL (meval.h:21-81) with the trace terms precomputed; accumulates into ax,ay,az,fPot.
__device__ __forceinline__ void meval(int iOrder, const EwaldKernelArgs &A, const double *g, double dx, double dy,
                                      double dz, double &ax, double &ay, double &az, double &fPot) {
    const double *R = A.root;
    double ta = 0.0;
    if (iOrder >= 4) {
        double hx, hy, hz;
        contract4(&R[20], dx, dy, dz, hx, hy, hz);
        double qr = 0.25 * (hx * dx + hy * dy + hz * dz);
        const double *T = A.trQ4;
        double Qhx = 0.5 * (T[0] * dx + T[1] * dy + T[2] * dz);
        double Qhy = 0.5 * (T[1] * dx + T[3] * dy + T[4] * dz);
        double Qhz = 0.5 * (T[2] * dx + T[4] * dy + T[5] * dz);
        double Qh = 0.5 * (Qhx * dx + Qhy * dy + Qhz * dz);
        fPot -= g[4] * qr - g[3] * Qh + g[2] * T[6];
        ta += g[5] * qr - g[4] * Qh + g[3] * T[6];
        ax += g[4] * hx - g[3] * Qhx;
        ay += g[4] * hy - g[3] * Qhy;
        az += g[4] * hz - g[3] * Qhz;
    }
    if (iOrder >= 3) {
        double ox, oy, oz;
        contract3(&R[10], dx, dy, dz, ox, oy, oz);
        double qr = (1.0 / 3.0) * (ox * dx + oy * dy + oz * dz);
        double Qtr = A.trQ3[0] * dx + A.trQ3[1] * dy + A.trQ3[2] * dz;
        fPot -= g[3] * qr - g[2] * Qtr;
        ta += g[4] * qr - g[3] * Qtr;
        ax += g[3] * ox - g[2] * A.trQ3[0];
        ay += g[3] * oy - g[2] * A.trQ3[1];
        az += g[3] * oz - g[2] * A.trQ3[2];
    }
    if (iOrder >= 2) {
        double qx = R[7] * dz + R[6] * dy + R[4] * dx;
        double qy = R[8] * dz + R[6] * dx + R[5] * dy;
        double qz = R[8] * dy + R[7] * dx + R[9] * dz;
        double qr = 0.5 * (qx * dx + qy * dy + qz * dz);
        fPot -= g[2] * qr - g[1] * A.trQ2;
        ta += g[3] * qr - g[2] * A.trQ2;
        ax += g[2] * qx;
        ay += g[2] * qy;
        az += g[2] * qz;
    }
    fPot -= g[0] * R[0];
    ta += g[1] * R[0];
    ax -= dx * ta;
    ay -= dy * ta;
    az -= dz * ta;
}

#ifndef GG_EWALD_EXPFAC
#define GG_EWALD_EXPFAC 0 // 1: exp(-alpha^2 r^2) as a product of per-axis factors (21 exponentials per particle instead of ~90).
                          // Measured: no change (128^3 3.453 vs 3.450 ms, 256^3 27.01 vs 27.00 ms) -- the kernel waits on its
                          // dependent FP64 chains (erfcx, the gam[] recursion, MEVAL), not on the instruction count.  Off.
#endif
#ifndef GG_EWALD_RSQRT
#define GG_EWALD_RSQRT 1
#endif
#ifndef GG_EWALD_MIN_CTAS
#define GG_EWALD_MIN_CTAS 5 // 96 registers, 20 warps per SM: 4.21 -> 3.70 ms on the 128^3 box (3 CTAs: 3.79, 6: 4.05)
#endif
__global__ void __launch_bounds__(128, GG_EWALD_MIN_CTAS) k_ewald(const EwaldKernelArgs A) {
    extern __shared__ double s_ewt[];
    for (int i = threadIdx.x; i < A.nEwh * 5; i += blockDim.x) s_ewt[i] = A.ewt[i];
    __syncthreads();
    const int i = A.first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    if (A.active && !A.active[i]) return;
    const double L = A.L;
    double fPot = A.root[0] * A.k1, ax = 0.0, ay = 0.0, az = 0.0;
    const double dx = A.parts[i].x - A.root[1], dy = A.parts[i].y - A.root[2], dz = A.parts[i].z - A.root[3];
    const int nE = A.nEwReps, nR = A.nReps;
    int nLoop = 0;
#if GG_EWALD_EXPFAC
    // exp(-alpha^2 r^2) = exp(-alpha^2 dx'^2) exp(-alpha^2 dy'^2) exp(-alpha^2 dz'^2) with dx' = dx + ix L: 3 (2 nE + 1)
    // exponentials per particle instead of one per accepted image (~90 of the 343), two multiplications per image
    constexpr int EW_TAB = 9; // nE <= 4
    const bool expFac = nE <= (EW_TAB - 1) / 2;
    double eyT[EW_TAB], ezT[EW_TAB];
    if (expFac)
        for (int k = -nE; k <= nE; ++k) {
            const double ty = dy + k * L, tz = dz + k * L;
            eyT[k + nE] = exp(-(ty * ty) * A.alpha2);
            ezT[k + nE] = exp(-(tz * tz) * A.alpha2);
        }
#endif
    for (int ix = -nE; ix <= nE; ++ix) {
        const bool holex = (ix >= -nR && ix <= nR);
        const double dxo = dx + ix * L;
#if GG_EWALD_EXPFAC
        const double exX = expFac ? exp(-(dxo * dxo) * A.alpha2) : 0.0;
#endif
        for (int iy = -nE; iy <= nE; ++iy) {
            const bool holexy = holex && (iy >= -nR && iy <= nR);
            const double dyo = dy + iy * L;
#if GG_EWALD_EXPFAC
            const double exXY = expFac ? exX * eyT[iy + nE] : 0.0;
#endif
            for (int iz = -nE; iz <= nE; ++iz) {
                const bool hole = holexy && (iz >= -nR && iz <= nR);
                const double dzo = dz + iz * L;
                double r2 = dxo * dxo + dyo * dyo + dzo * dzo;
                if (r2 > A.fEwCut2 && !hole) continue;
                double g[6];
                if (r2 < 3.0e-3 * L * L) { // series about the origin, ewald.c:98-117
                    double alphan = A.ka;
                    r2 *= A.alpha2;
                    g[0] = alphan * (r2 / 3 - 1); alphan *= 2 * A.alpha2;
                    g[1] = alphan * (r2 / 5 - 1.0 / 3.0); alphan *= 2 * A.alpha2;
                    g[2] = alphan * (r2 / 7 - 1.0 / 5.0); alphan *= 2 * A.alpha2;
                    g[3] = alphan * (r2 / 9 - 1.0 / 7.0); alphan *= 2 * A.alpha2;
                    g[4] = alphan * (r2 / 11 - 1.0 / 9.0); alphan *= 2 * A.alpha2;
                    g[5] = alphan * (r2 / 13 - 1.0 / 11.0);
                } else { // ewald.c:118-136
#if GG_EWALD_RSQRT
                    // 1/r from the FP64 reciprocal square root (MUFU.RSQ64H + Newton, <= 1 ulp) and r = r2 / r: spares the
                    // IEEE square root AND the division (together ~45 of the term's ~300 FP64-pipe instructions)
                    const double dir = rsqrt(r2), r = r2 * dir, dir2 = dir * dir;
#else
                    double r = sqrt(r2), dir = 1.0 / r, dir2 = dir * dir;
#endif
                    // erfc(x) = exp(-x^2) erfcx(x): the exponential is needed anyway (ewald.c:121), and the scaled function
                    // is the cheaper one; -erf(x) = erfc(x) - 1 (x > 0.1 here: the series branch took the small radii)
#if GG_EWALD_EXPFAC
                    const double ex = expFac ? exXY * ezT[iz + nE] : exp(-r2 * A.alpha2);
#else
                    const double ex = exp(-r2 * A.alpha2);
#endif
                    double a = ex * A.ka * dir2;
                    const double ec = ex * erfcx(A.alpha * r);
                    g[0] = (hole ? ec - 1.0 : ec) * dir;
                    double alphan = 2 * A.alpha2;
                    g[1] = g[0] * dir2 + a;
                    g[2] = 3 * g[1] * dir2 + alphan * a; alphan *= 2 * A.alpha2;
                    g[3] = 5 * g[2] * dir2 + alphan * a; alphan *= 2 * A.alpha2;
                    g[4] = 7 * g[3] * dir2 + alphan * a; alphan *= 2 * A.alpha2;
                    g[5] = 9 * g[4] * dir2 + alphan * a;
                }
                meval(A.iOrder, A, g, dxo, dyo, dzo, ax, ay, az, fPot);
                ++nLoop;
            }
        }
    }
    for (int k = 0; k < A.nEwh; ++k) { // k-space, ewald.c:156-164
        const double *e = &s_ewt[5 * k];
        double s, c;
        sincos(e[0] * dx + e[1] * dy + e[2] * dz, &s, &c);
        double t = e[3] * s - e[4] * c;
        fPot += e[3] * c + e[4] * s;
        ax += e[0] * t;
        ay += e[1] * t;
        az += e[2] * t;
    }
    A.pot[i] += fPot;
    A.acc[3 * (size_t)i] += ax;
    A.acc[3 * (size_t)i + 1] += ay;
    A.acc[3 * (size_t)i + 2] += az;
    A.nLoop[i] = nLoop;
}

// Per-bucket bookkeeping of pkdGravAll (pkd.c:2945-2998): interaction sums, the reference's flop score
// (grav.c:246-247, ewald.c:175-176) and fWeight for the active particles of the bucket.
__global__ void __launch_bounds__(256) k_stats(const StatsKernelArgs A) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    // v[0..5]: sums (nActive, part, cell, soft, flopI, flopE); v[6..8]: maxima of the per-bucket list lengths
    unsigned long long v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (t < A.nTasks && A.tasks[t].pass == 0) {
        const Task task = A.tasks[t];
        const NodeW bk = A.nodes[task.node];
        const int qflop[5] = {10, 10, 41, 120, 277}, mflop[5] = {10, 10, 48, 151, 343};
        int n = 0;
        long long nLoop = 0;
        for (int j = 0; j < bk.nP; ++j) {
            int pi = bk.pLower + j;
            if (A.active && !A.active[pi]) continue;
            ++n;
            if (A.nLoop) nLoop += A.nLoop[pi];
        }
        const int nP = A.counts[3 * task.node], nS = A.counts[3 * task.node + 1], nN = A.counts[3 * task.node + 2];
        const long long part = (long long)n * nP + (long long)(n * (2 * (bk.nP - 1) - n + 1) / 2); // pkd.c:2946-2947
        const long long flopI = (long long)n * ((long long)(nP + bk.nP) * 38 + (long long)nS * 82 +
                                                (long long)nN * (35 + qflop[A.iOrder]));
        const long long flopE = A.nLoop ? nLoop * (104 + mflop[A.iEwOrder]) + (long long)n * A.nEwh * 58 : 0;
        for (int j = 0; j < bk.nP; ++j) {
            int pi = bk.pLower + j;
            if (A.active && !A.active[pi]) continue;
            A.fWeight[pi] = (double)flopI + (double)flopE;
            if (A.hfWeight) A.hfWeight[pi] = (double)flopI + (double)flopE;
        }
        v[0] = (unsigned long long)n; v[1] = (unsigned long long)part; v[2] = (unsigned long long)((long long)n * nN);
        v[3] = (unsigned long long)((long long)n * nS); v[4] = (unsigned long long)flopI; v[5] = (unsigned long long)flopE;
        v[6] = (unsigned long long)nP; v[7] = (unsigned long long)nS; v[8] = (unsigned long long)nN;
    }
    // one atomic per CTA and quantity instead of one per bucket (the sums are integers: order does not matter)
    __shared__ unsigned long long s_v[8][9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long u = __shfl_xor_sync(0xffffffffu, v[k], o);
            v[k] = k < 6 ? v[k] + u : (u > v[k] ? u : v[k]);
        }
    }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 9; ++k) s_v[threadIdx.x >> 5][k] = v[k];
    __syncthreads();
    if (threadIdx.x < 9) {
        const int k = threadIdx.x;
        unsigned long long r = 0;
        for (int w = 0; w < 8; ++w) r = k < 6 ? r + s_v[w][k] : (s_v[w][k] > r ? s_v[w][k] : r);
        if (k < 6) atomicAdd(&A.sums[k], r);
        else atomicMax(&A.sums[k], r);
    }
}

} // namespace

cudaError_t gg_launch_ewald_kernel(const EwaldKernelArgs &a, cudaStream_t st) {
    if (a.n <= a.first) return cudaSuccess;
    size_t smem = (size_t)a.nEwh * 5 * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_ewald, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_ewald<<<(a.n - a.first + 127) / 128, 128, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t gg_launch_stats_kernel(const StatsKernelArgs &a, cudaStream_t st) {
    if (a.nTasks <= 0) return cudaSuccess;
    k_stats<<<(a.nTasks + 255) / 256, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// pkdEwaldInit (ewald.c:182-248) on the host: rows (2pi/L)h, hCfac, hSfac for 0 < |h|^2 <= fhCut^2; hCfac/hSfac
// are the reduced (QEVAL) contraction of the root moments with the even / odd members of gam[] only.
void gg_ewald_table_host(const double *root, double L, double fhCut, int iOrder, std::vector<double> &ewt) {
    ewt.clear();
    const int hReps = (int)ceil(fhCut);
    const double alpha = 2.0 / L, k4 = M_PI * M_PI / (alpha * alpha * L * L);
    for (int hx = -hReps; hx <= hReps; ++hx)
        for (int hy = -hReps; hy <= hReps; ++hy)
            for (int hz = -hReps; hz <= hReps; ++hz) {
                const int h2 = hx * hx + hy * hy + hz * hz;
                if (h2 == 0 || h2 > fhCut * fhCut) continue;
                double g[6];
                g[0] = exp(-k4 * h2) / (M_PI * h2 * L);
                g[1] = 2 * M_PI / L * g[0]; g[2] = -2 * M_PI / L * g[1]; g[3] = 2 * M_PI / L * g[2];
                g[4] = -2 * M_PI / L * g[3]; g[5] = 2 * M_PI / L * g[4];
                const double dx = hx, dy = hy, dz = hz;
                // reduced contraction; potential-like sums with even gam (cos factor) and odd gam (sin factor)
                double q4 = 0, q3 = 0, q2 = 0;
                if (iOrder >= 4) {
                    double a, b, c;
                    contract4(&root[20], dx, dy, dz, a, b, c);
                    q4 = 0.25 * (a * dx + b * dy + c * dz);
                }
                if (iOrder >= 3) {
                    double a, b, c;
                    contract3(&root[10], dx, dy, dz, a, b, c);
                    q3 = (1.0 / 3.0) * (a * dx + b * dy + c * dz);
                }
                if (iOrder >= 2) {
                    double a = root[7] * dz + root[6] * dy + root[4] * dx, b = root[8] * dz + root[6] * dx + root[5] * dy,
                           c = root[8] * dy + root[7] * dx + root[9] * dz;
                    q2 = 0.5 * (a * dx + b * dy + c * dz);
                }
                // QEVAL's potential: fPot -= g4*q4 + g3*q3 + g2*q2 + g0*m (qeval.h:31,42,52,58), from 0
                const double mfacc = -(g[4] * q4) - g[2] * q2 - g[0] * root[0];
                const double mfacs = -(g[3] * q3);
                const double k = 2 * M_PI / L;
                ewt.push_back(k * hx); ewt.push_back(k * hy); ewt.push_back(k * hz);
                ewt.push_back(mfacc); ewt.push_back(mfacs);
            }
}
