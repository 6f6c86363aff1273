// gg_orb.cu -- the per-rank halves of the reference's ORB domain decomposition on the device:
//   k_orb_bounds = pstCalcBound (pst.c:1937) / pkdCalcBound of the rank's particles, per cell of the rank tree (PST)
//   k_orb_weight = pstWeight / pkdWeight (pst.c:1405, pkd.c:945-989): how many of the rank's particles of a PST cell lie
//                  below / at-or-above a trial split (r[d] < fSplit is "low", pkdLowerPart pkd.c:1064), and their weight
//   k_orb_split  = the outcome of pkdColRejects/pkdSwapRejects after _pstRootSplit (pst.c:1275-1334): every particle of
//                  a split cell now belongs to LOWER(cell) = 2 cell or UPPER(cell) = 2 cell + 1
// The reference partitions pStore in place on every trial (pkdLowerPart/pkdUpperPart) to count; here the particles
// stay where they are and carry the heap index of their PST cell, so ALL cells of one level of the rank tree are
// weighed by one launch (one 8 B coordinate + 4 B cell id per particle and trial = HBM-bound integer/compare work;
// 1 M particles = 12 MB per trial).  Counts are integers (exact, order-free).  Weights are summed in a FIXED order
// (xor-shuffle tree per warp, warps of a CTA in order, CTAs in order by k_orb_weight_sum), so a result is reproducible
// run to run; the reference's own sum follows the history of its in-place partition and is not reproducible bit for
// bit by anyone -- with fWeight = 1 (every first decomposition of a run) all sums are exact integers.
#include "gg_internal.h"

#ifndef GG_ORB_PPT
#define GG_ORB_PPT 1 // particles per thread in the weighing kernels (4 measured slower: 3.4 -> 4.3 ms per 8-domain decomposition of 1 M)
#endif

namespace {

__device__ __forceinline__ unsigned long long enc(double v) {
    const long long b = __double_as_longlong(v);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long warp_min64(unsigned long long v) {
    const unsigned h = (unsigned)(v >> 32), mh = __reduce_min_sync(0xffffffffu, h);
    const unsigned ml = __reduce_min_sync(0xffffffffu, h == mh ? (unsigned)v : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ unsigned long long warp_max64(unsigned long long v) {
    const unsigned h = (unsigned)(v >> 32), mh = __reduce_max_sync(0xffffffffu, h);
    const unsigned ml = __reduce_max_sync(0xffffffffu, h == mh ? (unsigned)v : 0u);
    return ((unsigned long long)mh << 32) | ml;
}

// slot of every PST heap index in this query (-1: not asked about) -> shared memory
__device__ __forceinline__ void load_slots(const OrbQuery &Q, signed char *slotOf) {
    for (int i = threadIdx.x; i < GG_ORB_MAX_CELL; i += blockDim.x) slotOf[i] = -1;
    __syncthreads();
    for (int s = threadIdx.x; s < Q.nSlots; s += blockDim.x) slotOf[Q.cell[s]] = (signed char)s;
    __syncthreads();
}

__global__ void __launch_bounds__(256) k_orb_init(int n, int *cellOf) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) cellOf[i] = 1; // ROOT of the rank tree (pkd.h:77)
}

// out[slot][0..2] = min, [3..5] = max as ordered integers (initialised to +inf / -inf keys), cnt[slot] = particles
__global__ void __launch_bounds__(256) k_orb_bounds(const OrbQuery Q, int n, const double *x, const double *y, const double *z,
                                                    const int *cellOf, unsigned long long *out, int *cnt) {
    __shared__ signed char slotOf[GG_ORB_MAX_CELL];
    load_slots(Q, slotOf);
    const int i = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
    int mine = -1;
    double p[3] = {0.0, 0.0, 0.0};
    if (i < n) {
        mine = slotOf[cellOf[i]];
        p[0] = x[i]; p[1] = y[i]; p[2] = z[i];
    }
    unsigned todo = __ballot_sync(0xffffffffu, mine >= 0);
    while (todo) {
        const int s = __shfl_sync(0xffffffffu, mine, __ffs(todo) - 1);
        const bool in = mine == s;
        const unsigned m = __ballot_sync(0xffffffffu, in);
        todo &= ~m;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const unsigned long long lo = warp_min64(in ? enc(p[k]) : ~0ull), hi = warp_max64(in ? enc(p[k]) : 0ull);
            if (lane == 0) {
                atomicMin(&out[6 * s + k], lo);
                atomicMax(&out[6 * s + 3 + k], hi);
            }
        }
        if (lane == 0) atomicAdd(&cnt[s], __popc(m));
    }
}

// cnt[slot][2] += (low, high) particle counts; part[blockIdx][slot][2] = this CTA's (low, high) weight (w != null)
__global__ void __launch_bounds__(256) k_orb_weight(const OrbQuery Q, int n, const double *x, const double *y, const double *z,
                                                    const double *w, const int *cellOf, int *cnt, double *part) {
    __shared__ signed char slotOf[GG_ORB_MAX_CELL];
    __shared__ int sCnt[GG_ORB_MAX_SLOTS][2];
    __shared__ double sW[8][GG_ORB_MAX_SLOTS][2];
    load_slots(Q, slotOf);
    for (int t = threadIdx.x; t < 2 * Q.nSlots; t += 256) {
        sCnt[t >> 1][t & 1] = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) sW[k][t >> 1][t & 1] = 0.0;
    }
    __syncthreads();
    // GG_ORB_PPT particles per thread, a CTA's particles contiguous; the same geometry as k_orb_weight_d, so that the
    // host-driven and the device-driven bisection add the weights in the same order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = blockIdx.x * (256 * GG_ORB_PPT) + threadIdx.x;
    int mineA[GG_ORB_PPT];
    double cA[GG_ORB_PPT], wA[GG_ORB_PPT];
#pragma unroll
    for (int u = 0; u < GG_ORB_PPT; ++u) {
        const int i = base + u * 256;
        mineA[u] = i < n ? slotOf[cellOf[i]] : -1;
    }
#pragma unroll
    for (int u = 0; u < GG_ORB_PPT; ++u) {
        const int i = base + u * 256;
        cA[u] = 0.0; wA[u] = 0.0;
        if (mineA[u] >= 0) {
            const int d = Q.dim[mineA[u]];
            cA[u] = d == 0 ? x[i] : (d == 1 ? y[i] : z[i]);
            if (w) wA[u] = w[i];
        }
    }
#pragma unroll
    for (int u = 0; u < GG_ORB_PPT; ++u) {
        const int mine = mineA[u];
        const bool low = mine >= 0 && cA[u] < Q.split[mine];
        const double wi = wA[u];
        unsigned todo = __ballot_sync(0xffffffffu, mine >= 0);
        while (todo) {
            const int s = __shfl_sync(0xffffffffu, mine, __ffs(todo) - 1);
            const bool in = mine == s;
            const unsigned m = __ballot_sync(0xffffffffu, in), ml = __ballot_sync(0xffffffffu, in && low);
            todo &= ~m;
            if (lane == 0) {
                atomicAdd(&sCnt[s][0], __popc(ml));
                atomicAdd(&sCnt[s][1], __popc(m & ~ml));
            }
            if (w) {
                double a = in && low ? wi : 0.0, b = in && !low ? wi : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a = __dadd_rn(a, __shfl_xor_sync(0xffffffffu, a, o));
                    b = __dadd_rn(b, __shfl_xor_sync(0xffffffffu, b, o));
                }
                if (lane == 0) { sW[warp][s][0] = __dadd_rn(sW[warp][s][0], a); sW[warp][s][1] = __dadd_rn(sW[warp][s][1], b); }
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * Q.nSlots; t += 256) {
        const int s = t >> 1, side = t & 1;
        if (sCnt[s][side]) atomicAdd(&cnt[2 * s + side], sCnt[s][side]);
        if (w) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) a = __dadd_rn(a, sW[k][s][side]);
            part[((size_t)blockIdx.x * GG_ORB_MAX_SLOTS + s) * 2 + side] = a;
        }
    }
}

// sums[slot][side] = the CTAs' partial weights added in a fixed order: thread t takes CTAs t, t+256, ... in order, then
// the 256 threads are combined by a fixed tree
__global__ void __launch_bounds__(256) k_orb_weight_sum(int nSlots, int nBlocks, const double *part, double *sums) {
    __shared__ double sh[256];
    const int s = blockIdx.x >> 1, side = blockIdx.x & 1;
    double a = 0.0;
    for (int b = threadIdx.x; b < nBlocks; b += 256) a = __dadd_rn(a, part[((size_t)b * GG_ORB_MAX_SLOTS + s) * 2 + side]);
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[2 * s + side] = sh[0];
}

__global__ void __launch_bounds__(256) k_orb_split(const OrbQuery Q, int n, const double *x, const double *y, const double *z,
                                                   int *cellOf) {
    __shared__ signed char slotOf[GG_ORB_MAX_CELL];
    load_slots(Q, slotOf);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int c = cellOf[i], s = slotOf[c];
    if (s < 0) return;
    const int d = Q.dim[s];
    const double v = d == 0 ? x[i] : (d == 1 ? y[i] : z[i]);
    cellOf[i] = 2 * c + (v < Q.split[s] ? 0 : 1); // LOWER / UPPER (pkd.h:78-79)
}

// The split with a second boundary (the store-overflow outcome of pkdColRejects, pkd.c:1463-1485): the lower child takes
// the WRAPPED interval between fSplitInactive and fSplit (pkdLowerPartWrap, pkd.c:1165-1211), the upper child the rest.
__global__ void __launch_bounds__(256) k_orb_split_wrap(const OrbQuery Q, const OrbWrap W, int n, const double *x, const double *y,
                                                        const double *z, int *cellOf) {
    __shared__ signed char slotOf[GG_ORB_MAX_CELL];
    load_slots(Q, slotOf);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int c = cellOf[i], s = slotOf[c];
    if (s < 0) return;
    const int d = Q.dim[s];
    const double v = d == 0 ? x[i] : (d == 1 ? y[i] : z[i]);
    const double fS = Q.split[s], fI = W.inactive[s];
    const bool low = fI > fS ? (v < fS || v >= fI) : (v < fS && v >= fI);
    cellOf[i] = 2 * c + (low ? 0 : 1);
}

// ---- the bisection with its state on the device ---------------------------------------------------------------------
// k_orb_weight with the query read from the device state: cells that are no longer bisected have no slot
__global__ void __launch_bounds__(256) k_orb_weight_d(const OrbBisect *B, int n, const double *x, const double *y, const double *z,
                                                      const double *w, const int *cellOf, int *cnt, double *part) {
    if (B->nLive == 0) return;
    __shared__ signed char slotOf[GG_ORB_MAX_CELL];
    __shared__ int sCnt[GG_ORB_MAX_SLOTS][2], sDim[GG_ORB_MAX_SLOTS];
    __shared__ double sSplit[GG_ORB_MAX_SLOTS];
    __shared__ double sW[8][GG_ORB_MAX_SLOTS][2];
    const int nSlots = B->q.nSlots;
    for (int i = threadIdx.x; i < GG_ORB_MAX_CELL; i += blockDim.x) slotOf[i] = -1;
    __syncthreads();
    for (int s = threadIdx.x; s < nSlots; s += blockDim.x) {
        if (B->live[s]) slotOf[B->q.cell[s]] = (signed char)s;
        sDim[s] = B->q.dim[s];
        sSplit[s] = B->q.split[s];
    }
    for (int t = threadIdx.x; t < 2 * nSlots; t += 256) {
        sCnt[t >> 1][t & 1] = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) sW[k][t >> 1][t & 1] = 0.0;
    }
    __syncthreads();
    // GG_ORB_PPT particles per thread, a CTA's particles contiguous (the weights' summation order is a function of the
    // launch geometry only): the cell ids of all of them are loaded first, then the coordinates -- four loads in flight
    // per thread instead of one (the one-particle kernel is latency-bound: profiles/r01_k_orb_weight.md)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = blockIdx.x * (256 * GG_ORB_PPT) + threadIdx.x;
    int mineA[GG_ORB_PPT];
    double cA[GG_ORB_PPT], wA[GG_ORB_PPT];
#pragma unroll
    for (int u = 0; u < GG_ORB_PPT; ++u) {
        const int i = base + u * 256;
        mineA[u] = i < n ? slotOf[cellOf[i]] : -1;
    }
#pragma unroll
    for (int u = 0; u < GG_ORB_PPT; ++u) {
        const int i = base + u * 256;
        cA[u] = 0.0; wA[u] = 0.0;
        if (mineA[u] >= 0) {
            const int d = sDim[mineA[u]];
            cA[u] = d == 0 ? x[i] : (d == 1 ? y[i] : z[i]);
            if (w) wA[u] = w[i];
        }
    }
#pragma unroll
    for (int u = 0; u < GG_ORB_PPT; ++u) {
        const int mine = mineA[u];
        const bool low = mine >= 0 && cA[u] < sSplit[mine];
        const double wi = wA[u];
        unsigned todo = __ballot_sync(0xffffffffu, mine >= 0);
        while (todo) {
            const int s = __shfl_sync(0xffffffffu, mine, __ffs(todo) - 1);
            const bool in = mine == s;
            const unsigned m = __ballot_sync(0xffffffffu, in), ml = __ballot_sync(0xffffffffu, in && low);
            todo &= ~m;
            if (lane == 0) {
                atomicAdd(&sCnt[s][0], __popc(ml));
                atomicAdd(&sCnt[s][1], __popc(m & ~ml));
            }
            if (w) {
                double a = in && low ? wi : 0.0, b = in && !low ? wi : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a = __dadd_rn(a, __shfl_xor_sync(0xffffffffu, a, o));
                    b = __dadd_rn(b, __shfl_xor_sync(0xffffffffu, b, o));
                }
                if (lane == 0) { sW[warp][s][0] = __dadd_rn(sW[warp][s][0], a); sW[warp][s][1] = __dadd_rn(sW[warp][s][1], b); }
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * nSlots; t += 256) {
        const int s = t >> 1, side = t & 1;
        if (sCnt[s][side]) atomicAdd(&cnt[2 * s + side], sCnt[s][side]);
        if (w) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) a = __dadd_rn(a, sW[k][s][side]);
            part[((size_t)blockIdx.x * GG_ORB_MAX_SLOTS + s) * 2 + side] = a;
        }
    }
}

__global__ void __launch_bounds__(256) k_orb_weight_sum_d(const OrbBisect *B, int nBlocks, const double *part, double *sums) {
    const int s = blockIdx.x >> 1, side = blockIdx.x & 1;
    if (B->nLive == 0 || s >= B->q.nSlots || !B->live[s]) return;
    __shared__ double sh[256];
    double a = 0.0;
    for (int b = threadIdx.x; b < nBlocks; b += 256) a = __dadd_rn(a, part[((size_t)b * GG_ORB_MAX_SLOTS + s) * 2 + side]);
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[2 * s + side] = sh[0];
}

// One step of the root finder for every cell of the level (one thread per cell): digest the answer to the last trial
// (pst.c:1001-1030: stop on a one-one split or equal loads, else move the bracket's upper or lower end to the trial),
// then set up the next one (pst.c:984-999: while the midpoint lies strictly inside the bracket and ittr < MAX_ITTR).
// Several ranks (gg_orb_bisect_all): `all` holds every rank's answer record (GG_ORB_REC_BYTES each: sums, then counts) as
// gathered for this trial; the records are added in rank order, so every rank takes the same branch with the same bits --
// what the reference gets from adding outWtLow and outWtHigh up the rank tree (pst.c:1004-1010).
__global__ void __launch_bounds__(GG_ORB_MAX_SLOTS) k_orb_decide(OrbBisect *B, int *cnt, double *sums, int first, int useW,
                                                                 int nRanks, const unsigned char *all) {
    __shared__ int nLive;
    const int s = threadIdx.x;
    if (s == 0) nLive = 0;
    __syncthreads();
    if (!first && B->nLive == 0) return;
    if (s < B->q.nSlots) {
        int live = B->live[s];
        double fl = B->fl[s], fu = B->fu[s], fmm = B->fmm[s];
        if (!first && live) {
            int nLow = cnt[2 * s], nHigh = cnt[2 * s + 1];
            double wl = useW ? sums[2 * s] : 0.0, wh = useW ? sums[2 * s + 1] : 0.0;
            if (nRanks > 1) {
                nLow = nHigh = 0;
                wl = wh = 0.0;
                for (int r = 0; r < nRanks; ++r) {
                    const unsigned char *rec = all + (size_t)r * GG_ORB_REC_BYTES;
                    const double *rs = reinterpret_cast<const double *>(rec);
                    const int *rc = reinterpret_cast<const int *>(rec + GG_ORB_REC_CNT);
                    nLow += rc[2 * s]; nHigh += rc[2 * s + 1];
                    if (useW) {
                        wl = r ? __dadd_rn(wl, rs[2 * s]) : rs[2 * s];
                        wh = r ? __dadd_rn(wh, rs[2 * s + 1]) : rs[2 * s + 1];
                    }
                }
            }
            if (!useW) { wl = (double)nLow; wh = (double)nHigh; }
            const double a = B->splitWork ? __ddiv_rn(wl, B->nLower[s]) : __ddiv_rn((double)nLow, B->nLower[s]);
            const double b = B->splitWork ? __ddiv_rn(wh, B->nUpper[s]) : __ddiv_rn((double)nHigh, B->nUpper[s]);
            if ((nLow == 1 && nHigh == 1) || a == b) live = 0;
            else {
                if (a > b) fu = B->q.split[s];
                else fl = B->q.split[s];
                fmm = __dmul_rn(__dadd_rn(fl, fu), 0.5);
                B->ittr[s] += 1;
            }
        }
        if (live) live = (fl < fmm) && (fmm < fu) && (B->ittr[s] < B->maxIttr);
        if (live) {
            B->q.split[s] = fmm;
            B->hasSplit[s] = 1;
            atomicAdd(&nLive, 1);
        }
        B->live[s] = live; B->fl[s] = fl; B->fu[s] = fu; B->fmm[s] = fmm;
        cnt[2 * s] = cnt[2 * s + 1] = 0;
        if (useW) sums[2 * s] = sums[2 * s + 1] = 0.0;
    }
    __syncthreads();
    if (s == 0) B->nLive = nLive;
}

inline int blocks_for(int n) { return (n + 255) / 256; }

} // namespace

cudaError_t gg_launch_orb_bisect(OrbBisect *B, const OrbBisect &h, int n, const double *x, const double *y, const double *z,
                                 const double *w, const int *cellOf, int *cnt, double *part, double *sums, cudaStream_t st) {
    cudaError_t e = gg_launch_orb_bisect_begin(B, h, cnt, sums, w != nullptr, st);
    if (e != cudaSuccess) return e;
    // MAX_ITTR trials at most; once no cell is live the remaining launches return at their first instruction
    for (int t = 0; t <= h.maxIttr; ++t) {
        if ((e = gg_launch_orb_trial(B, h.q.nSlots, n, x, y, z, w, cellOf, cnt, part, sums, st)) != cudaSuccess) return e;
        if ((e = gg_launch_orb_decide(B, cnt, sums, w != nullptr, 1, nullptr, st)) != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// The pieces of the bisection, for a caller that puts a collective between a trial's weighing and its decision:
// state to the device + the first trial's set-up; this rank's answer to the current trial; the decision.
cudaError_t gg_launch_orb_bisect_begin(OrbBisect *B, const OrbBisect &h, int *cnt, double *sums, int useW, cudaStream_t st) {
    cudaError_t e = cudaMemcpyAsync(B, &h, sizeof(OrbBisect), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    k_orb_decide<<<1, GG_ORB_MAX_SLOTS, 0, st>>>(B, cnt, sums, 1, useW, 1, nullptr);
    return cudaGetLastError();
}

cudaError_t gg_launch_orb_trial(OrbBisect *B, int nSlots, int n, const double *x, const double *y, const double *z,
                                const double *w, const int *cellOf, int *cnt, double *part, double *sums, cudaStream_t st) {
    if (n > 0) {
        const int nb = (n + 256 * GG_ORB_PPT - 1) / (256 * GG_ORB_PPT);
        k_orb_weight_d<<<nb, 256, 0, st>>>(B, n, x, y, z, w, cellOf, cnt, part);
        if (w) k_orb_weight_sum_d<<<2 * nSlots, 256, 0, st>>>(B, nb, part, sums);
    }
    return cudaGetLastError();
}

cudaError_t gg_launch_orb_decide(OrbBisect *B, int *cnt, double *sums, int useW, int nRanks, const unsigned char *all,
                                 cudaStream_t st) {
    k_orb_decide<<<1, GG_ORB_MAX_SLOTS, 0, st>>>(B, cnt, sums, 0, useW, nRanks, all);
    return cudaGetLastError();
}

cudaError_t gg_launch_orb_init(int n, int *cellOf, cudaStream_t st) {
    if (n > 0) k_orb_init<<<blocks_for(n), 256, 0, st>>>(n, cellOf);
    return cudaGetLastError();
}

// out: [nSlots][6] ordered-integer keys, cnt: [nSlots]; both initialised here
cudaError_t gg_launch_orb_bounds(const OrbQuery &q, int n, const double *x, const double *y, const double *z, const int *cellOf,
                                 unsigned long long *out, int *cnt, cudaStream_t st) {
    unsigned long long init[GG_ORB_MAX_SLOTS * 6];
    for (int s = 0; s < q.nSlots; ++s)
        for (int k = 0; k < 6; ++k) init[6 * s + k] = k < 3 ? ~0ull : 0ull;
    cudaError_t e = cudaMemcpyAsync(out, init, sizeof(unsigned long long) * 6 * q.nSlots, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(cnt, 0, sizeof(int) * q.nSlots, st);
    if (e != cudaSuccess) return e;
    if (n > 0) k_orb_bounds<<<blocks_for(n), 256, 0, st>>>(q, n, x, y, z, cellOf, out, cnt);
    return cudaGetLastError();
}

// cnt: [nSlots][2] ints, sums: [nSlots][2] doubles (only written when w != null), part: [blocks_for(n)][MAX_SLOTS][2]
cudaError_t gg_launch_orb_weight(const OrbQuery &q, int n, const double *x, const double *y, const double *z, const double *w,
                                 const int *cellOf, int *cnt, double *part, double *sums, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int) * 2 * q.nSlots, st);
    if (e != cudaSuccess) return e;
    if (w) {
        e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * q.nSlots, st);
        if (e != cudaSuccess) return e;
    }
    if (n > 0) {
        const int nb = (n + 256 * GG_ORB_PPT - 1) / (256 * GG_ORB_PPT);
        k_orb_weight<<<nb, 256, 0, st>>>(q, n, x, y, z, w, cellOf, cnt, part);
        if (w) k_orb_weight_sum<<<2 * q.nSlots, 256, 0, st>>>(q.nSlots, nb, part, sums);
    }
    return cudaGetLastError();
}

cudaError_t gg_launch_orb_split(const OrbQuery &q, int n, const double *x, const double *y, const double *z, int *cellOf,
                                cudaStream_t st) {
    if (n > 0) k_orb_split<<<blocks_for(n), 256, 0, st>>>(q, n, x, y, z, cellOf);
    return cudaGetLastError();
}

cudaError_t gg_launch_orb_split_wrap(const OrbQuery &q, const OrbWrap &w, int n, const double *x, const double *y, const double *z,
                                     int *cellOf, cudaStream_t st) {
    if (n > 0) k_orb_split_wrap<<<blocks_for(n), 256, 0, st>>>(q, w, n, x, y, z, cellOf);
    return cudaGetLastError();
}

size_t gg_orb_part_bytes(int n) { return sizeof(double) * 2 * GG_ORB_MAX_SLOTS * (size_t)(n > 0 ? blocks_for(n) : 1); }
