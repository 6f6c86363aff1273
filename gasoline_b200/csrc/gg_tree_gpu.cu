// gg_tree_gpu.cu -- the gravity tree built ON THE DEVICE (pkdBuildBinary + pkdCalcCell + pkdCalcOpen +
// pkdThreadTree on the GPU), bit-identical to the reference's host build.
//
// What "identical" rests on (reference: BuildBinary pkd.c:2437-2587, pkdUpperPart pkd.c:1106-1133, pkdCombine pkd.c:1973,
// pkdCalcCell pkd.c:2018-2135, OPEN_JOSH pkd.c:2253-2260, pkdThreadTree pkd.c:2590-2620):
//   * the SHAPE of the tree depends only on exact min/max (squeezed bounds), one rounded midpoint and `<` comparisons;
//   * the ORDER of the particles inside a bucket fixes the summation order of its centre of mass, and with it the last
//     bit of every centre and opening radius above it.  The reference's partition exchanges the k-th element >= split
//     found from the left with the k-th element < split found from the right until the pointers cross at
//     mid = lo + #(< split).  That permutation is a function of prefix counts: element i < mid with r >= split is the
//     k-th "misplaced left" one with k = (i - lo) - #(< split in [lo, i)); element i >= mid with r < split is the k-th
//     "misplaced right" one with k = #(< split in (i, hi]).  One exclusive scan per level gives every particle its k,
//     and pair k is swapped -- the same array the serial loop leaves, for every cell of a tree level at once;
//   * centre of mass, mass and mass-weighted softening: buckets sum their particles in that order, interior cells
//     combine (lower, upper) child, with the reference's operation order and no FMA contraction (__dmul_rn/__dadd_rn);
//   * Bmax = max |r_p - r_cell| over ALL particles of the cell is a maximum (order-free); each distance is computed
//     with the reference's expression; fOpen2 = max(Bmax, 2/sqrt(3) Bmax/theta)^2;
//   * cells are numbered in depth-first pre-order, as pkd->iFreeCell++ numbers them during the recursion.
// The multipole moments come from gg_moments.cu (bottom-up M2M, FP64), exactly as for gg_set_local with mom = NULL.
//
// Level-synchronous construction, three kernels + one scan per tree level over all n particles (cellOf[i] = the deepest
// cell holding position i):  k_flag (split decision recomputed per particle from the cell's bounds; flag = r[dim] < split;
// the first particle of a cell allocates the two children) -> cub exclusive scan -> k_split (ranks -> exchange table;
// children's ranges; children's squeezed bounds by ordered-integer atomic min/max, block/warp pre-reduced -- a particle's
// side is its flag, wherever the exchange will put it) -> k_swap (exchange pair k, re-home position i).  A cell that
// still splits but holds <= GGB_WCAP particles leaves this path: k_finish builds its whole sub-tree with one warp in
// shared memory (same rules, ballots instead of the scan), which removes the deepest levels of launches (Plummer 1 M:
// 34 -> 28 levels).  Then one bottom-up kernel (arrival counters, like gg_moments.cu), one numbering kernel (pre-order index
// and threaded "next" from subtree sizes), Bmax by warp-aggregated atomic max while climbing, and the emit kernel.
#include <cub/cub.cuh>
#include <math.h>
#include <stdio.h>
#include "gg_internal.h"

namespace {

struct __align__(8) BNode { // construction record, breadth-first numbering
    int lo, hi;             // particle range (inclusive)
    int left, right;        // children (breadth-first ids), -1: bucket
    int parent;
    int dim;                // split axis, -1: not split
    int mid;                // first particle of the upper child
    int nMis;               // exchanged pairs
    double split;
    unsigned long long b[6]; // fMin[3], fMax[3] as order-preserving integers
};

#define GGB_MAX_LEVELS 192
#ifndef GGB_WCAP
#define GGB_WCAP 128 // largest cell (particles) finished by one warp in shared memory (64 / 128 / 256 measured: 3.08 / 2.94 / 2.98 ms)
#endif

__device__ __forceinline__ unsigned long long enc(double v) {
    const long long b = __double_as_longlong(v);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec(unsigned long long k) {
    return __longlong_as_double((k & 0x8000000000000000ull) ? (long long)(k & 0x7fffffffffffffffull) : (long long)~k);
}

// min/max of six ordered-integer keys over the lanes of a warp: two 32-bit REDUX per 64-bit key (high halves, then the
// low halves of the lanes that hold the winning high half) instead of ten shuffles
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
__device__ __forceinline__ void warp_minmax(unsigned long long *lo, unsigned long long *hi) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = warp_min64(lo[k]); hi[k] = warp_max64(hi[k]); }
}

__global__ void k_b_init(int n, int *iord, int *cellOf, BNode *nodes, int *ctr, int *levelStart) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { iord[i] = i; cellOf[i] = 0; }
    if (i == 0) {
        BNode r;
        r.lo = 0; r.hi = n - 1; r.left = r.right = -1; r.parent = -1; r.dim = -1; r.mid = 0; r.nMis = 0; r.split = 0.0;
        for (int k = 0; k < 3; ++k) { r.b[k] = ~0ull; r.b[3 + k] = 0ull; }
        nodes[0] = r;
        ctr[0] = 1;
        ctr[1] = 0; // cells handed to k_finish
        levelStart[0] = 0; // the root is fresh at level 0
        levelStart[1] = 1; // cells allocated by level 0 start here
    }
}

// squeezed bounds of the root (level 0); the children's bounds are accumulated by k_split of the level that makes them
__global__ void __launch_bounds__(256) k_bounds(int n, const double *x, const double *y, const double *z, const int *cellOf,
                                                const int *levelStart, int level, BNode *nodes) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int first = levelStart[level];
    int c = -1;
    unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
    if (i < n) {
        c = cellOf[i];
        if (c < first) c = -1;
        else {
            lo[0] = hi[0] = enc(x[i]); lo[1] = hi[1] = enc(y[i]); lo[2] = hi[2] = enc(z[i]);
        }
    }
    __shared__ int s_c;
    __shared__ unsigned long long s_v[8][6];
    if (threadIdx.x == 0) s_c = c;
    __syncthreads();
    const int cb = s_c;
    const bool blockUniform = __syncthreads_and(c == cb);
    if (blockUniform && cb < 0) return;
    const int c0 = __shfl_sync(0xffffffffu, c, 0);
    const bool warpUniform = __all_sync(0xffffffffu, c == c0);
    if (warpUniform) warp_minmax(lo, hi);
    if (blockUniform) {
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0)
            for (int k = 0; k < 3; ++k) { s_v[w][k] = lo[k]; s_v[w][3 + k] = hi[k]; }
        __syncthreads();
        if (threadIdx.x < 6) {
            unsigned long long v = s_v[0][threadIdx.x];
            for (int q = 1; q < 8; ++q) {
                const unsigned long long u = s_v[q][threadIdx.x];
                v = threadIdx.x < 3 ? (u < v ? u : v) : (u > v ? u : v);
            }
            if (threadIdx.x < 3) atomicMin(&nodes[cb].b[threadIdx.x], v);
            else atomicMax(&nodes[cb].b[threadIdx.x], v);
        }
        return;
    }
    if (warpUniform) {
        if (c0 >= 0 && (threadIdx.x & 31) == 0)
            for (int k = 0; k < 3; ++k) { atomicMin(&nodes[c0].b[k], lo[k]); atomicMax(&nodes[c0].b[3 + k], hi[k]); }
        return;
    }
    if (c >= 0)
        for (int k = 0; k < 3; ++k) { atomicMin(&nodes[c].b[k], lo[k]); atomicMax(&nodes[c].b[3 + k], hi[k]); }
}

// BuildBinary's decision for a fresh cell (pkd.c:2437-2587): split the longest axis of the squeezed box at its midpoint
// when the cell holds more than nBucket particles and has extent; first axis wins ties.
__device__ __forceinline__ int decide(const BNode &nd, int nBucket, double *pSplit) {
    const double mn0 = dec(nd.b[0]), mn1 = dec(nd.b[1]), mn2 = dec(nd.b[2]);
    const double mx0 = dec(nd.b[3]), mx1 = dec(nd.b[4]), mx2 = dec(nd.b[5]);
    const bool good = (mx0 > mn0) || (mx1 > mn1) || (mx2 > mn2);
    if (!(nd.hi - nd.lo + 1 > nBucket && good)) return -1;
    int d = 0;
    double e = __dsub_rn(mx0, mn0), lo = mn0, hi = mx0;
    const double e1 = __dsub_rn(mx1, mn1), e2 = __dsub_rn(mx2, mn2);
    if (e1 > e) { d = 1; e = e1; lo = mn1; hi = mx1; }
    if (e2 > e) { d = 2; lo = mn2; hi = mx2; }
    *pSplit = __dmul_rn(0.5, __dadd_rn(lo, hi));
    return d;
}

__global__ void __launch_bounds__(256) k_flag(int n, const double *x, const double *y, const double *z, const int *cellOf,
                                              const int *levelStart, int level, BNode *nodes, int nBucket, int *flag,
                                              int *ctr, int *retired) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i > n) return;
    int f = 0;
    if (i < n) {
        const int c = cellOf[i];
        if (c >= levelStart[level]) {
            // only the range and the bounds are read: the cell's first particle writes dim / split / children below while
            // the others are still here
            BNode nd;
            nd.lo = nodes[c].lo; nd.hi = nodes[c].hi;
#pragma unroll
            for (int k = 0; k < 6; ++k) nd.b[k] = nodes[c].b[k];
            double split;
            int d = decide(nd, nBucket, &split);
            // a cell that still splits but holds <= GGB_WCAP particles leaves the level-synchronous path here: one warp
            // of k_finish builds its whole sub-tree in shared memory
            if (d >= 0 && nd.hi - nd.lo + 1 <= GGB_WCAP) d = -2;
            if (d >= 0) {
                const double v = d == 0 ? x[i] : (d == 1 ? y[i] : z[i]);
                f = v < split;
            }
            if (i == nd.lo) { // one thread per cell records the decision and, for a split, allocates the two children
                nodes[c].dim = d;
                nodes[c].split = d >= 0 ? split : 0.0;
                if (d == -2) retired[atomicAdd(&ctr[1], 1)] = c;
                if (d >= 0) {
                    const int id = atomicAdd(&ctr[0], 2);
                    BNode ch;
                    ch.lo = ch.hi = 0; // filled by k_split once the scan has placed the boundary
                    ch.left = ch.right = -1; ch.parent = c; ch.dim = -1; ch.mid = 0; ch.nMis = 0; ch.split = 0.0;
                    for (int k = 0; k < 3; ++k) { ch.b[k] = ~0ull; ch.b[3 + k] = 0ull; }
                    nodes[id] = ch;
                    nodes[id + 1] = ch;
                    nodes[c].left = id;
                    nodes[c].right = id + 1;
                }
            }
        }
    }
    flag[i] = f; // flag[n] = 0 closes the exclusive scan
}

// Ranks of the misplaced particles -> exchange table; the squeezed bounds of the two children (a particle goes below
// iff its flag is set, wherever the exchange puts it); the first particle of the cell fills in the children's ranges.
__global__ void __launch_bounds__(256) k_split(int n, const double *x, const double *y, const double *z, const int *cellOf,
                                               int *levelStart, int level, BNode *nodes, const int *flag, const int *S,
                                               int *tabL, int *tabR, const int *ctr) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i == 0) levelStart[level + 2] = ctr[0]; // every allocation of this level happened in k_flag
    int c = -1, less = 0;
    if (i < n) {
        c = cellOf[i];
        if (c < levelStart[level] || nodes[c].dim < 0) c = -1;
    }
    unsigned long long lo0[3] = {~0ull, ~0ull, ~0ull}, hi0[3] = {0ull, 0ull, 0ull}, lo1[3] = {~0ull, ~0ull, ~0ull},
                       hi1[3] = {0ull, 0ull, 0ull}, key[3] = {0ull, 0ull, 0ull};
    int left = -1;
    if (c >= 0) {
        const int clo = nodes[c].lo, chi = nodes[c].hi;
        left = nodes[c].left;
        const int base = S[clo], nLeft = S[chi + 1] - base, mid = clo + nLeft;
        less = flag[i];
        if (i < mid) {
            if (!less) tabL[clo + (i - clo) - (S[i] - base)] = i;
        } else if (less) tabR[clo + (S[chi + 1] - S[i + 1])] = i;
        if (i == clo) {
            nodes[left].lo = clo; nodes[left].hi = mid - 1;
            nodes[left + 1].lo = mid; nodes[left + 1].hi = chi;
            nodes[c].mid = mid;
            nodes[c].nMis = (mid - clo) - (S[mid] - base);
        }
        key[0] = enc(x[i]); key[1] = enc(y[i]); key[2] = enc(z[i]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (less) lo0[k] = hi0[k] = key[k];
            else lo1[k] = hi1[k] = key[k];
        }
    }
    // children's bounds: pre-reduce over the block / the warp when all of it stands in one cell
    __shared__ int s_c;
    __shared__ unsigned long long s_v[8][12];
    if (threadIdx.x == 0) s_c = c;
    __syncthreads();
    const int cb = s_c;
    const bool blockUniform = __syncthreads_and(c == cb);
    if (blockUniform && cb < 0) return;
    const int c0 = __shfl_sync(0xffffffffu, c, 0);
    const bool warpUniform = __all_sync(0xffffffffu, c == c0);
    if (warpUniform && c0 >= 0) { warp_minmax(lo0, hi0); warp_minmax(lo1, hi1); }
    if (blockUniform) {
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0)
            for (int k = 0; k < 3; ++k) {
                s_v[w][k] = lo0[k]; s_v[w][3 + k] = hi0[k]; s_v[w][6 + k] = lo1[k]; s_v[w][9 + k] = hi1[k];
            }
        __syncthreads();
        if (threadIdx.x < 12) {
            const int k = threadIdx.x, isMin = (k % 6) < 3;
            unsigned long long v = s_v[0][k];
            for (int q = 1; q < 8; ++q) {
                const unsigned long long u = s_v[q][k];
                v = isMin ? (u < v ? u : v) : (u > v ? u : v);
            }
            unsigned long long *dst = &nodes[left + k / 6].b[k % 6];
            if (isMin) { if (v != ~0ull) atomicMin(dst, v); }
            else if (v != 0ull) atomicMax(dst, v);
        }
        return;
    }
    if (warpUniform) {
        if (c0 >= 0 && (threadIdx.x & 31) == 0) {
            if (lo0[0] != ~0ull)
#pragma unroll
                for (int k = 0; k < 3; ++k) { atomicMin(&nodes[left].b[k], lo0[k]); atomicMax(&nodes[left].b[3 + k], hi0[k]); }
            if (lo1[0] != ~0ull)
#pragma unroll
                for (int k = 0; k < 3; ++k) { atomicMin(&nodes[left + 1].b[k], lo1[k]); atomicMax(&nodes[left + 1].b[3 + k], hi1[k]); }
        }
        return;
    }
    if (c >= 0) {
        BNode *ch = &nodes[left + (less ? 0 : 1)];
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&ch->b[k], key[k]); atomicMax(&ch->b[3 + k], key[k]); }
    }
}

__global__ void __launch_bounds__(256) k_swap(int n, int *cellOf, const int *levelStart, int level, const BNode *nodes,
                                              const int *tabL, const int *tabR, double *x, double *y, double *z, double *m,
                                              double *h, int *act, int *iord) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const int c = cellOf[j];
    if (c < levelStart[level]) return;
    const BNode &nd = nodes[c];
    if (nd.dim < 0) return;
    const int lo = nd.lo, mid = nd.mid;
    if (j - lo < nd.nMis) {
        const int a = tabL[j], b = tabR[j];
        double t;
        t = x[a]; x[a] = x[b]; x[b] = t;
        t = y[a]; y[a] = y[b]; y[b] = t;
        t = z[a]; z[a] = z[b]; z[b] = t;
        t = m[a]; m[a] = m[b]; m[b] = t;
        t = h[a]; h[a] = h[b]; h[b] = t;
        int u = iord[a]; iord[a] = iord[b]; iord[b] = u;
        if (act) { u = act[a]; act[a] = act[b]; act[b] = u; }
    }
    cellOf[j] = j < mid ? nd.left : nd.right;
}

// The sub-tree of one retired cell (<= GGB_WCAP particles), built by ONE WARP in shared memory with the same rules as the
// level-synchronous path: squeezed bounds as ordered integers, BuildBinary's decision, the reference's exchange
// partition reproduced from prefix counts (ballots), children allocated from the same node counter.  Depth-first with
// an explicit stack; the final cell numbering does not depend on allocation order (k_number).  At the end the warp
// applies its permutation to the particle payload and re-homes its positions (cellOf = the bucket holding them).
struct FinishSmem {
    double x[GGB_WCAP], y[GGB_WCAP], z[GGB_WCAP];
    double pm[GGB_WCAP], ph[GGB_WCAP];
    int pa[GGB_WCAP], po[GGB_WCAP];
    int stack[GGB_WCAP + 2][3]; // cell, first, one past last (offsets into the retired cell's range)
    unsigned char idx[GGB_WCAP], tabL[GGB_WCAP / 2], tabR[GGB_WCAP / 2];
};
#ifndef GGB_FINISH_WARPS
#define GGB_FINISH_WARPS 4
#endif

__global__ void __launch_bounds__(GGB_FINISH_WARPS * 32) k_finish(int nRet, const int *retired, BNode *nodes, int *cellOf,
                                                                  double *x, double *y, double *z, double *m, double *h,
                                                                  int *act, int *iord, int nBucket, int *ctr) {
    __shared__ FinishSmem s_all[GGB_FINISH_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * GGB_FINISH_WARPS + warp;
    if (w >= nRet) return;
    FinishSmem &S = s_all[warp];
    const unsigned lt = (1u << lane) - 1u;
    const int c0 = retired[w];
    const int lo = nodes[c0].lo, n = nodes[c0].hi - lo + 1;
    for (int i = lane; i < n; i += 32) {
        S.x[i] = x[lo + i]; S.y[i] = y[lo + i]; S.z[i] = z[lo + i];
        S.idx[i] = (unsigned char)i;
    }
    if (lane == 0) { S.stack[0][0] = c0; S.stack[0][1] = 0; S.stack[0][2] = n; }
    int sp = 1;
    __syncwarp();
    while (sp > 0) {
        --sp;
        const int cell = S.stack[sp][0], a = S.stack[sp][1], b = S.stack[sp][2];
        __syncwarp();
        // squeezed bounds
        unsigned long long mn[3] = {~0ull, ~0ull, ~0ull}, mx[3] = {0ull, 0ull, 0ull};
        for (int i = a + lane; i < b; i += 32) {
            const unsigned long long kx = enc(S.x[i]), ky = enc(S.y[i]), kz = enc(S.z[i]);
            mn[0] = kx < mn[0] ? kx : mn[0]; mx[0] = kx > mx[0] ? kx : mx[0];
            mn[1] = ky < mn[1] ? ky : mn[1]; mx[1] = ky > mx[1] ? ky : mx[1];
            mn[2] = kz < mn[2] ? kz : mn[2]; mx[2] = kz > mx[2] ? kz : mx[2];
        }
        warp_minmax(mn, mx);
        BNode nd;
        nd.lo = lo + a; nd.hi = lo + b - 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) { nd.b[k] = mn[k]; nd.b[3 + k] = mx[k]; }
        if (cell != c0 && lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { nodes[cell].b[k] = mn[k]; nodes[cell].b[3 + k] = mx[k]; }
        }
        double split;
        const int d = decide(nd, nBucket, &split);
        if (d < 0) { // bucket
            for (int i = a + lane; i < b; i += 32) cellOf[lo + i] = cell;
            if (lane == 0) nodes[cell].dim = -1;
            continue;
        }
        const double *cd = d == 0 ? S.x : (d == 1 ? S.y : S.z);
        // exchange partition (pkdUpperPart): totals first, then every misplaced element's pair index
        int total = 0;
        for (int base = a; base < b; base += 32) {
            const int i = base + lane;
            total += __popc(__ballot_sync(0xffffffffu, i < b && cd[i] < split));
        }
        const int mid = a + total;
        int before = 0, nMis = 0;
        for (int base = a; base < b; base += 32) {
            const int i = base + lane;
            const bool in = i < b, less = in && cd[i] < split;
            const unsigned bal = __ballot_sync(0xffffffffu, less);
            const int lb = before + __popc(bal & lt); // elements < split in [a, i)
            const bool misL = in && i < mid && !less;
            if (misL) S.tabL[(i - a) - lb] = (unsigned char)i;
            if (in && i >= mid && less) S.tabR[total - lb - 1] = (unsigned char)i;
            nMis += __popc(__ballot_sync(0xffffffffu, misL));
            before += __popc(bal);
        }
        __syncwarp();
        for (int k = lane; k < nMis; k += 32) {
            const int p = S.tabL[k], q = S.tabR[k];
            double t;
            t = S.x[p]; S.x[p] = S.x[q]; S.x[q] = t;
            t = S.y[p]; S.y[p] = S.y[q]; S.y[q] = t;
            t = S.z[p]; S.z[p] = S.z[q]; S.z[q] = t;
            const unsigned char u = S.idx[p]; S.idx[p] = S.idx[q]; S.idx[q] = u;
        }
        int id = 0;
        if (lane == 0) {
            id = atomicAdd(&ctr[0], 2);
            BNode ch;
            ch.left = ch.right = -1; ch.parent = cell; ch.dim = -1; ch.mid = 0; ch.nMis = 0; ch.split = 0.0;
            for (int k = 0; k < 3; ++k) { ch.b[k] = ~0ull; ch.b[3 + k] = 0ull; }
            ch.lo = lo + a; ch.hi = lo + mid - 1;
            nodes[id] = ch;
            ch.lo = lo + mid; ch.hi = lo + b - 1;
            nodes[id + 1] = ch;
            nodes[cell].left = id; nodes[cell].right = id + 1; nodes[cell].dim = d; nodes[cell].split = split;
            nodes[cell].mid = lo + mid; nodes[cell].nMis = nMis;
            S.stack[sp][0] = id + 1; S.stack[sp][1] = mid; S.stack[sp][2] = b;
            S.stack[sp + 1][0] = id; S.stack[sp + 1][1] = a; S.stack[sp + 1][2] = mid;
        }
        sp += 2;
        __syncwarp();
    }
    // the payload follows the permutation: staged through shared memory so that every read of the old order has
    // completed before the first write of the new one
    for (int i = lane; i < n; i += 32) {
        const int src = lo + S.idx[i];
        S.pm[i] = m[src]; S.ph[i] = h[src]; S.po[i] = iord[src];
        S.pa[i] = act ? act[src] : 0;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        x[lo + i] = S.x[i]; y[lo + i] = S.y[i]; z[lo + i] = S.z[i];
        m[lo + i] = S.pm[i]; h[lo + i] = S.ph[i]; iord[lo + i] = S.po[i];
        if (act) act[lo + i] = S.pa[i];
    }
}

// mass, centre of mass, mass-weighted softening and subtree size of every cell, children before parents
__global__ void __launch_bounds__(128) k_up(int nn, const BNode *nodes, const double *x, const double *y, const double *z,
                                            const double *m, const double *h, double *cmass, double *csoft, double *ccom,
                                            int *csize, int *arrive) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nn) return;
    if (nodes[t].left >= 0) return;
    {
        const int lo = nodes[t].lo, hi = nodes[t].hi;
        double M = 0.0, S = 0.0, rx = 0.0, ry = 0.0, rz = 0.0;
        for (int i = lo; i <= hi; ++i) {
            const double fm = m[i];
            M = __dadd_rn(M, fm);
            S = __dadd_rn(S, __dmul_rn(fm, h[i]));
            rx = __dadd_rn(rx, __dmul_rn(fm, x[i]));
            ry = __dadd_rn(ry, __dmul_rn(fm, y[i]));
            rz = __dadd_rn(rz, __dmul_rn(fm, z[i]));
        }
        if (M > 0) { S = __ddiv_rn(S, M); rx = __ddiv_rn(rx, M); ry = __ddiv_rn(ry, M); rz = __ddiv_rn(rz, M); }
        cmass[t] = M; csoft[t] = S;
        ccom[3 * (size_t)t] = rx; ccom[3 * (size_t)t + 1] = ry; ccom[3 * (size_t)t + 2] = rz;
        csize[t] = 1;
    }
    int node = t;
    for (;;) {
        const int p = nodes[node].parent;
        if (p < 0) break;
        __threadfence();
        if (atomicAdd(&arrive[p], 1) == 0) break;
        __threadfence();
        const int kids[2] = {nodes[p].left, nodes[p].right};
        double M = 0.0, S = 0.0, rx = 0.0, ry = 0.0, rz = 0.0;
        int sz = 1;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int q = kids[k];
            const double fm = __ldcg(&cmass[q]);
            M = __dadd_rn(M, fm);
            S = __dadd_rn(S, __dmul_rn(fm, __ldcg(&csoft[q])));
            rx = __dadd_rn(rx, __dmul_rn(fm, __ldcg(&ccom[3 * (size_t)q])));
            ry = __dadd_rn(ry, __dmul_rn(fm, __ldcg(&ccom[3 * (size_t)q + 1])));
            rz = __dadd_rn(rz, __dmul_rn(fm, __ldcg(&ccom[3 * (size_t)q + 2])));
            sz += __ldcg(&csize[q]);
        }
        if (M > 0) { S = __ddiv_rn(S, M); rx = __ddiv_rn(rx, M); ry = __ddiv_rn(ry, M); rz = __ddiv_rn(rz, M); }
        cmass[p] = M; csoft[p] = S;
        ccom[3 * (size_t)p] = rx; ccom[3 * (size_t)p + 1] = ry; ccom[3 * (size_t)p + 2] = rz;
        csize[p] = sz;
        node = p;
    }
}

// pre-order index (cell, lower subtree, upper subtree) and the threaded "next" cell of every cell
__global__ void __launch_bounds__(256) k_number(int nn, const BNode *nodes, const int *csize, int *pre, int *nextB) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nn) return;
    int idx = 0, cur = t, nxt = -2;
    for (;;) {
        const int p = nodes[cur].parent;
        if (p < 0) break;
        const int l = nodes[p].left;
        if (cur == l) {
            idx += 1;
            if (nxt == -2) nxt = nodes[p].right; // first ancestor (or self) that is a lower child: its sibling follows
        } else idx += 1 + csize[l];
        cur = p;
    }
    pre[t] = idx;
    nextB[t] = nxt == -2 ? -1 : nxt;
}

// Bmax of every cell: each particle climbs from its bucket to the root; lanes of a warp that stand on the same cell
// combine their distances first (adjacent particles share all but their deepest ancestors)
__global__ void __launch_bounds__(256) k_bmax(int n, const int *cellOf, const BNode *nodes, const double *x, const double *y,
                                              const double *z, const double *ccom, unsigned long long *bmaxBits) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    int cur = i < n ? cellOf[i] : -1;
    double px = 0, py = 0, pz = 0;
    if (i < n) { px = x[i]; py = y[i]; pz = z[i]; }
    while (__any_sync(0xffffffffu, cur >= 0)) {
        unsigned long long bits = 0ull;
        if (cur >= 0) {
            const double dx = __dsub_rn(px, ccom[3 * (size_t)cur]), dy = __dsub_rn(py, ccom[3 * (size_t)cur + 1]),
                         dz = __dsub_rn(pz, ccom[3 * (size_t)cur + 2]);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            bits = (unsigned long long)__double_as_longlong(__dsqrt_rn(d2));
        }
        const unsigned grp = __match_any_sync(0xffffffffu, cur);
        const unsigned hi = (unsigned)(bits >> 32), mh = __reduce_max_sync(grp, hi);
        const unsigned lo = hi == mh ? (unsigned)bits : 0u, ml = __reduce_max_sync(grp, lo);
        if (cur >= 0) {
            if ((int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31))
                atomicMax(&bmaxBits[cur], ((unsigned long long)mh << 32) | ml);
            cur = nodes[cur].parent;
        }
    }
}

__global__ void __launch_bounds__(256) k_emit(int nn, const BNode *nodes, const int *pre, const int *nextB, const double *cmass,
                                              const double *csoft, const double *ccom, const unsigned long long *bmaxBits,
                                              double c23, double dTheta, double *bnd, double *r, double *fMass,
                                              double *fSoft, double *fOpen2, int *pLower, int *pUpper, int *iLower,
                                              int *iUpper, int *iDim, double *fSplit, double *fBmax) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= nn) return;
    const BNode nd = nodes[t];
    const int g = pre[t];
#pragma unroll
    for (int k = 0; k < 6; ++k) bnd[6 * (size_t)g + k] = dec(nd.b[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) r[3 * (size_t)g + k] = ccom[3 * (size_t)t + k];
    fMass[g] = cmass[t];
    fSoft[g] = csoft[t];
    const double bmax = __longlong_as_double((long long)bmaxBits[t]);
    // OPEN_JOSH, pkd.c:2253-2260: c23 = 2/sqrt(3); the criteria that take "the minimal, i.e., Bmax" (pkd.c:2261-2264:
    // OPEN_RELPAR, OPEN_ABSTOT, OPEN_RELTOT) arrive as c23 = 0
    double dOpen = __ddiv_rn(__dmul_rn(c23, bmax), dTheta);
    if (dOpen < bmax) dOpen = bmax;
    fOpen2[g] = __dmul_rn(dOpen, dOpen);
    pLower[g] = nd.lo;
    pUpper[g] = nd.hi;
    iLower[g] = nd.left >= 0 ? pre[nd.left] : -1;
    iUpper[g] = nextB[t] >= 0 ? pre[nextB[t]] : -1;
    iDim[g] = nd.left >= 0 ? nd.dim : -1; // KDN.iDim / fSplit (pkd.h:454-456): -1 / 0 for a bucket
    fSplit[g] = nd.left >= 0 ? nd.split : 0.0;
    fBmax[g] = bmax;
}

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t need(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    ~Buf() { if (p) cudaFree(p); }
};

struct Builder {
    Buf part, ipart, nodes, cell, scan, tab, cub, ctr, up, num, out, outi, ret;
    int *hCtr = nullptr; // pinned
    ~Builder() { if (hCtr) cudaFreeHost(hCtr); }
};

} // namespace

#define BCK(call)                                                                                             \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            snprintf(err, errLen, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));       \
            return GG_ERR_CUDA;                                                                               \
        }                                                                                                     \
    } while (0)

int gg_builder_run(void **pBuilder, const gg_particles *pp, int nBucket, double dTheta, double c23, cudaStream_t st, GGBuiltDev *out,
                   int *pnLaunches, char *err, size_t errLen) {
    if (!*pBuilder) *pBuilder = new Builder();
    Builder &B = *(Builder *)*pBuilder;
    const int n = pp->n;
    const size_t maxNodes = 2 * (size_t)n + 2;
    BCK(B.part.need(sizeof(double) * 5 * (size_t)n));
    BCK(B.ipart.need(sizeof(int) * 2 * (size_t)n));
    BCK(B.nodes.need(sizeof(BNode) * maxNodes));
    BCK(B.cell.need(sizeof(int) * (size_t)n));
    BCK(B.scan.need(sizeof(int) * 2 * ((size_t)n + 1)));
    BCK(B.tab.need(sizeof(int) * 2 * (size_t)n));
    BCK(B.ret.need(sizeof(int) * ((size_t)n + 1)));
    BCK(B.ctr.need(sizeof(int) * (8 + GGB_MAX_LEVELS + 4)));
    if (!B.hCtr) BCK(cudaMallocHost((void **)&B.hCtr, sizeof(int) * 4));
    static_assert(GGB_WCAP <= 256 && GG_MAX_BUCKET <= GGB_WCAP, "k_finish indexes its particles with unsigned char");
    double *x = (double *)B.part.p, *y = x + n, *z = y + n, *m = z + n, *h = m + n;
    int *iord = (int *)B.ipart.p, *act = pp->active ? iord + n : nullptr;
    BNode *nodes = (BNode *)B.nodes.p;
    int *cellOf = (int *)B.cell.p, *flag = (int *)B.scan.p, *S = flag + n + 1;
    int *tabL = (int *)B.tab.p, *tabR = tabL + n;
    int *ctr = (int *)B.ctr.p, *levelStart = ctr + 8;
    size_t cubBytes = 0;
    BCK(cub::DeviceScan::ExclusiveSum(nullptr, cubBytes, flag, S, n + 1, st));
    BCK(B.cub.need(cubBytes));
    int nl = 0;
    // ---- particles in (input order)
    BCK(cudaMemcpyAsync(x, pp->x, sizeof(double) * n, cudaMemcpyDefault, st));
    BCK(cudaMemcpyAsync(y, pp->y, sizeof(double) * n, cudaMemcpyDefault, st));
    BCK(cudaMemcpyAsync(z, pp->z, sizeof(double) * n, cudaMemcpyDefault, st));
    BCK(cudaMemcpyAsync(m, pp->fMass, sizeof(double) * n, cudaMemcpyDefault, st));
    BCK(cudaMemcpyAsync(h, pp->fSoft, sizeof(double) * n, cudaMemcpyDefault, st));
    if (act) BCK(cudaMemcpyAsync(act, pp->active, sizeof(int) * n, cudaMemcpyDefault, st));
    const int gridP = (n + 255) / 256, gridP1 = (n + 1 + 255) / 256;
    k_b_init<<<gridP, 256, 0, st>>>(n, iord, cellOf, nodes, ctr, levelStart);
    ++nl;
    // ---- shape, level by level
    int nn = 1, level = 0;
    for (;; ++level) {
        if (level >= GGB_MAX_LEVELS) {
            snprintf(err, errLen, "gg_build_local: tree deeper than %d levels (coincident particles with nBucket=%d?)",
                     GGB_MAX_LEVELS, nBucket);
            return GG_ERR_UNSUPPORTED;
        }
        if (level == 0) { k_bounds<<<gridP, 256, 0, st>>>(n, x, y, z, cellOf, levelStart, level, nodes); ++nl; }
        k_flag<<<gridP1, 256, 0, st>>>(n, x, y, z, cellOf, levelStart, level, nodes, nBucket, flag, ctr, (int *)B.ret.p);
        BCK(cub::DeviceScan::ExclusiveSum(B.cub.p, cubBytes, flag, S, n + 1, st));
        k_split<<<gridP, 256, 0, st>>>(n, x, y, z, cellOf, levelStart, level, nodes, flag, S, tabL, tabR, ctr);
        k_swap<<<gridP, 256, 0, st>>>(n, cellOf, levelStart, level, nodes, tabL, tabR, x, y, z, m, h, act, iord);
        nl += 5;
        // while 2^level cells of GGB_WCAP particles cannot hold all n, some cell still splits here; afterwards look every level
        if ((1ll << (level + 1)) * (long long)GGB_WCAP < (long long)n) continue;
        BCK(cudaMemcpyAsync(B.hCtr, ctr, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        BCK(cudaMemcpyAsync(B.hCtr + 2, levelStart + level + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        nn = B.hCtr[0];
        if (B.hCtr[0] == B.hCtr[2]) break; // this level split nothing
    }
    BCK(cudaGetLastError());
    // ---- the sub-trees of the retired cells, one warp each
    const int nRet = B.hCtr[1];
    if (nRet > 0) {
        k_finish<<<(nRet + GGB_FINISH_WARPS - 1) / GGB_FINISH_WARPS, GGB_FINISH_WARPS * 32, 0, st>>>(
            nRet, (const int *)B.ret.p, nodes, cellOf, x, y, z, m, h, act, iord, nBucket, ctr);
        ++nl;
        BCK(cudaMemcpyAsync(B.hCtr, ctr, sizeof(int), cudaMemcpyDeviceToHost, st));
        BCK(cudaStreamSynchronize(st));
        nn = B.hCtr[0];
    }
    BCK(cudaGetLastError());
    // ---- per-cell quantities
    BCK(B.up.need(sizeof(double) * 6 * (size_t)nn + sizeof(int) * 2 * (size_t)nn));
    double *cmass = (double *)B.up.p, *csoft = cmass + nn, *ccom = csoft + nn;
    unsigned long long *bmaxBits = (unsigned long long *)(ccom + 3 * (size_t)nn);
    int *csize = (int *)(bmaxBits + nn), *arrive = csize + nn;
    BCK(B.num.need(sizeof(int) * 2 * (size_t)nn));
    int *pre = (int *)B.num.p, *nextB = pre + nn;
    BCK(B.out.need(sizeof(double) * 15 * (size_t)nn));
    BCK(B.outi.need(sizeof(int) * 5 * (size_t)nn));
    double *obnd = (double *)B.out.p, *orr = obnd + 6 * (size_t)nn, *oM = orr + 3 * (size_t)nn, *oS = oM + nn, *oO = oS + nn;
    int *oPL = (int *)B.outi.p, *oPU = oPL + nn, *oIL = oPU + nn, *oIU = oIL + nn, *oDim = oIU + nn;
    double *oSplit = oO + nn, *oBmax = oSplit + nn;
    BCK(cudaMemsetAsync(bmaxBits, 0, sizeof(unsigned long long) * nn + sizeof(int) * 2 * (size_t)nn, st));
    const int gridN = (nn + 255) / 256;
    k_up<<<(nn + 127) / 128, 128, 0, st>>>(nn, nodes, x, y, z, m, h, cmass, csoft, ccom, csize, arrive);
    k_number<<<gridN, 256, 0, st>>>(nn, nodes, csize, pre, nextB);
    k_bmax<<<gridP, 256, 0, st>>>(n, cellOf, nodes, x, y, z, ccom, bmaxBits);
    k_emit<<<gridN, 256, 0, st>>>(nn, nodes, pre, nextB, cmass, csoft, ccom, bmaxBits, c23, dTheta, obnd, orr, oM,
                                  oS, oO, oPL, oPU, oIL, oIU, oDim, oSplit, oBmax);
    nl += 4;
    BCK(cudaGetLastError());
    out->nNodes = nn; out->nPart = n; out->nLevels = level + 1;
    out->bnd = obnd; out->r = orr; out->fMass = oM; out->fSoft = oS; out->fOpen2 = oO;
    out->pLower = oPL; out->pUpper = oPU; out->iLower = oIL; out->iUpper = oIU;
    out->iDim = oDim; out->fSplit = oSplit; out->fBmax = oBmax;
    out->x = x; out->y = y; out->z = z; out->m = m; out->h = h; out->active = act; out->iorder = iord;
    if (pnLaunches) *pnLaunches = nl;
    return GG_OK;
}

void gg_builder_free(void *builder) { delete (Builder *)builder; }
