// gg_moments.cu -- the cells' multipole moments computed ON THE DEVICE (optional replacement for uploading
// pkd->kdNodes[].mom, 248 B per cell = 58 % of the bytes gg_set_local moves).
//
// pkdCalcCell (pkd.c:2018-2135) sums, for every cell, the reduced multipoles of ALL its particles about the cell's
// centre of mass: O(N log N) work per tree build.  Here each bucket forms the RAW Cartesian moments (orders 2..4, 31
// values) of its own particles, and every interior cell is the sum of its two children's raw moments translated to its
// own centre (exact binomial shift; the dipole about a centre of mass vanishes) -- O(cells).  The reduction to the
// reference's reduced multipoles (gg_m2m.h: gg_raw_reduce, the formulas of pkd.c:2056-2131) happens per cell at the end.
// All of it in FP64; the result differs from the host's particle-by-particle sums only by summation order (~1e-15 of
// M*Bmax^l), far below the FP32 the evaluation kernel rounds the moments to.  The centres r, fMass, fSoft and fOpen2 --
// everything the FP64 opening decisions read -- stay the host's, bit for bit, so the interaction lists do not change.
//
// Bottom-up without level lists: one thread per bucket computes its leaf, then climbs; at each parent the first child
// to arrive leaves, the second (after a fence) combines both children in the fixed order (c0, c1) -> deterministic.
#include "gg_internal.h"
#include "gg_m2m.h"

namespace {

__global__ void k_mom_parent(int nn, const NodeW *nodes, int nodeBase, int iRootLocal, int *parent, int *arrive) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const NodeW w = nodes[nodeBase + i];
    if (w.c0 >= 0) {
        parent[w.c0 - nodeBase] = i;
        parent[w.c1 - nodeBase] = i;
    }
    arrive[i] = 0;
    if (i == iRootLocal) parent[i] = -1;
}

__device__ __forceinline__ void store_raw(double *raw, const GGRawMom &r) {
    raw[0] = r.M;
#pragma unroll
    for (int k = 0; k < 31; ++k) raw[1 + k] = r.q[k];
}

__device__ __forceinline__ void load_raw(const double *raw, GGRawMom &r) { // L2 (another SM wrote it)
    r.M = __ldcg(raw);
#pragma unroll
    for (int k = 0; k < 31; ++k) r.q[k] = __ldcg(raw + 1 + k);
}

// the evaluation records of one finished cell: FP32 reduced moments with the quadrupole made traceless like SETILIST
// (walk.c:41-48), and the raw FP64 quadrupole of the softened-cell path -- the same records k_pack_mom writes
__device__ __forceinline__ void finish_cell(const GGRawMom &r, float4 *momf, double *momq) {
    double q[31];
    gg_raw_reduce(r, q);
    float f[32];
    gg_pack_momf(q, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) momf[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
#pragma unroll
    for (int k = 0; k < 6; ++k) momq[k] = q[k];
}

__global__ void __launch_bounds__(128) k_mom_up(int nn, const NodeW *nodes, int nodeBase, int partBase, const double *x,
                                                const double *y, const double *z, const double *m, const int *parent,
                                                int *arrive, double *raw, float4 *momf, double *momq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    NodeW w = nodes[nodeBase + i];
    if (w.c0 >= 0) return; // interior cells are formed by whichever child arrives second
    GGRawMom acc;
    gg_raw_zero(acc);
    for (int j = 0; j < w.nP; ++j) {
        const int pi = w.pLower - partBase + j;
        gg_raw_add_particle(acc, m[pi], x[pi] - w.rx, y[pi] - w.ry, z[pi] - w.rz);
    }
    store_raw(raw + (size_t)i * 32, acc);
    finish_cell(acc, momf + (size_t)(nodeBase + i) * 8, momq + (size_t)(nodeBase + i) * 6);
    int node = i;
    for (;;) {
        const int p = parent[node];
        if (p < 0) break;
        __threadfence(); // this cell's raw moments are visible before the arrival is
        if (atomicAdd(&arrive[p], 1) == 0) break;
        __threadfence();
        w = nodes[nodeBase + p];
        gg_raw_zero(acc);
        const int kids[2] = {w.c0, w.c1};
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            GGRawMom s;
            load_raw(raw + (size_t)(kids[k] - nodeBase) * 32, s);
            const double *cr = reinterpret_cast<const double *>(&nodes[kids[k]]); // rx, ry, rz lead the record
            gg_raw_shift_add(acc, s, cr[0] - w.rx, cr[1] - w.ry, cr[2] - w.rz);
        }
        store_raw(raw + (size_t)p * 32, acc);
        finish_cell(acc, momf + (size_t)(nodeBase + p) * 8, momq + (size_t)(nodeBase + p) * 6);
        node = p;
    }
}

// raw records -> the reference's reduced multipoles in FP64 (pkdCalcCell's definition), [nn][GG_NMOM]: for hosts that
// want kdNodes[].mom back (gg_tree_fetch)
__global__ void __launch_bounds__(128) k_mom_reduce(int nn, const double *raw, double *mom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    GGRawMom r;
    load_raw(raw + (size_t)i * 32, r);
    double q[31];
    gg_raw_reduce(r, q);
#pragma unroll
    for (int k = 0; k < 31; ++k) mom[(size_t)i * GG_NMOM + k] = q[k];
}

} // namespace

cudaError_t gg_launch_mom_reduce(int nn, const double *raw, double *mom, cudaStream_t st) {
    if (nn > 0) k_mom_reduce<<<(nn + 127) / 128, 128, 0, st>>>(nn, raw, mom);
    return cudaGetLastError();
}

cudaError_t gg_launch_device_moments(int nn, const NodeW *nodes, int nodeBase, int partBase, int iRootLocal,
                                     const double *x, const double *y, const double *z, const double *m, int *parent,
                                     int *arrive, double *raw, float4 *momf, double *momq, cudaStream_t st) {
    if (nn <= 0) return cudaSuccess;
    k_mom_parent<<<(nn + 255) / 256, 256, 0, st>>>(nn, nodes, nodeBase, iRootLocal, parent, arrive);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_mom_up<<<(nn + 127) / 128, 128, 0, st>>>(nn, nodes, nodeBase, partBase, x, y, z, m, parent, arrive, raw, momf, momq);
    return cudaGetLastError();
}
