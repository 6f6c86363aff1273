// gg_tree_kernel.cu -- fused tree walk + interaction-list evaluation, one warp per sink bucket (sm_100a).
//
// Replaces, per sink bucket, the reference's pkdBucketWalk (walk.c:306) + pkdBucketInteract (grav.c:23):
//
//  * WALK.  The warp keeps a frontier of (cell, periodic image) pairs in shared memory and tests 32 of them per
//    iteration, one per lane.  The test is the reference's, operation for operation and in FP64 without FMA
//    contraction (INTERSECTNP walk.h:12-30; "< 4 particles => open" walk.c:81; softened-cell classification
//    walk.c:118-127), so each bucket ends up with exactly the reference's three lists -- only their order differs,
//    which the reference's results do not depend on beyond rounding.  Opened cells push their two children, opened
//    buckets append their particles to a particle buffer, accepted cells go to a cell buffer.
//  * INTERACT.  Whenever a buffer holds 32 entries the warp evaluates them: lane j loads source j (a 128 B FP32
//    moment record, or a 32 B particle record) into registers once and loops over the bucket's <= 8 active sinks,
//    whose positions sit in shared memory relative to the bucket centre (the FP64 subtraction source-centre is
//    done at staging, so FP32 displacements keep ~1e-7 relative accuracy).  Per-sink accelerations, potentials
//    and max 1/dt^2 accumulate in registers and are reduced across the warp with shuffles once per bucket.
//    Lists never touch HBM.  1/r comes from MUFU.RSQ plus one Newton step.
//
// The arithmetic is FP32 CUDA-core work: ~130 FFMA-class instructions per (sink, hexadecapole cell)
// pair against the reference's score of 312 flops (grav.c:156-162), ~22 per (sink, particle) pair against 38.
#include "gg_internal.h"

#define FULL 0xffffffffu

namespace {

struct WarpSmem {
    float4 acc[GG_MAX_SINKS][32]; // per lane, per sink: ax, ay, az, potential
    float dtm[GG_MAX_SINKS][32];  // per lane, per sink: max 1/dt^2
    unsigned stack[GG_STACK_CAP];
    unsigned cbuf[64];
    float4 sink[GG_MAX_SINKS];    // x, y, z relative to the bucket centre, mass
    float sh[GG_MAX_SINKS];
    int sidx[GG_MAX_SINKS];
    unsigned pbuf[32]; // really 32*maxBucket + 32
};

__host__ __device__ inline size_t warp_smem_bytes(int maxBucket) {
    size_t b = sizeof(WarpSmem) + (size_t)32 * maxBucket * sizeof(unsigned);
    return (b + 15) & ~(size_t)15;
}

// INTERSECTNP (walk.h:12-30): squared distance from (x,y,z) to the box <= fBall2.  Intrinsics pin the rounding of
// every product and sum (no FMA), matching the reference's x86-64 build.
__device__ __forceinline__ bool intersect_np(const double *box, double fBall2, double x, double y, double z) {
    double dx = box[0] - x, dx1 = x - box[3];
    double dy = box[1] - y, dy1 = y - box[4];
    double dz = box[2] - z, dz1 = z - box[5];
    double d2;
    if (dx > 0.0) d2 = __dmul_rn(dx, dx);
    else if (dx1 > 0.0) d2 = __dmul_rn(dx1, dx1);
    else d2 = 0.0;
    if (dy > 0.0) d2 = __dadd_rn(d2, __dmul_rn(dy, dy));
    else if (dy1 > 0.0) d2 = __dadd_rn(d2, __dmul_rn(dy1, dy1));
    if (dz > 0.0) d2 = __dadd_rn(d2, __dmul_rn(dz, dz));
    else if (dz1 > 0.0) d2 = __dadd_rn(d2, __dmul_rn(dz1, dz1));
    return d2 <= fBall2;
}

__device__ __forceinline__ float rsqrt_nr(float d2) {
    float y = rsqrtf(d2);
    return y * fmaf(-0.5f * d2, y * y, 1.5f); // one Newton step: MUFU.RSQ is good to ~2^-22
}

__device__ __forceinline__ NodeW load_node(const NodeW *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    int4 d = __ldg(reinterpret_cast<const int4 *>(q + 3));
    NodeW n;
    n.rx = a.x; n.ry = a.y; n.rz = b.x; n.fOpen2 = b.y; n.fSoft = c.x; n.fMass = c.y;
    n.c0 = d.x; n.c1 = d.y; n.pLower = d.z; n.nP = d.w;
    return n;
}

__device__ __forceinline__ PartS load_part(const PartS *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = __ldg(q);
    float4 b = __ldg(reinterpret_cast<const float4 *>(q + 1)); // z (as 2 floats), m, h
    PartS s;
    s.x = a.x; s.y = a.y;
    s.z = __hiloint2double(__float_as_int(b.y), __float_as_int(b.x));
    s.m = b.z; s.h = b.w;
    return s;
}

// FP32 moment record of one cell held in registers (8 x float4 = the 128 B record):
//   m0 = Qxx Qyy Qzz Qxy | m1 = Qxz Qyz Oxxx Oxyy | m2 = Oxxy Oyyy Oxxz Oyyz | m3 = Oxyz Oxzz Oyzz Ozzz
//   m4 = Hxxxx Hxyyy Hxxxy Hyyyy | m5 = Hxxxz Hyyyz Hxxyy Hxxyz | m6 = Hxyyz Hxxzz Hxyzz Hxzzz
//   m7 = Hyyzz Hyzzz Hzzzz pad          (Q traceless; O, H reduced -- pkdCalcCell pkd.c:2093-2131)
struct CellMom {
    float4 m0, m1, m2, m3, m4, m5, m6, m7;
};

// Reduced-multipole evaluation of one Newtonian cell on one sink (QEVAL qeval.h:21-64 + gam[] grav.c:172-191),
// restructured around scaled monomials so every moment is used in exactly one FMA per force component.
// Returns the contribution (fx,fy,fz) to the acceleration, fp to -potential and 1/dt^2.
template <int ORDER>
__device__ __forceinline__ void cell_on_sink(const CellMom &c, float M, float dx, float dy, float dz, float ms,
                                             float &ox_, float &oy_, float &oz_, float &op_, float &odt) {
    const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    const float d2 = xx + yy + zz;
    const float g0 = rsqrt_nr(d2);
    const float dir2 = g0 * g0;
    const float g1 = g0 * dir2;
    float fx = 0.f, fy = 0.f, fz = 0.f, ta = g1 * M, fp = g0 * M;
    if (ORDER >= 2) {
        const float g2 = 3.f * g1 * dir2, g3 = 5.f * g2 * dir2;
        const float Qxx = c.m0.x, Qyy = c.m0.y, Qzz = c.m0.z, Qxy = c.m0.w, Qxz = c.m1.x, Qyz = c.m1.y;
        const float qx = fmaf(Qxz, dz, fmaf(Qxy, dy, Qxx * dx));
        const float qy = fmaf(Qyz, dz, fmaf(Qxy, dx, Qyy * dy));
        const float qz = fmaf(Qyz, dy, fmaf(Qxz, dx, Qzz * dz));
        const float qr = 0.5f * fmaf(qz, dz, fmaf(qy, dy, qx * dx));
        fp = fmaf(g2, qr, fp);
        ta = fmaf(g3, qr, ta);
        fx = g2 * qx; fy = g2 * qy; fz = g2 * qz;
        if (ORDER >= 3) {
            const float g4 = 7.f * g3 * dir2;
            const float hxx = 0.5f * xx, hyy = 0.5f * yy, hzz = 0.5f * zz;
            const float xy = dx * dy, xz = dx * dz, yz = dy * dz;
            const float Oxxx = c.m1.z, Oxyy = c.m1.w, Oxxy = c.m2.x, Oyyy = c.m2.y, Oxxz = c.m2.z, Oyyz = c.m2.w,
                        Oxyz = c.m3.x, Oxzz = c.m3.y, Oyzz = c.m3.z, Ozzz = c.m3.w;
            const float ox = fmaf(Oxzz, hzz, fmaf(Oxyz, yz, fmaf(Oxyy, hyy, fmaf(Oxxz, xz, fmaf(Oxxy, xy, Oxxx * hxx)))));
            const float oy = fmaf(Oyzz, hzz, fmaf(Oyyz, yz, fmaf(Oyyy, hyy, fmaf(Oxyz, xz, fmaf(Oxyy, xy, Oxxy * hxx)))));
            const float oz = fmaf(Ozzz, hzz, fmaf(Oyzz, yz, fmaf(Oyyz, hyy, fmaf(Oxzz, xz, fmaf(Oxyz, xy, Oxxz * hxx)))));
            const float orr = (1.f / 3.f) * fmaf(oz, dz, fmaf(oy, dy, ox * dx));
            fp = fmaf(g3, orr, fp);
            ta = fmaf(g4, orr, ta);
            fx = fmaf(g3, ox, fx); fy = fmaf(g3, oy, fy); fz = fmaf(g3, oz, fz);
            if (ORDER >= 4) {
                const float g5 = 9.f * g4 * dir2;
                // cubic monomials with multiplicity/6: x^3/6, x^2 y/2, xyz, ...
                const float cxxx = (1.f / 3.f) * hxx * dx, cyyy = (1.f / 3.f) * hyy * dy, czzz = (1.f / 3.f) * hzz * dz;
                const float cxxy = hxx * dy, cxxz = hxx * dz, cxyy = hyy * dx, cyyz = hyy * dz, cxzz = hzz * dx,
                            cyzz = hzz * dy, cxyz = xy * dz;
                const float Hxxxx = c.m4.x, Hxyyy = c.m4.y, Hxxxy = c.m4.z, Hyyyy = c.m4.w, Hxxxz = c.m5.x,
                            Hyyyz = c.m5.y, Hxxyy = c.m5.z, Hxxyz = c.m5.w, Hxyyz = c.m6.x, Hxxzz = c.m6.y,
                            Hxyzz = c.m6.z, Hxzzz = c.m6.w, Hyyzz = c.m7.x, Hyzzz = c.m7.y, Hzzzz = c.m7.z;
                const float hx = fmaf(Hxzzz, czzz, fmaf(Hxyzz, cyzz, fmaf(Hxyyz, cyyz, fmaf(Hxyyy, cyyy,
                                 fmaf(Hxxzz, cxzz, fmaf(Hxxyz, cxyz, fmaf(Hxxyy, cxyy, fmaf(Hxxxz, cxxz,
                                 fmaf(Hxxxy, cxxy, Hxxxx * cxxx)))))))));
                const float hy = fmaf(Hyzzz, czzz, fmaf(Hyyzz, cyzz, fmaf(Hyyyz, cyyz, fmaf(Hyyyy, cyyy,
                                 fmaf(Hxyzz, cxzz, fmaf(Hxyyz, cxyz, fmaf(Hxyyy, cxyy, fmaf(Hxxyz, cxxz,
                                 fmaf(Hxxyy, cxxy, Hxxxy * cxxx)))))))));
                const float hz = fmaf(Hzzzz, czzz, fmaf(Hyzzz, cyzz, fmaf(Hyyzz, cyyz, fmaf(Hyyyz, cyyy,
                                 fmaf(Hxzzz, cxzz, fmaf(Hxyzz, cxyz, fmaf(Hxyyz, cxyy, fmaf(Hxxzz, cxxz,
                                 fmaf(Hxxyz, cxxy, Hxxxz * cxxx)))))))));
                const float hr = 0.25f * fmaf(hz, dz, fmaf(hy, dy, hx * dx));
                fp = fmaf(g4, hr, fp);
                ta = fmaf(g5, hr, ta);
                fx = fmaf(g4, hx, fx); fy = fmaf(g4, hy, fy); fz = fmaf(g4, hz, fz);
            }
        }
    }
    op_ = fp;
    ox_ = fmaf(-dx, ta, fx);
    oy_ = fmaf(-dy, ta, fy);
    oz_ = fmaf(-dz, ta, fz);
    odt = (ms + M) * g1; // grav.c:189-190
}

// Particle-particle kernel with Hernquist-Katz K3 spline softening (SPLINEM grav.h:53-69, grav.c:89-108).
// Returns the contribution to the acceleration, to -potential, and 1/dt^2.
__device__ __forceinline__ void part_on_sink(float pm, float ph, float dx, float dy, float dz, float ms, float hs,
                                             float &ox_, float &oy_, float &oz_, float &op_, float &odt) {
    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float twoh = hs + ph;
    float a, b;
    if (d2 >= twoh * twoh) {
        a = rsqrt_nr(d2);
        b = a * a * a;
    } else {
        const float r = sqrtf(d2);
        const float dih = 2.0f / twoh;
        const float u = r * dih, u2 = u * u;
        const float dih3 = dih * dih * dih;
        if (u < 1.0f) {
            a = dih * (7.f / 5.f + u2 * (-2.f / 3.f + u2 * (3.f / 10.f - 0.1f * u)));
            b = dih3 * (4.f / 3.f + u2 * (-6.f / 5.f + 0.5f * u));
        } else {
            const float dir = 1.0f / r;
            a = fmaf(-1.f / 15.f, dir, dih * (8.f / 5.f + u2 * (-4.f / 3.f + u * (1.f + u * (-3.f / 10.f + u * (1.f / 30.f))))));
            b = fmaf(-1.f / 15.f, dir * dir * dir, dih3 * (8.f / 3.f + u * (-3.f + u * (6.f / 5.f - u * (1.f / 6.f)))));
        }
    }
    odt = (ms + pm) * b;
    b *= pm;
    op_ = a * pm;
    ox_ = -dx * b;
    oy_ = -dy * b;
    oz_ = -dz * b;
}

// Softened cell (ILCS) on one sink: SPLINEQ grav.h:17-50 + grav.c:126-150, FP64 (rare path).
__device__ __noinline__ void softcell_on_sink(double M, double hc, const double *Q, double dx, double dy, double dz,
                                              double ms, double hs, float4 *pacc, float *pdt) {
    double d2 = dx * dx + dy * dy + dz * dz;
    double dir = rsqrt(d2), twoh = hs + hc, a, b, c, d;
    if (d2 < twoh * twoh) {
        double dih = 2.0 / twoh, u = dih / dir;
        if (u < 1.0) {
            a = dih * (7.0 / 5.0 - 2.0 / 3.0 * u * u + 3.0 / 10.0 * u * u * u * u - 1.0 / 10.0 * u * u * u * u * u);
            b = dih * dih * dih * (4.0 / 3.0 - 6.0 / 5.0 * u * u + 1.0 / 2.0 * u * u * u);
            c = dih * dih * dih * dih * dih * (12.0 / 5.0 - 3.0 / 2.0 * u);
            d = 3.0 / 2.0 * dih * dih * dih * dih * dih * dih * dir;
        } else {
            a = -1.0 / 15.0 * dir + dih * (8.0 / 5.0 - 4.0 / 3.0 * u * u + u * u * u - 3.0 / 10.0 * u * u * u * u +
                                           1.0 / 30.0 * u * u * u * u * u);
            b = -1.0 / 15.0 * dir * dir * dir +
                dih * dih * dih * (8.0 / 3.0 - 3.0 * u + 6.0 / 5.0 * u * u - 1.0 / 6.0 * u * u * u);
            c = -1.0 / 5.0 * dir * dir * dir * dir * dir + 3.0 * dih * dih * dih * dih * dir +
                dih * dih * dih * dih * dih * (-12.0 / 5.0 + 1.0 / 2.0 * u);
            d = -dir * dir * dir * dir * dir * dir * dir + 3.0 * dih * dih * dih * dih * dir * dir * dir -
                1.0 / 2.0 * dih * dih * dih * dih * dih * dih * dir;
        }
    } else {
        a = dir; b = a * a * a; c = 3.0 * b * a * a; d = 5.0 * c * a * a;
    }
    double qirx = Q[0] * dx + Q[3] * dy + Q[4] * dz;
    double qiry = Q[3] * dx + Q[1] * dy + Q[5] * dz;
    double qirz = Q[4] * dx + Q[5] * dy + Q[2] * dz;
    double qir = 0.5 * (qirx * dx + qiry * dy + qirz * dz);
    double tr = 0.5 * (Q[0] + Q[1] + Q[2]);
    double qir3 = b * M + d * qir - c * tr;
    float4 v = *pacc;
    v.w -= (float)(a * M + c * qir - b * tr);
    v.x -= (float)(qir3 * dx - c * qirx);
    v.y -= (float)(qir3 * dy - c * qiry);
    v.z -= (float)(qir3 * dz - c * qirz);
    *pacc = v;
    *pdt = fmaxf(*pdt, (float)((ms + M) * b));
}

template <int ORDER>
__global__ void __launch_bounds__(GG_WARPS_PER_CTA * 32) k_tree_gravity(const TreeKernelArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_off[GG_MAX_IMAGES * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < A.nImages * 3; i += blockDim.x) s_off[i] = A.imgOff[i];
    __syncthreads();
    WarpSmem &W = *reinterpret_cast<WarpSmem *>(smem_raw + warp * warp_smem_bytes(A.maxBucket));
    const unsigned imgMask = (1u << A.imgBits) - 1u;

    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(A.taskCounter, 1);
        t = __shfl_sync(FULL, t, 0);
        if (t >= A.nTasks) break;
        const Task task = A.tasks[t];
        const NodeW bk = load_node(&A.nodes[task.node]);

        // ---- stage the sinks: bbox of ACTIVE particles (pkd.c:2916-2932), fSoftMax over ALL (walk.c:319-324)
        double box[6] = {1.7976931348623157e308,  1.7976931348623157e308,  1.7976931348623157e308,
                         -1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
        double fSoftMax = 0.0;
        int nAct = 0;
        for (int base = 0; base < bk.nP; base += 32) {
            const int j = base + lane;
            bool act = false;
            if (j < bk.nP) {
                const int pi = bk.pLower + j;
                act = A.active ? (A.active[pi] != 0) : true;
                fSoftMax = fmax(fSoftMax, A.hsoft[pi]);
                if (act) {
                    const PartS p = load_part(&A.parts[pi]);
                    box[0] = fmin(box[0], p.x); box[3] = fmax(box[3], p.x);
                    box[1] = fmin(box[1], p.y); box[4] = fmax(box[4], p.y);
                    box[2] = fmin(box[2], p.z); box[5] = fmax(box[5], p.z);
                }
            }
            nAct += __popc(__ballot_sync(FULL, act));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                box[k] = fmin(box[k], __shfl_xor_sync(FULL, box[k], o));
                box[3 + k] = fmax(box[3 + k], __shfl_xor_sync(FULL, box[3 + k], o));
            }
            fSoftMax = fmax(fSoftMax, __shfl_xor_sync(FULL, fSoftMax, o));
        }
        const double cenx = 0.5 * (box[0] + box[3]), ceny = 0.5 * (box[1] + box[4]), cenz = 0.5 * (box[2] + box[5]);
        // sinks of this pass: active ranks [8*group, 8*group+8)
        const int rank0 = task.group * GG_MAX_SINKS;
        const int nS = min(GG_MAX_SINKS, nAct - rank0);
        __syncwarp();
        {
            int seen = 0;
            for (int base = 0; base < bk.nP; base += 32) {
                const int j = base + lane, pi = bk.pLower + j;
                bool act = false;
                if (j < bk.nP) act = A.active ? (A.active[pi] != 0) : true;
                const unsigned m = __ballot_sync(FULL, act);
                const int rank = seen + __popc(m & lt) - rank0;
                if (act && rank >= 0 && rank < GG_MAX_SINKS) {
                    const PartS p = load_part(&A.parts[pi]);
                    W.sink[rank] = make_float4((float)(p.x - cenx), (float)(p.y - ceny), (float)(p.z - cenz), p.m);
                    W.sh[rank] = p.h;
                    W.sidx[rank] = pi;
                }
                seen += __popc(m);
            }
        }
        for (int s = 0; s < nS; ++s) {
            W.acc[s][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            W.dtm[s][lane] = 0.f;
        }

        // ---- walk + evaluate
        int nStack = A.nImages, nCell = 0, nPart = 0;
        int cntP = 0, cntS = 0, cntN = 0;
        for (int i = lane; i < A.nImages; i += 32) W.stack[i] = ((unsigned)A.rootNode << A.imgBits) | (unsigned)i;
        __syncwarp();

        for (;;) {
            if (nStack > 0) {
                int k = min(32, nStack);
                if (nStack > GG_STACK_CAP - GG_STACK_DFS_MARGIN) k = 1; // near the cap: depth-first, growth <= 1 per step
                const bool has = lane < k;
                const unsigned item = has ? W.stack[nStack - 1 - lane] : 0xffffffffu;
                nStack -= k;
                __syncwarp();
                int action = 0; // 1 push children, 2 Newtonian cell, 3 source bucket, 4 own bucket, 5 softened cell
                NodeW nd;
                int node = 0, img = 0;
                nd.nP = 0; nd.c0 = -1; nd.c1 = -1; nd.pLower = 0; nd.fMass = 0; nd.fSoft = 0;
                double x = 0, y = 0, z = 0;
                if (item != 0xffffffffu) {
                    node = (int)(item >> A.imgBits);
                    img = (int)(item & imgMask);
                    nd = load_node(&A.nodes[node]);
                    x = nd.rx + s_off[3 * img];
                    y = nd.ry + s_off[3 * img + 1];
                    z = nd.rz + s_off[3 * img + 2];
                    bool open = intersect_np(box, nd.fOpen2, x, y, z);
                    if (nd.nP < 4) open = true; // walk.c:81 (pUpper - pLower < 3)
                    if (open) {
                        if (nd.c0 >= 0) action = 1;
                        else action = (node == task.node && img == A.homeImage) ? 4 : 3; // walk.c:93
                    } else {
                        double twoh2 = nd.fSoft + fSoftMax;
                        twoh2 = __dmul_rn(twoh2, twoh2);
                        bool soft = false;
                        if (!(twoh2 < nd.fOpen2)) soft = intersect_np(box, twoh2, x, y, z); // walk.c:122-127
                        action = soft ? 5 : 2;
                    }
                }
                // children (a single-child cell, pkdThreadTree pkd.c:2597-2609, pushes a no-op as second item)
                const unsigned mPush = __ballot_sync(FULL, action == 1);
                if (action == 1) {
                    const int pos = nStack + 2 * __popc(mPush & lt);
                    if (pos + 1 < GG_STACK_CAP) {
                        W.stack[pos] = nd.c1 >= 0 ? (((unsigned)nd.c1 << A.imgBits) | (unsigned)img) : 0xffffffffu;
                        W.stack[pos + 1] = ((unsigned)nd.c0 << A.imgBits) | (unsigned)img;
                    } else atomicExch(A.errFlag, 1);
                }
                nStack += 2 * __popc(mPush);
                // Newtonian cells
                const unsigned mCell = __ballot_sync(FULL, action == 2);
                if (action == 2) W.cbuf[nCell + __popc(mCell & lt)] = item;
                nCell += __popc(mCell);
                cntN += __popc(mCell);
                // source particles
                const int np = (action == 3 || action == 4) ? nd.nP : 0;
                const unsigned mBk = __ballot_sync(FULL, np > 0);
                if (mBk) {
                    int incl = np;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const int total = __shfl_sync(FULL, incl, 31);
                    const int wpos = nPart + incl - np;
                    for (int j = 0; j < np; ++j)
                        W.pbuf[wpos + j] = ((unsigned)(nd.pLower + j) << A.imgBits) | (unsigned)img;
                    const unsigned mOwn = __ballot_sync(FULL, action == 4);
                    const int own = mOwn ? __shfl_sync(FULL, np, __ffs(mOwn) - 1) : 0;
                    nPart += total;
                    cntP += total - own;
                }
                // softened cells: evaluated on the spot by the lane that found them
                const unsigned mSoft = __ballot_sync(FULL, action == 5);
                if (mSoft) {
                    cntS += __popc(mSoft);
                    if (action == 5 && !A.walkOnly) {
                        const double *Q = &A.momq[(size_t)node * 6];
                        for (int s = 0; s < nS; ++s) {
                            const float4 sk = W.sink[s];
                            softcell_on_sink(nd.fMass, nd.fSoft, Q, ((double)sk.x + cenx) - x, ((double)sk.y + ceny) - y,
                                             ((double)sk.z + cenz) - z, (double)sk.w, (double)W.sh[s], &W.acc[s][lane],
                                             &W.dtm[s][lane]);
                        }
                    }
                }
                __syncwarp();
            }
            const bool done = (nStack == 0);
            if (A.walkOnly) {
                nCell = 0;
                nPart = 0;
            }
            // ---- evaluate full chunks (from the top of each buffer: no shifting), everything once the walk is done
            while (nCell >= 32 || (done && nCell > 0)) {
                const int cnt = min(32, nCell);
                nCell -= cnt;
                if (lane < cnt) {
                    const unsigned item = W.cbuf[nCell + lane];
                    const int node = (int)(item >> A.imgBits), img = (int)(item & imgMask);
                    const double2 *nq = reinterpret_cast<const double2 *>(&A.nodes[node]);
                    const double2 p01 = __ldg(nq), p23 = __ldg(nq + 1), p45 = __ldg(nq + 2);
                    const float cx = (float)((p01.x + s_off[3 * img]) - cenx);
                    const float cy = (float)((p01.y + s_off[3 * img + 1]) - ceny);
                    const float cz = (float)((p23.x + s_off[3 * img + 2]) - cenz);
                    const float M = (float)p45.y;
                    CellMom c;
                    const float4 *mq = &A.momf[(size_t)node * 8];
                    c.m0 = __ldg(mq); c.m1 = __ldg(mq + 1);
                    if (ORDER >= 3) { c.m2 = __ldg(mq + 2); c.m3 = __ldg(mq + 3); }
                    if (ORDER >= 4) { c.m4 = __ldg(mq + 4); c.m5 = __ldg(mq + 5); c.m6 = __ldg(mq + 6); c.m7 = __ldg(mq + 7); }
#pragma unroll 2
                    for (int s = 0; s < nS; ++s) {
                        const float4 sk = W.sink[s];
                        float fx, fy, fz, fp, fdt;
                        cell_on_sink<ORDER>(c, M, sk.x - cx, sk.y - cy, sk.z - cz, sk.w, fx, fy, fz, fp, fdt);
                        float4 v = W.acc[s][lane];
                        v.x += fx; v.y += fy; v.z += fz; v.w -= fp;
                        W.acc[s][lane] = v;
                        W.dtm[s][lane] = fmaxf(W.dtm[s][lane], fdt);
                    }
                }
                __syncwarp();
            }
            while (nPart >= 32 || (done && nPart > 0)) {
                const int cnt = min(32, nPart);
                nPart -= cnt;
                if (lane < cnt) {
                    const unsigned item = W.pbuf[nPart + lane];
                    const int pi = (int)(item >> A.imgBits), img = (int)(item & imgMask);
                    const PartS p = load_part(&A.parts[pi]);
                    const float px = (float)((p.x + s_off[3 * img]) - cenx);
                    const float py = (float)((p.y + s_off[3 * img + 1]) - ceny);
                    const float pz = (float)((p.z + s_off[3 * img + 2]) - cenz);
                    const bool home = (img == A.homeImage);
#pragma unroll 2
                    for (int s = 0; s < nS; ++s) {
                        if (home && pi == W.sidx[s]) continue; // a particle does not act on itself (grav.c:211)
                        const float4 sk = W.sink[s];
                        float fx, fy, fz, fp, fdt;
                        part_on_sink(p.m, p.h, sk.x - px, sk.y - py, sk.z - pz, sk.w, W.sh[s], fx, fy, fz, fp, fdt);
                        float4 v = W.acc[s][lane];
                        v.x += fx; v.y += fy; v.z += fz; v.w -= fp;
                        W.acc[s][lane] = v;
                        W.dtm[s][lane] = fmaxf(W.dtm[s][lane], fdt);
                    }
                }
                __syncwarp();
            }
            if (done) break;
        }

        // ---- reduce across the warp and write out
        if (task.group == 0 && lane == 0) {
            A.counts[3 * task.node] = cntP;
            A.counts[3 * task.node + 1] = cntS;
            A.counts[3 * task.node + 2] = cntN;
        }
        if (!A.walkOnly) {
            for (int s = 0; s < nS; ++s) {
                const float4 v = W.acc[s][lane];
                double vx = v.x, vy = v.y, vz = v.z, vp = v.w;
                float vd = W.dtm[s][lane];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    vx += __shfl_xor_sync(FULL, vx, o);
                    vy += __shfl_xor_sync(FULL, vy, o);
                    vz += __shfl_xor_sync(FULL, vz, o);
                    vp += __shfl_xor_sync(FULL, vp, o);
                    vd = fmaxf(vd, __shfl_xor_sync(FULL, vd, o));
                }
                if (lane == 0) {
                    const int pi = W.sidx[s];
                    A.acc[3 * (size_t)pi] = vx;
                    A.acc[3 * (size_t)pi + 1] = vy;
                    A.acc[3 * (size_t)pi + 2] = vz;
                    A.pot[pi] = vp;
                    A.dtg[pi] = (double)vd;
                }
            }
        }
        __syncwarp();
    }
}

} // namespace

size_t gg_tree_kernel_smem(int maxBucket) { return GG_WARPS_PER_CTA * warp_smem_bytes(maxBucket); }

cudaError_t gg_launch_tree_kernel(const TreeKernelArgs &a, int nSM, cudaStream_t st) {
    void (*fn)(const TreeKernelArgs) = nullptr;
    switch (a.iOrder) {
    case 1: fn = k_tree_gravity<1>; break;
    case 2: fn = k_tree_gravity<2>; break;
    case 3: fn = k_tree_gravity<3>; break;
    default: fn = k_tree_gravity<4>; break;
    }
    size_t smem = gg_tree_kernel_smem(a.maxBucket);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int perSM = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, fn, GG_WARPS_PER_CTA * 32, smem);
    if (e != cudaSuccess) return e;
    if (perSM < 1) perSM = 1;
    int grid = nSM * perSM;
    int need = (a.nTasks + GG_WARPS_PER_CTA - 1) / GG_WARPS_PER_CTA;
    if (grid > need) grid = need > 0 ? need : 1;
    fn<<<grid, GG_WARPS_PER_CTA * 32, smem, st>>>(a);
    return cudaGetLastError();
}
