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
    unsigned stack[GG_STACK_CAP];
    unsigned cbuf[64];
    float sx[GG_MAX_SINKS], sy[GG_MAX_SINKS], sz[GG_MAX_SINKS], sh[GG_MAX_SINKS], sm[GG_MAX_SINKS];
    int sidx[GG_MAX_SINKS];
    double box[6];
    double cen[3];
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

// Per-sink accumulators of one lane.
struct Acc {
    float ax[GG_MAX_SINKS], ay[GG_MAX_SINKS], az[GG_MAX_SINKS], po[GG_MAX_SINKS], dt[GG_MAX_SINKS];
};

// Reduced-multipole evaluation of one Newtonian cell on one sink (QEVAL qeval.h:21-64 + gam[] grav.c:172-191),
// restructured around scaled monomials so every moment is used in exactly one FMA per force component.
// q[] = traceless Q (xx,yy,zz,xy,xz,yz), O (xxx,xyy,xxy,yyy,xxz,yyz,xyz,xzz,yzz,zzz),
//       H (xxxx,xyyy,xxxy,yyyy,xxxz,yyyz,xxyy,xxyz,xyyz,xxzz,xyzz,xzzz,yyzz,yzzz,zzzz).
template <int ORDER>
__device__ __forceinline__ void cell_on_sink(const float (&q)[32], float M, float dx, float dy, float dz, float ms,
                                             float &ax, float &ay, float &az, float &po, float &dtmax) {
    float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float g0 = rsqrt_nr(d2);
    float dir2 = g0 * g0;
    float g1 = g0 * dir2;
    float fx = 0.f, fy = 0.f, fz = 0.f, ta = g1 * M, fp = g0 * M;
    if (ORDER >= 2) {
        float g2 = 3.f * g1 * dir2, g3 = 5.f * g2 * dir2;
        float qx = fmaf(q[4], dz, fmaf(q[3], dy, q[0] * dx));
        float qy = fmaf(q[5], dz, fmaf(q[3], dx, q[1] * dy));
        float qz = fmaf(q[5], dy, fmaf(q[4], dx, q[2] * dz));
        float qr = 0.5f * fmaf(qz, dz, fmaf(qy, dy, qx * dx));
        fp = fmaf(g2, qr, fp);
        ta = fmaf(g3, qr, ta);
        fx = g2 * qx; fy = g2 * qy; fz = g2 * qz;
        if (ORDER >= 3) {
            float g4 = 7.f * g3 * dir2;
            float hxx = 0.5f * dx * dx, hyy = 0.5f * dy * dy, hzz = 0.5f * dz * dz;
            float xy = dx * dy, xz = dx * dz, yz = dy * dz;
            float ox = fmaf(q[13], hzz, fmaf(q[12], yz, fmaf(q[7], hyy, fmaf(q[10], xz, fmaf(q[8], xy, q[6] * hxx)))));
            float oy = fmaf(q[14], hzz, fmaf(q[11], yz, fmaf(q[9], hyy, fmaf(q[12], xz, fmaf(q[7], xy, q[8] * hxx)))));
            float oz = fmaf(q[15], hzz, fmaf(q[14], yz, fmaf(q[11], hyy, fmaf(q[13], xz, fmaf(q[12], xy, q[10] * hxx)))));
            float orr = (1.f / 3.f) * fmaf(oz, dz, fmaf(oy, dy, ox * dx));
            fp = fmaf(g3, orr, fp);
            ta = fmaf(g4, orr, ta);
            fx = fmaf(g3, ox, fx); fy = fmaf(g3, oy, fy); fz = fmaf(g3, oz, fz);
            if (ORDER >= 4) {
                float g5 = 9.f * g4 * dir2;
                // cubic monomials with multiplicity/6: x^3/6, x^2 y/2, xyz, ...
                float cxxx = (1.f / 3.f) * hxx * dx, cyyy = (1.f / 3.f) * hyy * dy, czzz = (1.f / 3.f) * hzz * dz;
                float cxxy = hxx * dy, cxxz = hxx * dz, cxyy = hyy * dx, cyyz = hyy * dz, cxzz = hzz * dx,
                      cyzz = hzz * dy, cxyz = xy * dz;
                const float *H = &q[16];
                float hx = fmaf(H[11], czzz, fmaf(H[10], cyzz, fmaf(H[8], cyyz, fmaf(H[1], cyyy,
                           fmaf(H[9], cxzz, fmaf(H[7], cxyz, fmaf(H[6], cxyy, fmaf(H[4], cxxz,
                           fmaf(H[2], cxxy, H[0] * cxxx)))))))));
                float hy = fmaf(H[13], czzz, fmaf(H[12], cyzz, fmaf(H[5], cyyz, fmaf(H[3], cyyy,
                           fmaf(H[10], cxzz, fmaf(H[8], cxyz, fmaf(H[1], cxyy, fmaf(H[7], cxxz,
                           fmaf(H[6], cxxy, H[2] * cxxx)))))))));
                float hz = fmaf(H[14], czzz, fmaf(H[13], cyzz, fmaf(H[12], cyyz, fmaf(H[5], cyyy,
                           fmaf(H[11], cxzz, fmaf(H[10], cxyz, fmaf(H[8], cxyy, fmaf(H[9], cxxz,
                           fmaf(H[7], cxxy, H[4] * cxxx)))))))));
                float hr = 0.25f * fmaf(hz, dz, fmaf(hy, dy, hx * dx));
                fp = fmaf(g4, hr, fp);
                ta = fmaf(g5, hr, ta);
                fx = fmaf(g4, hx, fx); fy = fmaf(g4, hy, fy); fz = fmaf(g4, hz, fz);
            }
        }
    }
    po -= fp;
    ax += fmaf(-dx, ta, fx);
    ay += fmaf(-dy, ta, fy);
    az += fmaf(-dz, ta, fz);
    dtmax = fmaxf(dtmax, (ms + M) * g1); // grav.c:189-190
}

// Particle-particle kernel with Hernquist-Katz K3 spline softening (SPLINEM grav.h:53-69, grav.c:89-108).
__device__ __forceinline__ void part_on_sink(float pm, float ph, float dx, float dy, float dz, float ms, float hs,
                                             float &ax, float &ay, float &az, float &po, float &dtmax) {
    float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float twoh = hs + ph;
    float a, b;
    if (d2 >= twoh * twoh) {
        a = rsqrt_nr(d2);
        b = a * a * a;
    } else {
        float r = sqrtf(d2);
        float dih = 2.0f / twoh;
        float u = r * dih, u2 = u * u;
        float dih3 = dih * dih * dih;
        if (u < 1.0f) {
            a = dih * (7.f / 5.f + u2 * (-2.f / 3.f + u2 * (3.f / 10.f - 0.1f * u)));
            b = dih3 * (4.f / 3.f + u2 * (-6.f / 5.f + 0.5f * u));
        } else {
            float dir = 1.0f / r;
            a = fmaf(-1.f / 15.f, dir, dih * (8.f / 5.f + u2 * (-4.f / 3.f + u * (1.f + u * (-3.f / 10.f + u * (1.f / 30.f))))));
            b = fmaf(-1.f / 15.f, dir * dir * dir, dih3 * (8.f / 3.f + u * (-3.f + u * (6.f / 5.f - u * (1.f / 6.f)))));
        }
    }
    dtmax = fmaxf(dtmax, (ms + pm) * b);
    a *= pm;
    b *= pm;
    po -= a;
    ax = fmaf(-dx, b, ax);
    ay = fmaf(-dy, b, ay);
    az = fmaf(-dz, b, az);
}

// Softened cell (ILCS) on one sink: SPLINEQ grav.h:17-50 + grav.c:126-150, FP64 (rare path).
__device__ __noinline__ void softcell_on_sink(double M, double hc, const double *Q, double dx, double dy, double dz,
                                              double ms, double hs, float &ax, float &ay, float &az, float &po,
                                              float &dtmax) {
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
    po -= (float)(a * M + c * qir - b * tr);
    ax -= (float)(qir3 * dx - c * qirx);
    ay -= (float)(qir3 * dy - c * qiry);
    az -= (float)(qir3 * dz - c * qirz);
    dtmax = fmaxf(dtmax, (float)((ms + M) * b));
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
    const int pbufCap = 32 * A.maxBucket + 32;

    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(A.taskCounter, 1);
        t = __shfl_sync(FULL, t, 0);
        if (t >= A.nTasks) break;
        const Task task = A.tasks[t];
        const NodeW bk = load_node(&A.nodes[task.node]);

        // ---- stage the sinks: bbox of ACTIVE particles (pkd.c:2916-2932), fSoftMax over ALL (walk.c:319-324)
        double mn[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
        double mx[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
        double fSoftMax = 0.0;
        int nAct = 0;
        for (int base = 0; base < bk.nP; base += 32) {
            int j = base + lane;
            bool in = j < bk.nP;
            bool act = false;
            if (in) {
                int pi = bk.pLower + j;
                act = A.active ? (A.active[pi] != 0) : true;
                double h = A.hsoft[pi];
                if (h > fSoftMax) fSoftMax = h;
                if (act) {
                    PartS p = load_part(&A.parts[pi]);
                    mn[0] = fmin(mn[0], p.x); mx[0] = fmax(mx[0], p.x);
                    mn[1] = fmin(mn[1], p.y); mx[1] = fmax(mx[1], p.y);
                    mn[2] = fmin(mn[2], p.z); mx[2] = fmax(mx[2], p.z);
                }
            }
            nAct += __popc(__ballot_sync(FULL, act));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                mn[k] = fmin(mn[k], __shfl_xor_sync(FULL, mn[k], o));
                mx[k] = fmax(mx[k], __shfl_xor_sync(FULL, mx[k], o));
            }
            fSoftMax = fmax(fSoftMax, __shfl_xor_sync(FULL, fSoftMax, o));
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                W.box[k] = mn[k];
                W.box[3 + k] = mx[k];
                W.cen[k] = 0.5 * (mn[k] + mx[k]);
            }
        }
        __syncwarp();
        const double cenx = W.cen[0], ceny = W.cen[1], cenz = W.cen[2];
        // sinks of this pass: active ranks [8*group, 8*group+8)
        const int rank0 = task.group * GG_MAX_SINKS;
        const int nS = min(GG_MAX_SINKS, nAct - rank0);
        {
            int seen = 0;
            for (int base = 0; base < bk.nP; base += 32) {
                int j = base + lane;
                bool act = false;
                int pi = bk.pLower + j;
                if (j < bk.nP) act = A.active ? (A.active[pi] != 0) : true;
                unsigned m = __ballot_sync(FULL, act);
                int rank = seen + __popc(m & lt) - rank0;
                if (act && rank >= 0 && rank < GG_MAX_SINKS) {
                    PartS p = load_part(&A.parts[pi]);
                    W.sx[rank] = (float)(p.x - cenx);
                    W.sy[rank] = (float)(p.y - ceny);
                    W.sz[rank] = (float)(p.z - cenz);
                    W.sh[rank] = p.h;
                    W.sm[rank] = p.m;
                    W.sidx[rank] = pi;
                }
                seen += __popc(m);
            }
        }
        Acc acc;
#pragma unroll
        for (int s = 0; s < GG_MAX_SINKS; ++s) acc.ax[s] = acc.ay[s] = acc.az[s] = acc.po[s] = acc.dt[s] = 0.f;

        // ---- walk
        int nStack = A.nImages, nCell = 0, nPart = 0;
        int cntP = 0, cntS = 0, cntN = 0;
        for (int i = lane; i < A.nImages; i += 32) W.stack[i] = ((unsigned)A.rootNode << A.imgBits) | (unsigned)i;
        __syncwarp();

        auto eval_cells = [&](int first, int count) {
            const bool valid = lane < count;
            if (valid) {
                unsigned item = W.cbuf[first + lane];
                int node = (int)(item >> A.imgBits), img = (int)(item & imgMask);
                const double2 *nq = reinterpret_cast<const double2 *>(&A.nodes[node]);
                double2 p01 = __ldg(nq), p23 = __ldg(nq + 1), p45 = __ldg(nq + 2);
                const float cx = (float)((p01.x + s_off[3 * img]) - cenx);
                const float cy = (float)((p01.y + s_off[3 * img + 1]) - ceny);
                const float cz = (float)((p23.x + s_off[3 * img + 2]) - cenz);
                const float M = (float)p45.y;
                float q[32];
                const float4 *mq = &A.momf[(size_t)node * 8];
#pragma unroll
                for (int k = 0; k < (ORDER >= 4 ? 8 : (ORDER == 3 ? 4 : 2)); ++k) {
                    float4 v = __ldg(mq + k);
                    q[4 * k] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w;
                }
#pragma unroll
                for (int s = 0; s < GG_MAX_SINKS; ++s)
                    if (s < nS)
                        cell_on_sink<ORDER>(q, M, W.sx[s] - cx, W.sy[s] - cy, W.sz[s] - cz, W.sm[s], acc.ax[s],
                                            acc.ay[s], acc.az[s], acc.po[s], acc.dt[s]);
            }
        };
        auto eval_parts = [&](int first, int count) {
            const bool valid = lane < count;
            if (valid) {
                unsigned item = W.pbuf[first + lane];
                int pi = (int)(item >> A.imgBits), img = (int)(item & imgMask);
                PartS p = load_part(&A.parts[pi]);
                const float px = (float)((p.x + s_off[3 * img]) - cenx);
                const float py = (float)((p.y + s_off[3 * img + 1]) - ceny);
                const float pz = (float)((p.z + s_off[3 * img + 2]) - cenz);
                const bool home = (img == A.homeImage);
#pragma unroll
                for (int s = 0; s < GG_MAX_SINKS; ++s)
                    if (s < nS && !(home && pi == W.sidx[s]))
                        part_on_sink(p.m, p.h, W.sx[s] - px, W.sy[s] - py, W.sz[s] - pz, W.sm[s], W.sh[s],
                                     acc.ax[s], acc.ay[s], acc.az[s], acc.po[s], acc.dt[s]);
            }
        };

        while (nStack > 0) {
            int k = min(32, nStack);
            if (nStack > GG_STACK_CAP - GG_STACK_DFS_MARGIN) k = 1; // near the cap: depth-first, growth <= 1 per step
            const bool has = lane < k;
            unsigned item = has ? W.stack[nStack - 1 - lane] : 0u;
            nStack -= k;
            __syncwarp();
            int action = 0; // 1 push children, 2 Newtonian cell, 3 source bucket, 4 own bucket, 5 softened cell
            NodeW nd;
            int node = 0, img = 0;
            nd.nP = 0; nd.c0 = -1; nd.c1 = -1; nd.pLower = 0;
            double x = 0, y = 0, z = 0;
            if (has && item != 0xffffffffu) {
                node = (int)(item >> A.imgBits);
                img = (int)(item & imgMask);
                nd = load_node(&A.nodes[node]);
                x = nd.rx + s_off[3 * img];
                y = nd.ry + s_off[3 * img + 1];
                z = nd.rz + s_off[3 * img + 2];
                bool open = intersect_np(W.box, nd.fOpen2, x, y, z);
                if (nd.nP < 4) open = true; // walk.c:81 (pUpper - pLower < 3)
                if (open) {
                    if (nd.c0 >= 0) action = 1;
                    else action = (node == task.node && img == A.homeImage) ? 4 : 3; // walk.c:93
                } else {
                    double twoh2 = nd.fSoft + fSoftMax;
                    twoh2 = __dmul_rn(twoh2, twoh2);
                    bool soft = false;
                    if (!(twoh2 < nd.fOpen2)) soft = intersect_np(W.box, twoh2, x, y, z); // walk.c:122-127
                    action = soft ? 5 : 2;
                }
            }
            // children
            unsigned mPush = __ballot_sync(FULL, action == 1);
            if (action == 1) {
                int pos = nStack + 2 * __popc(mPush & lt);
                if (nd.c1 >= 0) {
                    if (pos + 1 < GG_STACK_CAP) {
                        W.stack[pos] = ((unsigned)nd.c1 << A.imgBits) | (unsigned)img;
                        W.stack[pos + 1] = ((unsigned)nd.c0 << A.imgBits) | (unsigned)img;
                    } else atomicExch(A.errFlag, 1);
                } else { // single-child cell (pkdThreadTree pkd.c:2597-2609): second slot is a no-op item
                    if (pos + 1 < GG_STACK_CAP) {
                        W.stack[pos] = ((unsigned)nd.c0 << A.imgBits) | (unsigned)img;
                        W.stack[pos + 1] = 0xffffffffu;
                    } else atomicExch(A.errFlag, 1);
                }
            }
            nStack += 2 * __popc(mPush);
            // Newtonian cells
            unsigned mCell = __ballot_sync(FULL, action == 2);
            if (action == 2) W.cbuf[nCell + __popc(mCell & lt)] = item;
            nCell += __popc(mCell);
            cntN += __popc(mCell);
            // source particles
            int np = (action == 3 || action == 4) ? nd.nP : 0;
            unsigned mBk = __ballot_sync(FULL, np > 0);
            if (mBk) {
                int incl = np;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                int total = __shfl_sync(FULL, incl, 31);
                int wpos = nPart + incl - np;
                for (int j = 0; j < np; ++j)
                    W.pbuf[wpos + j] = ((unsigned)(nd.pLower + j) << A.imgBits) | (unsigned)img;
                unsigned mOwn = __ballot_sync(FULL, action == 4);
                int own = mOwn ? __shfl_sync(FULL, np, __ffs(mOwn) - 1) : 0;
                nPart += total;
                cntP += total - own;
            }
            // softened cells: evaluated on the spot by the lane that found them
            unsigned mSoft = __ballot_sync(FULL, action == 5);
            cntS += __popc(mSoft);
            if (action == 5 && !A.walkOnly) {
                const double *Q = &A.momq[(size_t)node * 6];
#pragma unroll
                for (int s = 0; s < GG_MAX_SINKS; ++s)
                    if (s < nS)
                        softcell_on_sink(nd.fMass, nd.fSoft, Q, ((double)W.sx[s] + cenx) - x,
                                         ((double)W.sy[s] + ceny) - y, ((double)W.sz[s] + cenz) - z,
                                         (double)W.sm[s], (double)W.sh[s], acc.ax[s], acc.ay[s], acc.az[s],
                                         acc.po[s], acc.dt[s]);
            }
            __syncwarp();
            // drain full chunks (from the top of each buffer: no shifting)
            if (A.walkOnly) {
                nCell = 0;
                nPart = 0;
            } else {
                while (nCell >= 32) {
                    nCell -= 32;
                    eval_cells(nCell, 32);
                }
                while (nPart >= 32) {
                    nPart -= 32;
                    eval_parts(nPart, 32);
                }
            }
            __syncwarp();
        }
        if (!A.walkOnly) {
            if (nCell > 0) eval_cells(0, nCell);
            if (nPart > 0) eval_parts(0, nPart);
        }
        (void)pbufCap;

        // ---- reduce across the warp and write out
        if (task.group == 0 && lane == 0) {
            A.counts[3 * task.node] = cntP;
            A.counts[3 * task.node + 1] = cntS;
            A.counts[3 * task.node + 2] = cntN;
        }
        if (!A.walkOnly) {
#pragma unroll
            for (int s = 0; s < GG_MAX_SINKS; ++s) {
                if (s < nS) {
                    double vx = acc.ax[s], vy = acc.ay[s], vz = acc.az[s], vp = acc.po[s];
                    float vd = acc.dt[s];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        vx += __shfl_xor_sync(FULL, vx, o);
                        vy += __shfl_xor_sync(FULL, vy, o);
                        vz += __shfl_xor_sync(FULL, vz, o);
                        vp += __shfl_xor_sync(FULL, vp, o);
                        vd = fmaxf(vd, __shfl_xor_sync(FULL, vd, o));
                    }
                    if (lane == s) {
                        int pi = W.sidx[s];
                        A.acc[3 * (size_t)pi] = vx;
                        A.acc[3 * (size_t)pi + 1] = vy;
                        A.acc[3 * (size_t)pi + 2] = vz;
                        A.pot[pi] = vp;
                        A.dtg[pi] = (double)vd;
                    }
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
