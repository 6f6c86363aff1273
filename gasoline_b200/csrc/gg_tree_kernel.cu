// gg_tree_kernel.cu -- the two tree-gravity kernels (sm_100a): k_walk builds every sink bucket's interaction lists,
// k_eval evaluates them.  Together they replace pkdBucketWalk (walk.c:306) + pkdBucketInteract (grav.c:23).
//
//  * k_walk -- one warp per sink bucket, compiled for high occupancy (the walk is L2-latency bound: every step waits
//    for 32 scattered 64 B node records).  The warp keeps a frontier of (cell, periodic image) pairs in shared memory
//    and tests 32 of them per step, one per lane.  The test is the reference's, operation for operation and in FP64
//    without FMA contraction (INTERSECTNP walk.h:12-30; "< 4 particles => open" walk.c:81; softened-cell
//    classification walk.c:118-127), so each bucket ends up with exactly the reference's three lists -- only their
//    order differs, which the reference's results do not depend on beyond rounding.  Opened cells push their two
//    children, opened buckets contribute their particles.  List entries are 4-byte references (index << imgBits |
//    image) written in 128 B blocks of 32 into a pool in HBM; each bucket's blocks form a chain (nextBlk).  Blocks
//    come from per-warp slabs, so the global cursor sees one atomic per 32 blocks.
//  * k_eval -- one warp per (bucket, group of <= 8 active sinks).  For every block of a chain the warp stages the 32
//    sources in shared memory -- each lane converts one source to FP32 coordinates relative to the bucket centre
//    (the FP64 subtraction is done here, so FP32 displacements keep ~1e-7 relative accuracy) and copies its 128 B
//    FP32 moment record -- and evaluates them with the classic N-body tiling: every lane owns ONE sink (position,
//    mass and accumulators in registers) and loops over staged sources read with broadcast LDS.128; the 32 lanes
//    are G = 32/nSinks sub-groups that split the block between them.  Per-lane FP32 partial sums are folded into
//    FP64 accumulators after every block, and the G sub-groups are combined in FP64 in a fixed order at the end
//    (deterministic).  1/r comes from MUFU.RSQ plus one Newton step.
//
// The arithmetic is FP32 CUDA-core work: 138 FFMA/FMUL/FADD per (sink, hexadecapole cell) pair
// against the reference's score of 312 flops (grav.c:156-162), ~25 per (sink, particle) pair against 38.
#include "gg_internal.h"

#define FULL 0xffffffffu
#define CSTRIDE 9 // float4 per staged cell: (x,y,z,M) + 8 x float4 moments = 36 words -> bank offset 4 per source
#define PSTRIDE 3 // float4 per staged particle: (x,y,z,m), (h, index, -, -), pad = 12 words -> conflict-free for G <= 8

namespace {

// INTERSECTNP (walk.h:12-30): squared distance from (x,y,z) to the box <= fBall2.  Branch-free but bit-identical:
// with fMin <= fMax at most one of (fMin - x), (x - fMax) is positive, so max(max(.,.),0) selects the term the
// reference's if/else-if picks, and adding the +0.0 of an axis inside the box changes nothing.  Intrinsics pin the
// rounding of every product and sum (no FMA), matching the reference's x86-64 build.
__device__ __forceinline__ double pos_part(double a, double b) { // max(a, b, 0) for finite inputs (no NaN fix-up code)
    const double t = a > b ? a : b;
    return t > 0.0 ? t : 0.0;
}
__device__ __forceinline__ bool intersect_np(const double *box, double fBall2, double x, double y, double z) {
    const double dx = pos_part(box[0] - x, x - box[3]);
    const double dy = pos_part(box[1] - y, y - box[4]);
    const double dz = pos_part(box[2] - z, z - box[5]);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return d2 <= fBall2;
}

__device__ __forceinline__ float rsqrt_nr(float d2) {
    float y; // the bare MUFU.RSQ: rsqrtf() wraps it in a denormal-range fix-up that no squared distance here needs
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d2));
    return y * fmaf(-0.5f * d2, y * y, 1.5f); // one Newton step: MUFU.RSQ is good to ~2^-22
}

__device__ __forceinline__ NodeW load_node(const NodeW *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    int4 d = __ldg(reinterpret_cast<const int4 *>(q + 3));
    NodeW n;
    n.rx = a.x; n.ry = a.y; n.rz = b.x; n.fMass = b.y; n.fOpen2 = c.x; n.fSoft = c.y;
    n.c0 = d.x; n.c1 = d.y; n.pLower = d.z; n.nP = d.w;
    return n;
}

// 16-byte asynchronous global->shared copy THROUGH L1 (.ca): the top of the tree is re-read by every bucket and
// every periodic image, so the L1 hit rate matters (the .cg form cuda_pipeline.h picks for 16 B bypasses L1).
__device__ __forceinline__ void cp_async_ca16(void *smem, const void *gmem) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ NodeW load_node_smem(const uint4 *q) {
    const uint4 a = q[0], b = q[1], c = q[2], d = q[3];
    NodeW n;
    n.rx = __hiloint2double(a.y, a.x); n.ry = __hiloint2double(a.w, a.z);
    n.rz = __hiloint2double(b.y, b.x); n.fMass = __hiloint2double(b.w, b.z);
    n.fOpen2 = __hiloint2double(c.y, c.x); n.fSoft = __hiloint2double(c.w, c.z);
    n.c0 = (int)d.x; n.c1 = (int)d.y; n.pLower = (int)d.z; n.nP = (int)d.w;
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
// restructured so that every moment is used in exactly one FMA per force component and the per-order constants
// disappear: the record holds (2l-1)!! x the moments (gg_pack_momf), so the kernel needs only e_l = 1/r^(2l+1)
// (one FMUL each).  With A = e2 (Q'r).r, B = e3 (O'rr).r, C = e4 (H'rrr).r:
//     -phi  = M/r + A/2 + B/3 + C/4
//     a     = e2 Q'r + e3 O'rr/2 + e4 H'rrr/6 - r [ M e1 + (5A/2 + 7B/3 + 9C/4)/r^2 ]
// The three force components chain straight onto the lane's running sums (ax, ay, az); ap accumulates -phi and dtm
// the running maximum of 1/dt^2 = (m_sink + M)/r^3 (grav.c:189-190).  ~126 FP32 instructions at hexadecapole order
// (the reference scores 312 flops).  MONO64: the monopole term is left to the caller (FP64, see eval_cells), which
// gets the FP32 1/r back through pg0.
template <int ORDER, bool MONO64 = false>
__device__ __forceinline__ void cell_on_sink(const CellMom &c, float M, float dx, float dy, float dz, float ms,
                                             float &ax, float &ay, float &az, float &ap, float &dtm, float *pg0 = nullptr) {
    const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    const float d2 = xx + yy + zz;
    const float g0 = rsqrt_nr(d2);
    const float u = g0 * g0;
    const float e1 = g0 * u;
    if (MONO64) *pg0 = g0;
    dtm = fmaxf(dtm, (ms + M) * e1);
    float fp = MONO64 ? 0.f : g0 * M;
    float ts = 0.f;
    float fx = ax, fy = ay, fz = az;
    if (ORDER >= 2) {
        const float e2 = e1 * u;
        const float Qxx = c.m0.x, Qyy = c.m0.y, Qzz = c.m0.z, Qxy = c.m0.w, Qxz = c.m1.x, Qyz = c.m1.y;
        const float qx = fmaf(Qxz, dz, fmaf(Qxy, dy, Qxx * dx));
        const float qy = fmaf(Qyz, dz, fmaf(Qxy, dx, Qyy * dy));
        const float qz = fmaf(Qyz, dy, fmaf(Qxz, dx, Qzz * dz));
        const float A = e2 * fmaf(qz, dz, fmaf(qy, dy, qx * dx));
        fp = fmaf(0.5f, A, fp);
        ts = 2.5f * A;
        fx = fmaf(e2, qx, fx); fy = fmaf(e2, qy, fy); fz = fmaf(e2, qz, fz);
        if (ORDER >= 3) {
            const float e3 = e2 * u;
            const float hxx = 0.5f * xx, hyy = 0.5f * yy, hzz = 0.5f * zz;
            const float xy = dx * dy, xz = dx * dz, yz = dy * dz;
            const float Oxxx = c.m1.z, Oxyy = c.m1.w, Oxxy = c.m2.x, Oyyy = c.m2.y, Oxxz = c.m2.z, Oyyz = c.m2.w,
                        Oxyz = c.m3.x, Oxzz = c.m3.y, Oyzz = c.m3.z, Ozzz = c.m3.w;
            const float ox = fmaf(Oxzz, hzz, fmaf(Oxyz, yz, fmaf(Oxyy, hyy, fmaf(Oxxz, xz, fmaf(Oxxy, xy, Oxxx * hxx)))));
            const float oy = fmaf(Oyzz, hzz, fmaf(Oyyz, yz, fmaf(Oyyy, hyy, fmaf(Oxyz, xz, fmaf(Oxyy, xy, Oxxy * hxx)))));
            const float oz = fmaf(Ozzz, hzz, fmaf(Oyzz, yz, fmaf(Oyyz, hyy, fmaf(Oxzz, xz, fmaf(Oxyz, xy, Oxxz * hxx)))));
            const float B = e3 * fmaf(oz, dz, fmaf(oy, dy, ox * dx));
            fp = fmaf(1.f / 3.f, B, fp);
            ts = fmaf(7.f / 3.f, B, ts);
            fx = fmaf(e3, ox, fx); fy = fmaf(e3, oy, fy); fz = fmaf(e3, oz, fz);
            if (ORDER >= 4) {
                const float e4 = e3 * u;
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
                const float C = e4 * fmaf(hz, dz, fmaf(hy, dy, hx * dx));
                fp = fmaf(0.25f, C, fp);
                ts = fmaf(2.25f, C, ts);
                fx = fmaf(e4, hx, fx); fy = fmaf(e4, hy, fy); fz = fmaf(e4, hz, fz);
            }
        }
    }
    const float ta = MONO64 ? u * ts : fmaf(u, ts, e1 * M);
    ax = fmaf(-dx, ta, fx);
    ay = fmaf(-dy, ta, fy);
    az = fmaf(-dz, ta, fz);
    ap += fp;
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
struct SoftTerm {
    double ax, ay, az, pot, dt;
};
__device__ __noinline__ SoftTerm softcell_on_sink(double M, double hc, const double *Q, double dx, double dy, double dz,
                                                  double ms, double hs) {
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
    SoftTerm t;
    t.ax = -(qir3 * dx - c * qirx);
    t.ay = -(qir3 * dy - c * qiry);
    t.az = -(qir3 * dz - c * qirz);
    t.pot = -(a * M + c * qir - b * tr);
    t.dt = (ms + M) * b;
    return t;
}

// The sink box of a bucket: bbox of its ACTIVE particles (pkd.c:2916-2932), fSoftMax over ALL of them
// (walk.c:319-324) and the number of active ones.  Warp-collective.
__device__ __forceinline__ void sink_box(const TreeKernelArgs &A, const NodeW &bk, int lane, double box[6],
                                         double &fSoftMax, int &nAct) {
    box[0] = box[1] = box[2] = 1.7976931348623157e308;
    box[3] = box[4] = box[5] = -1.7976931348623157e308;
    fSoftMax = 0.0;
    nAct = 0;
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
}

// ------------------------------------------------------------------------------------------------ k_walk
#define NSTRIDE 5 // uint4 per staged node record: 64 B + 16 B pad -> LDS.128 of 8 consecutive lanes is conflict-free
struct WalkSmem {
    uint4 nstage[32 * NSTRIDE]; // the 32 node records of the current step, fetched cooperatively (4 lanes per record)
    unsigned stack[GG_STACK_CAP];     // frontier: (cell << imgBits) | image
    gg_mask_t smask[GG_STACK_CAP];     // ... and the buckets of the group for which that cell is still undecided
    double box[GG_WALK_GB][6];  // sink boxes (active particles) of the group's buckets
    double gbox[6];             // ... and the box around all of them
    double fSoftMax[GG_WALK_GB];
    int head[GG_NLIST][2], fill[GG_NLIST][2], cnt[GG_NLIST][2]; // chain state per list type (0 leaves, 1 soft, 2 Newtonian, 3 big Newtonian) x (shared, masked)
    int own[GG_WALK_GB];        // particles of the bucket itself met in the home image (walk.c:93)
    int bnode[GG_WALK_GB];
    int seedNext;               // first image whose root has not been put on the frontier yet
};

__host__ __device__ inline size_t walk_smem_bytes() { return (sizeof(WalkSmem) + 15) & ~(size_t)15; }

// squared distance from a point to the FARTHEST corner of a box: an upper bound, term by term and rounding by
// rounding, of INTERSECTNP's squared distance to any box inside it
__device__ __forceinline__ double far_dist2(const double *box, double x, double y, double z) {
    const double dx = fmax(x - box[0], box[3] - x), dy = fmax(y - box[1], box[4] - y), dz = fmax(z - box[2], box[5] - z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ double near_dist2(const double *box, double x, double y, double z) {
    const double dx = pos_part(box[0] - x, x - box[3]);
    const double dy = pos_part(box[1] - y, y - box[4]);
    const double dz = pos_part(box[2] - z, z - box[5]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

struct Slab {
    int base, used;
};

// Append <= 32 entries (one per lane with `has`) to chain (type, d); d = 1 also records the entry's bucket mask.
// The head block of a chain is the one being filled; older blocks are full.  Blocks come from the warp's slab; a
// new slab costs one atomic.  Warp-collective.
#ifndef GG_WALK_NOINLINE
#define GG_WALK_NOINLINE 0 // 1: append is a real function, 2: distribute is (k_walk's loop body is ~100 KB of code when both are inlined)
#endif
#if GG_WALK_NOINLINE == 1
#define GG_APPEND_INLINE __noinline__
#else
#define GG_APPEND_INLINE __forceinline__
#endif
#if GG_WALK_NOINLINE == 2
#define GG_DISTRIBUTE_INLINE __noinline__
#else
#define GG_DISTRIBUTE_INLINE __forceinline__
#endif
__device__ GG_APPEND_INLINE void append(const TreeKernelArgs &A, WalkSmem &W, int type, int d, bool has, unsigned entry,
                                       unsigned mask, int lane, unsigned lt, Slab &slab) {
    const unsigned m = __ballot_sync(FULL, has);
    if (!m) return;
    const int n = __popc(m), pos = __popc(m & lt);
    const int head = W.head[type][d], fill = W.fill[type][d];
    const int room = head >= 0 ? 32 - fill : 0;
    int nb = head;
    if (n > room) {
        if (slab.used == GG_SLAB_BLOCKS) {
            int b = 0;
            if (lane == 0) b = atomicAdd(A.poolCursor, GG_SLAB_BLOCKS);
            slab.base = __shfl_sync(FULL, b, 0);
            slab.used = 0;
        }
        nb = slab.base + slab.used++;
    }
    __syncwarp();
    const bool fits = nb < A.capBlocks; // beyond the pool: the host sees cursor > capBlocks, grows the pool and reruns
    if (has && !A.walkOnly) {
        size_t at = (size_t)head * 32 + fill + pos;
        if (pos >= room) at = (size_t)nb * 32 + (pos - room);
        if (pos < room || fits) {
            A.pool[at] = entry;
            if (d) A.poolMask[at] = (gg_mask_t)mask;
        }
    }
    if (lane == 0) {
        if (n > room) {
            if (fits && !A.walkOnly) A.nextBlk[nb] = head;
            W.head[type][d] = fits ? nb : head;
            W.fill[type][d] = fits ? n - room : fill;
        } else W.fill[type][d] = fill + n;
        W.cnt[type][d] += n;
    }
    __syncwarp();
}

// One list type of one step: lanes with dm == all go to the group's shared chain, the others to its masked chain.
// myCnt (lane b counts for bucket b of the group) gets the per-bucket number of masked entries -- of particles for
// leaves (np > 0).
__device__ GG_DISTRIBUTE_INLINE void distribute(const TreeKernelArgs &A, WalkSmem &W, int type, unsigned dm, unsigned all,
                                           int nB, unsigned entry, int np, int lane, unsigned lt, Slab &slab,
                                           int &myCnt, int &sharedP, int &myLeaves) {
    const unsigned mAny = __ballot_sync(FULL, dm != 0);
    if (!mAny) return;
    const bool shared = dm == all;
    const unsigned mSh = __ballot_sync(FULL, shared);
    if (mSh) {
        if (type == 0) sharedP += __reduce_add_sync(FULL, shared ? np : 0);
        append(A, W, type, 0, shared, entry, 0u, lane, lt, slab);
    }
    if (mAny != mSh) {
        const unsigned pm = shared ? 0u : dm;
#pragma unroll
        for (int b = 0; b < GG_WALK_GB; ++b) { // (bits >= nB are never set)
            const bool has = (pm >> b) & 1u;
            const unsigned mb = __ballot_sync(FULL, has);
            if (type == 0) {
                if (mb) {
                    const int c = __reduce_add_sync(FULL, has ? np : 0);
                    if (lane == b) { myCnt += c; myLeaves += __popc(mb); }
                }
            } else if (lane == b) myCnt += __popc(mb);
        }
        append(A, W, type, 1, pm != 0, entry, pm, lane, lt, slab);
    }
}

// The sink boxes of up to four buckets of <= 8 particles at once: lane = 8 * (bucket within the pass) + particle.
// Same results as sink_box (min / max / max are exact and order-independent).
__device__ __forceinline__ void sink_box4(const TreeKernelArgs &A, int pLower, int nP, int lane, double box[6],
                                          double &fSoftMax) {
    box[0] = box[1] = box[2] = 1.7976931348623157e308;
    box[3] = box[4] = box[5] = -1.7976931348623157e308;
    fSoftMax = 0.0;
    const int j = lane & 7;
    if (j < nP) {
        const int pi = pLower + j;
        fSoftMax = A.hsoft[pi];
        if (A.active ? (A.active[pi] != 0) : true) {
            const PartS p = load_part(&A.parts[pi]);
            box[0] = box[3] = p.x; box[1] = box[4] = p.y; box[2] = box[5] = p.z;
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            box[k] = fmin(box[k], __shfl_xor_sync(FULL, box[k], o));
            box[3 + k] = fmax(box[3 + k], __shfl_xor_sync(FULL, box[3 + k], o));
        }
        fSoftMax = fmax(fSoftMax, __shfl_xor_sync(FULL, fSoftMax, o));
    }
}

// One warp walks the tree ONCE for a group of up to GG_WALK_GB consecutive sink buckets.  Every frontier item
// carries the set of buckets that have opened all of its ancestors; the reference's per-bucket opening test
// (walk.c:81-127) is evaluated for exactly those buckets, short-cut by two conservative tests against the box of
// the whole group that are monotone in floating point (so they can never disagree with the per-bucket result).
template <bool SEED_BATCHES>
__global__ void __launch_bounds__(GG_WALK_WARPS * 32, GG_WALK_MIN_CTAS) k_walk(const TreeKernelArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the image offsets live behind the warps' walk state, sized by the call's image count (27 in a periodic box)
    double *s_off = reinterpret_cast<double *>(smem_raw + GG_WALK_WARPS * walk_smem_bytes());
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < A.nImages * 3; i += blockDim.x) s_off[i] = A.imgOff[i];
    __syncthreads();
    WalkSmem &W = *reinterpret_cast<WalkSmem *>(smem_raw + warp * walk_smem_bytes());
    const unsigned imgMask = (1u << A.imgBits) - 1u;
    Slab slab{0, GG_SLAB_BLOCKS};
    const int nGroups = (A.nBuckets + GG_WALK_GB - 1) / GG_WALK_GB;
    const bool haveBig = A.bigMass < 1.0e308; // periodic boxes only (open boundaries: the class is empty, nothing is spent on it)

    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(A.taskCounter, 1);
        g = __shfl_sync(FULL, g, 0);
        if (g >= nGroups) break;
        const int b0 = g * GG_WALK_GB, nB = min(GG_WALK_GB, A.nBuckets - b0);
        const unsigned all = (1u << nB) - 1u;
        // ---- the group's sink boxes
        double gbox[6] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308,
                          -1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
        double gSoftMax = 0.0;
        {
            // lane b < nB: bucket b's node
            int myNode = 0, myLower = 0, myNP = 0;
            if (lane < nB) {
                myNode = A.bucketNode[b0 + lane];
                const int4 d = __ldg(reinterpret_cast<const int4 *>(&A.nodes[myNode]) + 3);
                myLower = d.z; myNP = d.w;
                W.bnode[lane] = myNode; W.own[lane] = 0;
            }
            const bool small = __all_sync(FULL, myNP <= 8);
            for (int bb = 0; bb < nB; bb += 4) {
                double box[6], fSoftMax;
                if (small) { // four buckets per pass, eight lanes each
                    const int b = min(bb + (lane >> 3), nB - 1);
                    const int pl = __shfl_sync(FULL, myLower, b), np = __shfl_sync(FULL, myNP, b);
                    sink_box4(A, pl, np, lane, box, fSoftMax);
                    if (A.sunNode >= 0 && __shfl_sync(FULL, myNode, b) == A.sunNode) {
                        // the dummy sink of the bDoSun pass: the reference gives its cell a box of +-1e-14 about the
                        // origin instead of the particle's own position (pkd.c:3017-3021)
                        box[0] = box[1] = box[2] = -A.sunBox;
                        box[3] = box[4] = box[5] = A.sunBox;
                    }
                    if (bb + (lane >> 3) < nB) {
                        const int k = lane & 7;
                        const double v = k == 0 ? box[0] : k == 1 ? box[1] : k == 2 ? box[2] : k == 3 ? box[3]
                                         : k == 4 ? box[4] : k == 5 ? box[5] : fSoftMax;
                        if (k < 6) W.box[b][k] = v;
                        else if (k == 6) W.fSoftMax[b] = v;
                    }
                } else {
                    for (int b = bb; b < min(bb + 4, nB); ++b) {
                        NodeW bk;
                        bk.pLower = __shfl_sync(FULL, myLower, b); bk.nP = __shfl_sync(FULL, myNP, b);
                        int nAct;
                        sink_box(A, bk, lane, box, fSoftMax, nAct);
                        const double v = lane == 0 ? box[0] : lane == 1 ? box[1] : lane == 2 ? box[2] : lane == 3 ? box[3]
                                         : lane == 4 ? box[4] : lane == 5 ? box[5] : fSoftMax;
                        if (lane < 6) W.box[b][lane] = v;
                        else if (lane == 6) W.fSoftMax[b] = v;
                    }
                }
            }
            __syncwarp();
            for (int b = 0; b < nB; ++b) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { gbox[k] = fmin(gbox[k], W.box[b][k]); gbox[3 + k] = fmax(gbox[3 + k], W.box[b][3 + k]); }
                gSoftMax = fmax(gSoftMax, W.fSoftMax[b]);
            }
            {
                const double v = lane == 0 ? gbox[0] : lane == 1 ? gbox[1] : lane == 2 ? gbox[2] : lane == 3 ? gbox[3]
                                 : lane == 4 ? gbox[4] : gbox[5];
                if (lane < 6) W.gbox[lane] = v;
            }
        }
        if (lane < 2 * GG_NLIST) {
            (&W.head[0][0])[lane] = -1; (&W.fill[0][0])[lane] = 0; (&W.cnt[0][0])[lane] = 0;
        }
        int myP = 0, myS = 0, myN = 0, myB = 0, myL = 0, sharedP = 0, unused = 0; // lane b: masked entries of bucket b
        // the frontier starts with the root under every image offset (walk.c:325-337).  More images than GG_SEED_IMAGES
        // (nReplicas > 3; SEED_BATCHES) are seeded in batches, the next one when the frontier has run empty; the batch
        // cursor lives in shared memory (k_walk has no register to spare for it).
        int nStack = 0;
        if (SEED_BATCHES) {
            if (lane == 0) W.seedNext = 0;
        } else {
            nStack = A.nImages;
            for (int i = lane; i < A.nImages; i += 32) {
                W.stack[i] = ((unsigned)A.rootNode << A.imgBits) | (unsigned)i;
                W.smask[i] = (gg_mask_t)all;
            }
        }
        __syncwarp();

        for (;;) {
            if (SEED_BATCHES) {
                if (nStack == 0) {
                    const int nextImg = W.seedNext;
                    if (nextImg >= A.nImages) break;
                    const int nSeed = min(GG_SEED_IMAGES, A.nImages - nextImg);
                    for (int i = lane; i < nSeed; i += 32) {
                        W.stack[i] = ((unsigned)A.rootNode << A.imgBits) | (unsigned)(nextImg + i);
                        W.smask[i] = (gg_mask_t)all;
                    }
                    nStack = nSeed;
                    __syncwarp();
                    if (lane == 0) W.seedNext = nextImg + nSeed;
                    __syncwarp();
                }
            } else if (nStack <= 0) break;
            int k = min(32, nStack);
            if (nStack > GG_STACK_CAP - GG_STACK_DFS_MARGIN) k = 1; // near the cap: depth-first, growth <= 1 per step
            unsigned item = 0xffffffffu, mask = 0;
            if (lane < k) { item = W.stack[nStack - 1 - lane]; mask = W.smask[nStack - 1 - lane]; }
            nStack -= k;
            __syncwarp();
            // fetch the k node records COALESCED: 4 lanes per 64 B record, so a warp-wide 16 B access touches 8 cache
            // lines instead of 32; cp.async lands them in shared memory
            const int node = item != 0xffffffffu ? (int)(item >> A.imgBits) : -1;
            {
                const int piece = lane & 3, sub = lane >> 2;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rec = 8 * i + sub;
                    const int rn = __shfl_sync(FULL, node, rec);
                    if (rn >= 0)
                        cp_async_ca16(&W.nstage[rec * NSTRIDE + piece], reinterpret_cast<const uint4 *>(&A.nodes[rn]) + piece);
                }
                cp_async_wait_all();
            }
            __syncwarp();
            // ---- decide, per bucket of the item's mask: open / Newtonian cell / softened cell
            unsigned mOpen = 0, mSoft = 0, mNewt = 0, mBig = 0, amb = 0;
            int img = 0, np = 0, c0 = -1, c1 = -1, nPnode = 0;
            double x = 0.0, y = 0.0, z = 0.0, fOpen2 = 0.0, fSoftC = 0.0;
            bool isBig = false;
            if (node >= 0) {
                img = (int)(item & imgMask);
                const NodeW nd = load_node_smem(&W.nstage[lane * NSTRIDE]);
                x = nd.rx + s_off[3 * img]; y = nd.ry + s_off[3 * img + 1]; z = nd.rz + s_off[3 * img + 2];
                fOpen2 = nd.fOpen2; fSoftC = nd.fSoft; isBig = nd.fMass >= A.bigMass; c0 = nd.c0; c1 = nd.c1; nPnode = nd.nP;
                if (nd.nP < 4) mOpen = mask; // walk.c:81 (pUpper - pLower < 3)
                else if (near_dist2(W.gbox, x, y, z) <= fOpen2) { // else: no bucket inside the group box opens it
                    if (far_dist2(W.gbox, x, y, z) <= fOpen2) mOpen = mask; // every bucket inside the group box does
                    else amb = mask;                                      // the buckets have to be asked one by one
                }
            }
            // the per-bucket tests of all lanes, spread over the warp: one (item, bucket) pair per lane and round
            {
                const int na = __popc(amb);
                int incl = na;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const int nPairs = __shfl_sync(FULL, incl, 31);
                if (nPairs > 0) {
                    __syncwarp(); // every lane has its record out of nstage: reuse it as scratch
                    double *tq = reinterpret_cast<double *>(W.nstage);                 // [32][4] x, y, z, fOpen2
                    unsigned short *pairs = reinterpret_cast<unsigned short *>(tq + 128); // [<= 32 * GG_WALK_GB]
                    unsigned *mres = reinterpret_cast<unsigned *>(pairs + 32 * GG_WALK_GB); // [32]
                    static_assert(1024 + 64 * GG_WALK_GB + 128 <= sizeof(W.nstage), "pair scratch must fit the node stage");
                    mres[lane] = 0u;
                    if (amb) {
                        tq[4 * lane] = x; tq[4 * lane + 1] = y; tq[4 * lane + 2] = z; tq[4 * lane + 3] = fOpen2;
                        int o = incl - na;
                        for (unsigned mm = amb; mm; mm &= mm - 1) pairs[o++] = (unsigned short)((lane << 4) | (__ffs(mm) - 1));
                    }
                    __syncwarp();
                    for (int p = lane; p < nPairs; p += 32) {
                        const int pr = pairs[p], o = pr >> 4, b = pr & 15;
                        if (intersect_np(W.box[b], tq[4 * o + 3], tq[4 * o], tq[4 * o + 1], tq[4 * o + 2]))
                            atomicOr(&mres[o], 1u << b);
                    }
                    __syncwarp();
                    if (amb) mOpen = mres[lane];
                    __syncwarp();
                }
            }
            if (node >= 0) {
                const unsigned mAcc = mask & ~mOpen;
                if (mAcc) { // walk.c:118-127
                    double t2 = fSoftC + gSoftMax;
                    t2 = __dmul_rn(t2, t2);
                    if (!(t2 < fOpen2)) // (fSoft + fSoftMax_b)^2 <= t2 for every bucket: only then can one be soft
                        for (unsigned mm = mAcc; mm; mm &= mm - 1) {
                            const int b = __ffs(mm) - 1;
                            double twoh2 = fSoftC + W.fSoftMax[b];
                            twoh2 = __dmul_rn(twoh2, twoh2);
                            if (!(twoh2 < fOpen2) && intersect_np(W.box[b], twoh2, x, y, z)) mSoft |= 1u << b;
                        }
                    mNewt = mAcc & ~mSoft;
                    // periodic boxes: the few massive far cells go to a list of their own, whose monopoles k_eval
                    // evaluates in FP64 (see eval_cells); same cell = same class for every bucket of the group
                    if (isBig) { mBig = mNewt; mNewt = 0; }
                }
                if (mOpen && c0 < 0) { // an opened bucket: all its particles are sources (walk.c:93-114)
                    np = nPnode;
                    if (img == A.homeImage)
                        for (unsigned mm = mOpen; mm; mm &= mm - 1) {
                            const int b = __ffs(mm) - 1;
                            if (W.bnode[b] == node) W.own[b] = np;
                        }
                }
            }
            // children (a single-child cell, pkdThreadTree pkd.c:2597-2609, pushes a no-op as second item)
            const bool push = mOpen && c0 >= 0;
            const unsigned mPush = __ballot_sync(FULL, push);
            if (push) {
                const int pos = nStack + 2 * __popc(mPush & lt);
                if (pos + 1 < GG_STACK_CAP) {
                    W.stack[pos] = c1 >= 0 ? (((unsigned)c1 << A.imgBits) | (unsigned)img) : 0xffffffffu;
                    W.stack[pos + 1] = ((unsigned)c0 << A.imgBits) | (unsigned)img;
                    W.smask[pos] = (gg_mask_t)mOpen;
                    W.smask[pos + 1] = (gg_mask_t)mOpen;
                } else atomicExch(A.errFlag, 1);
            }
            nStack += 2 * __popc(mPush);
            distribute(A, W, 2, mNewt, all, nB, item, 0, lane, lt, slab, myN, unused, unused);
            if (haveBig) distribute(A, W, 3, mBig, all, nB, item, 0, lane, lt, slab, myB, unused, unused);
            distribute(A, W, 1, mSoft, all, nB, item, 0, lane, lt, slab, myS, unused, unused);
            distribute(A, W, 0, np > 0 ? mOpen : 0u, all, nB, item, np, lane, lt, slab, myP, sharedP, myL);
            __syncwarp();
        }
        // ---- hand the chains over, and what pkdBucketWalk reports per bucket (walk.c:175-177)
        if (lane < nB) {
            int *c = &A.counts[3 * W.bnode[lane]];
            c[0] = sharedP + myP - W.own[lane];
            c[1] = W.cnt[1][0] + myS;
            c[2] = W.cnt[2][0] + myN + W.cnt[3][0] + myB; // pkd->nCellNewt: ordinary + big
            int *e = &A.bucketCnt[GG_NLIST * (b0 + lane)]; // list ENTRIES (a leaf entry stands for all particles of a bucket)
            e[0] = W.cnt[0][0] + myL; e[1] = c[1]; e[2] = W.cnt[2][0] + myN; e[3] = W.cnt[3][0] + myB;
            A.bucketTot[b0 + lane] = (long long)e[0] + e[1] + e[2] + e[3];
        }
        if (lane < 2 * GG_NLIST) {
            A.groupHead[2 * GG_NLIST * g + lane] = (&W.head[0][0])[lane];
            A.groupCnt[2 * GG_NLIST * g + lane] = (&W.cnt[0][0])[lane];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ k_scatter
// One warp per (walk group, list type, chain): copy the chain into the contiguous per-bucket lists k_eval streams
// through (shared entries to every bucket of the group, masked entries to the buckets of their mask).  Bucket b's
// lists start at bucketOff[b]: big Newtonian cells, Newtonian cells, softened cells, leaves; within a type the group's
// shared entries come first, so all eight warps of a group know where to write without talking to each other.
// One thread: do the walk's results fit what k_scatter / k_eval were launched with?  (frontier overflow flag, blocks taken
// from the chain pool, total list entries.)  The host reads the same numbers after the evaluation and, if not, grows the
// buffers and runs the evaluation again -- the common case costs no synchronisation between the walk and the evaluation.
__global__ void k_guard(const int *errFlag, const int *poolCursor, int capBlocks, const long long *bucketOff, int nBuckets,
                        long long capListEntries, int *okFlag) {
    *okFlag = (*errFlag == 0 && *poolCursor <= capBlocks && bucketOff[nBuckets] <= capListEntries) ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_scatter(const TreeKernelArgs A) {
    if (A.okFlag && !*A.okFlag) return;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int nGroups = (A.nBuckets + GG_WALK_GB - 1) / GG_WALK_GB;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int g = w / (2 * GG_NLIST), sub = w - 2 * GG_NLIST * g, type = sub >> 1, src = sub & 1;
    if (g >= nGroups) return;
    int blk = A.groupHead[2 * GG_NLIST * g + 2 * type + src];
    if (blk < 0) return;
    const int total = A.groupCnt[2 * GG_NLIST * g + 2 * type + src];
    const int b0 = g * GG_WALK_GB, nB = min(GG_WALK_GB, A.nBuckets - b0);
    const unsigned all = (1u << nB) - 1u;
    // lane b < nB: write cursor of bucket b
    long long cur = 0;
    if (lane < nB) {
        const int *e = &A.bucketCnt[GG_NLIST * (b0 + lane)];
        cur = A.bucketOff[b0 + lane] + (type <= 2 ? e[3] : 0) + (type <= 1 ? e[2] : 0) + (type == 0 ? e[1] : 0) +
              (src ? A.groupCnt[2 * GG_NLIST * g + 2 * type] : 0);
    }
    const int cnt0 = total - 32 * ((total - 1) / 32); // the head block is the partial one
    // the chain is a linked list (one dependent load per block): keep the NEXT block's entries and link in flight while
    // the current block is distributed
    unsigned it = 0, mk = 0;
    if (lane < cnt0) {
        it = A.pool[(size_t)blk * 32 + lane];
        mk = src ? (unsigned)A.poolMask[(size_t)blk * 32 + lane] : all;
    }
    int nxt = A.nextBlk[blk];
    while (blk >= 0) {
        unsigned itN = 0, mkN = 0;
        int nxtN = -1;
        if (nxt >= 0) {
            itN = A.pool[(size_t)nxt * 32 + lane];
            mkN = src ? (unsigned)A.poolMask[(size_t)nxt * 32 + lane] : all;
            nxtN = A.nextBlk[nxt];
        }
#pragma unroll 2
        for (int b = 0; b < nB; ++b) {
            const bool has = (mk >> b) & 1u;
            const unsigned m = __ballot_sync(FULL, has);
            if (!m) continue;
            const long long base = __shfl_sync(FULL, cur, b);
            if (has) A.lists[base + __popc(m & lt)] = it;
            if (lane == b) cur += __popc(m);
        }
        blk = nxt; nxt = nxtN; it = itN; mk = mkN;
    }
}

// ------------------------------------------------------------------------------------------------ k_eval
#define PCAP (32 * CSTRIDE / PSTRIDE) // particles staged at once (96)
#ifndef MONO_BS
#define MONO_BS 31
#endif
template <int NRAW>
struct EvalSmemT {
    float4 stage[2][32 * CSTRIDE]; // double-buffered staged block; [0] also the sink hand-out / final reduction scratch
    union {
        // cells: FP64 (x, y, z, M) of the block in flight, converted to sink-centred FP32 on arrival.  The MONO64 loop
        // (big cells of periodic boxes; NRAW = 2) double-buffers it like `stage`: its conversion leaves the FP64
        // sink-centred position behind for the hot loop.  Its blocks hold <= MONO_BS = 31 cells: with 32 the CTA would be
        // 42 bytes over the 45 670 B that let five of them share an SM (profiles/r02_k_eval_c4.md: half blocks of 16 cost
        // 130 staging instructions per 2.9 trips of the hot loop)
        double raw[NRAW == 2 ? 2 * MONO_BS * 4 : 32 * 4];
        struct {                   // leaves (never in flight together with cells)
            int lstart[32], lpart[32]; // first staging slot and first particle of each leaf of the batch
            unsigned char owner[PCAP]; // staging slot -> leaf of the batch
            unsigned short limg[32];
        };
    };
};

__device__ __forceinline__ void cp_async_cg16(void *smem, const void *gmem) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory");
}

// Per-lane evaluation state of one k_eval task: the lane's sink, its FP32 block sums and FP64 running sums.
struct Sink {
    float sx, sy, sz, ms, hs;
    int sidx;
    float ax, ay, az, ap, dtm;
    double dax, day, daz, dap;
    double sxd, syd, szd; // MONO64 only: the sink's position relative to the bucket centre in FP64
    __device__ __forceinline__ void fold() { // FP32 partial sums of one block -> FP64
        dax += (double)ax; day += (double)ay; daz += (double)az; dap -= (double)ap; // ap accumulates -phi
        ax = ay = az = ap = 0.f;
    }
};

struct EvalCtx {
    double cenx, ceny, cenz;
    int nS, G, q;
    bool worker;
    unsigned imgMask;
};

// Start the asynchronous gather of one block of <= 32 Newtonian cells (entry `it` per lane, lanes < cnt) into stage
// buffer `buf`: the FP64 (x,y,z,M) half of the walk record by its own lane, the 128 B FP32 moment record
// COALESCED, LPR lanes per record (a warp-wide 16 B access touches 32/LPR lines instead of 32).
template <int ORDER, class SM>
__device__ __forceinline__ void gather_cells(const TreeKernelArgs &A, SM &W, int buf, int rbuf, unsigned it, int cnt, int lane) {
    int cn = 0;
    if (lane < cnt) {
        cn = (int)(it >> A.imgBits);
        const char *src = reinterpret_cast<const char *>(&A.nodes[cn]);
        cp_async_cg16(&W.raw[rbuf + 4 * lane], src);
        cp_async_cg16(&W.raw[rbuf + 4 * lane + 2], src + 16);
    }
    if (ORDER >= 2) {
        constexpr int NP = ORDER == 2 ? 2 : (ORDER == 3 ? 4 : 8); // float4 pieces of the record in use
        constexpr int PPL = GG_GATHER_PPL < NP ? GG_GATHER_PPL : NP; // consecutive pieces one lane copies
        constexpr int LPR = NP / PPL;                                // lanes per record
        const int piece = (lane & (LPR - 1)) * PPL, sub = lane / LPR;
        // one shuffle, one 64-bit multiply-add and the copies per record: the lane's piece offset is folded into both
        // bases once, and the shared-memory address of trip i is the base plus a compile-time constant
        const unsigned sdst = (unsigned)__cvta_generic_to_shared(&W.stage[buf][sub * CSTRIDE + 1 + piece]);
        const char *gsrc = reinterpret_cast<const char *>(A.momf) + piece * 16;
#pragma unroll
        for (int i = 0; i < LPR; ++i) {
            const int rec = i * (32 / LPR) + sub;
            const unsigned rn = (unsigned)__shfl_sync(FULL, cn, rec);
            unsigned long long src;
            asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(src) : "r"(rn), "l"(gsrc));
            // .ca: the moment records go through L1 -- consecutive sink buckets (and, in periodic boxes, the 27 images of
            // one bucket's walk) list the same cells again and again; measured on the 128^3 box: k_eval 6.84 -> 5.85 ms
            // (the wait for this gather was 18 % of all stall samples), Plummer unchanged.  The 32 B node halves stay
            // .cg: through L1 they cost 5 % on both workloads.
            if (rec < cnt) {
#pragma unroll
                for (int pc = 0; pc < PPL; ++pc)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sdst + i * (32 / LPR) * CSTRIDE * 16 + pc * 16),
                                 "l"(src + pc * 16) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// The Newtonian-cell list (ILCN) of the bucket: n entries at L.  Software pipeline: while block i is evaluated
// (QEVAL to ORDER), block i+1 is being gathered into the other stage buffer and the entries of block i+2 are
// being loaded.
//
// MONO64 (periodic boxes): the monopole term of every cell is evaluated in FP64.  In a nearly uniform periodic box the
// 27 image sums cancel to a small net force and potential (|a_net| << sum |a_term|, |phi| ~ 1e-3 of the tree sum), so
// the FP32 rounding of the big far-image monopoles (1e-7 each) is what limits the result, and the smaller the
// perturbations (the larger the box in particles, at fixed displacement in grid units) the worse.  One FP64 Newton
// step on the FP32 1/r against the FP64 displacement (19 DFMA-pipe instructions per pair) removes that error; the
// quadrupole and higher terms (<= ~10 % of the monopole, FP32) stay as they are.
template <int ORDER, bool MONO64, class SM>
__device__ __forceinline__ void eval_cells(const TreeKernelArgs &A, SM &W, const double *s_off, const EvalCtx &E,
                                           Sink &K, const unsigned *L, int n, int lane) {
    if (n <= 0) return;
    // block size = the largest multiple of G that fits the 32 staging slots: the G sub-groups then make the same number
    // of trips through a full block (with 32 cells and G = 6 two sub-groups would run a 6th trip alone: 11 % idle)
    const int BSMAX = MONO64 ? MONO_BS : 32; // (MONO64: the FP64 staging is double-buffered, see EvalSmemT)
    const int BS = (GG_EVAL_BSG && E.G <= BSMAX) ? BSMAX - (BSMAX % E.G) : BSMAX; // (G = 32 sub-groups of one sink: whole blocks)
    const int nBlk = (n + BS - 1) / BS;
    unsigned itCur = lane < min(BS, n) ? L[lane] : 0u;
    gather_cells<ORDER>(A, W, 0, 0, itCur, min(BS, n), lane);
    unsigned itNext = (lane < BS && BS + lane < n) ? L[BS + lane] : 0u;
#pragma unroll 1
    for (int i = 0; i < nBlk; ++i) {
        const int buf = i & 1, rbuf = MONO64 ? 4 * MONO_BS * buf : 0, cnt = min(BS, n - BS * i);
        cp_async_wait_all();
        __syncwarp();
        if (lane < cnt) { // FP64 subtraction of the sink-bucket centre, then FP32
            const int ci = (int)(itCur & E.imgMask);
            const double2 p01 = *reinterpret_cast<const double2 *>(&W.raw[rbuf + 4 * lane]);
            const double2 p23 = *reinterpret_cast<const double2 *>(&W.raw[rbuf + 4 * lane + 2]);
            const double rx = (p01.x + s_off[3 * ci]) - E.cenx, ry = (p01.y + s_off[3 * ci + 1]) - E.ceny,
                         rz = (p23.x + s_off[3 * ci + 2]) - E.cenz;
            W.stage[buf][lane * CSTRIDE] = make_float4((float)rx, (float)ry, (float)rz, (float)p23.y);
            if (MONO64) {
                *reinterpret_cast<double2 *>(&W.raw[rbuf + 4 * lane]) = make_double2(rx, ry);
                W.raw[rbuf + 4 * lane + 2] = rz;
            }
        }
        __syncwarp();
        if (i + 1 < nBlk) {
            gather_cells<ORDER>(A, W, buf ^ 1, MONO64 ? 4 * MONO_BS * (buf ^ 1) : 0, itNext, min(BS, n - BS * (i + 1)), lane);
            itCur = itNext;
            const int k = BS * (i + 2) + lane;
            itNext = (lane < BS && k < n) ? L[k] : 0u;
        }
        if (E.worker) {
            for (int j = E.q; j < cnt; j += E.G) {
                const float4 *S = &W.stage[buf][j * CSTRIDE];
                const float4 pc = S[0];
                CellMom c;
                c.m0 = S[1]; c.m1 = S[2];
                if (ORDER >= 3) { c.m2 = S[3]; c.m3 = S[4]; }
                if (ORDER >= 4) { c.m4 = S[5]; c.m5 = S[6]; c.m6 = S[7]; c.m7 = S[8]; }
                float g0;
                cell_on_sink<ORDER, MONO64>(c, pc.w, K.sx - pc.x, K.sy - pc.y, K.sz - pc.z, K.ms, K.ax, K.ay, K.az, K.ap,
                                            K.dtm, &g0);
                if (MONO64) {
                    const double2 q01 = *reinterpret_cast<const double2 *>(&W.raw[rbuf + 4 * j]);
                    const double2 q23 = *reinterpret_cast<const double2 *>(&W.raw[rbuf + 4 * j + 2]);
                    const double ddx = K.sxd - q01.x, ddy = K.syd - q01.y, ddz = K.szd - q23.x;
                    const double d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                    const double y = (double)g0;
                    const double e = fma(-(d2 * y), y, 1.0); // 1 - d2 y^2: what the FP32 1/r (and FP32 dx) got wrong
                    const double w = fma(0.5 * y, e, y);     // one Newton step: 1/r to ~1e-14
                    const double Mw = q23.y * w;
                    const double Mw3 = Mw * (w * w);
                    K.dap -= Mw;
                    K.dax -= ddx * Mw3; K.day -= ddy * Mw3; K.daz -= ddz * Mw3;
                }
            }
        }
        K.fold();
        __syncwarp();
    }
}

// <= 32 opened source BUCKETS (leaves), each standing for all its particles -- incl. the sink bucket itself
// (intra-bucket pairs, grav.c:211-242).  SPLINE-softened monopoles (ILP, grav.c:89-108).
template <class SM>
__device__ __forceinline__ void eval_leaves(const TreeKernelArgs &A, SM &W, const double *s_off, const EvalCtx &E,
                                            Sink &K, unsigned it, int cnt, int lane) {
    int np = 0, pl = 0, ci = 0;
    if (lane < cnt) {
        ci = (int)(it & E.imgMask);
        const int4 d = __ldg(reinterpret_cast<const int4 *>(&A.nodes[it >> A.imgBits]) + 3);
        pl = d.z; np = d.w;
    }
    int incl = np;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    // batches of consecutive leaves that fit the staging buffer (a leaf holds <= GG_MAX_BUCKET <= PCAP particles)
    int r0 = 0;
    while (r0 < cnt) {
        const int base = __shfl_sync(FULL, incl - np, r0);
        const unsigned mIn = __ballot_sync(FULL, lane >= r0 && lane < cnt && incl - base <= PCAP);
        const int r1 = r0 + max(1, __popc(mIn)); // (a leaf beyond PCAP cannot exist: GG_MAX_BUCKET <= PCAP; never spin)
        const int nStaged = min(__shfl_sync(FULL, incl, r1 - 1) - base, PCAP);
        if (lane >= r0 && lane < r1) {
            const int s0 = incl - np - base;
            W.lstart[lane] = s0; W.lpart[lane] = pl; W.limg[lane] = (unsigned short)ci;
            for (int j = 0; j < np; ++j) W.owner[s0 + j] = (unsigned char)lane;
        }
        __syncwarp();
        for (int s = lane; s < nStaged; s += 32) {
            const int r = W.owner[s], ri = W.limg[r];
            const int pi = W.lpart[r] + (s - W.lstart[r]);
            const PartS p = load_part(&A.parts[pi]);
            float4 *S = &W.stage[0][s * PSTRIDE];
            S[0] = make_float4((float)((p.x + s_off[3 * ri]) - E.cenx), (float)((p.y + s_off[3 * ri + 1]) - E.ceny),
                               (float)((p.z + s_off[3 * ri + 2]) - E.cenz), p.m);
            // a particle does not act on itself (grav.c:211): remember who it is, in the home image only
            S[1] = make_float4(p.h, __int_as_float(ri == A.homeImage ? pi : -1), 0.f, 0.f);
        }
        __syncwarp();
        if (E.worker) {
#pragma unroll 2
            for (int j = E.q; j < nStaged; j += E.G) {
                const float4 pp = W.stage[0][j * PSTRIDE];
                const float2 ph = *reinterpret_cast<const float2 *>(&W.stage[0][j * PSTRIDE + 1]);
                if (__float_as_int(ph.y) == K.sidx) continue;
                float fx, fy, fz, fp, fdt;
                part_on_sink(pp.w, ph.x, K.sx - pp.x, K.sy - pp.y, K.sz - pp.z, K.ms, K.hs, fx, fy, fz, fp, fdt);
                K.ax += fx; K.ay += fy; K.az += fz; K.ap += fp;
                K.dtm = fmaxf(K.dtm, fdt);
            }
        }
        K.fold();
        __syncwarp();
        r0 = r1;
    }
}

// <= 32 softened cells (ILCS, rare): FP64, the first nS lanes each take their sink.  Returns the lane's sums.
__device__ __noinline__ SoftTerm eval_soft(const NodeW *nodes, const double *momq, int imgBits, const double *s_off,
                                           double sx, double sy, double sz, double ms, double hs, int nS, unsigned it,
                                           int cnt, int lane) {
    SoftTerm sum;
    sum.ax = sum.ay = sum.az = sum.pot = sum.dt = 0.0;
    const unsigned imgMask = (1u << imgBits) - 1u;
    for (int j = 0; j < cnt; ++j) {
        const unsigned e = __shfl_sync(FULL, it, j);
        const int cn = (int)(e >> imgBits), ci = (int)(e & imgMask);
        if (lane < nS) {
            const NodeW nd = load_node(&nodes[cn]);
            const double cx = nd.rx + s_off[3 * ci], cy = nd.ry + s_off[3 * ci + 1], cz = nd.rz + s_off[3 * ci + 2];
            const SoftTerm st = softcell_on_sink(nd.fMass, nd.fSoft, &momq[(size_t)cn * 6], sx - cx, sy - cy, sz - cz, ms, hs);
            sum.ax += st.ax; sum.ay += st.ay; sum.az += st.az; sum.pot += st.pot;
            sum.dt = fmax(sum.dt, st.dt);
        }
    }
    return sum;
}

// One warp per (bucket, pass of <= 8 active sinks), streaming through the bucket's three contiguous lists.
template <int ORDER, bool MONO64>
#ifdef GG_EVAL_MAXREG
__global__ void __launch_bounds__(GG_WARPS_PER_CTA * 32) __maxnreg__(MONO64 ? 128 : GG_EVAL_MAXREG) k_eval(const TreeKernelArgs A) {
#else
__global__ void __launch_bounds__(GG_WARPS_PER_CTA * 32, MONO64 ? GG_MONO_MIN_CTAS : GG_MIN_CTAS) k_eval(const TreeKernelArgs A) {
#endif
    typedef EvalSmemT<MONO64 ? 2 : 1> EvalSmem;
    extern __shared__ __align__(16) unsigned char eval_smem_raw[];
    EvalSmem *s_w = reinterpret_cast<EvalSmem *>(eval_smem_raw);
    double *s_off = reinterpret_cast<double *>(s_w + GG_WARPS_PER_CTA);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    if (A.okFlag && !*A.okFlag) return;
    for (int i = threadIdx.x; i < A.nImages * 3; i += blockDim.x) s_off[i] = A.imgOff[i];
    __syncthreads();
    EvalSmem &W = s_w[warp];
    EvalCtx E;
    E.imgMask = (1u << A.imgBits) - 1u;

    for (;;) {
        int t = 0;
        if (lane == 0) t = A.taskBegin + atomicAdd(A.taskCounter + 2, 1);
        t = __shfl_sync(FULL, t, 0);
        if (t >= A.taskEnd) break;
        const Task task = A.tasks[t];
        const NodeW bk = load_node(&A.nodes[task.node]);
        double box[6], fSoftMax;
        int nAct;
        sink_box(A, bk, lane, box, fSoftMax, nAct);
        E.cenx = 0.5 * (box[0] + box[3]); E.ceny = 0.5 * (box[1] + box[4]); E.cenz = 0.5 * (box[2] + box[5]);

        // ---- hand the sinks of this pass (active ranks [8*pass, 8*pass+8)) to their lanes: lane = q*nS + s owns
        //      sink s; the G = 32/nS sub-groups q split every staged block between them
        const int rank0 = task.pass * GG_MAX_SINKS;
        E.nS = min(GG_MAX_SINKS, nAct - rank0);
        E.G = 32 / E.nS;
        E.q = lane / E.nS;
        const int sI = lane - E.q * E.nS;
        E.worker = E.q < E.G;
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
                    W.stage[0][rank] = make_float4((float)(p.x - E.cenx), (float)(p.y - E.ceny), (float)(p.z - E.cenz), p.m);
                    W.stage[0][GG_MAX_SINKS + rank] = make_float4(p.h, __int_as_float(pi), 0.f, 0.f);
                    if (MONO64) { W.raw[4 * rank] = p.x - E.cenx; W.raw[4 * rank + 1] = p.y - E.ceny; W.raw[4 * rank + 2] = p.z - E.cenz; }
                }
                seen += __popc(m);
            }
        }
        __syncwarp();
        Sink K;
        {
            const float4 sk = W.stage[0][sI], sk2 = W.stage[0][GG_MAX_SINKS + sI];
            K.sx = sk.x; K.sy = sk.y; K.sz = sk.z; K.ms = sk.w; K.hs = sk2.x;
            K.sidx = __float_as_int(sk2.y);
            K.sxd = K.syd = K.szd = 0.0;
            if (MONO64) { K.sxd = W.raw[4 * sI]; K.syd = W.raw[4 * sI + 1]; K.szd = W.raw[4 * sI + 2]; }
        }
        K.ax = K.ay = K.az = K.ap = K.dtm = 0.f;
        K.dax = K.day = K.daz = K.dap = 0.0;
        __syncwarp();
        const int4 nL = __ldg(reinterpret_cast<const int4 *>(&A.bucketCnt[GG_NLIST * task.ord]));
        const int nLeaf = nL.x, nSoft = nL.y, nNewt = nL.z, nBig = nL.w;
        const unsigned *L = A.lists + A.bucketOff[task.ord];

        if (MONO64) { // the massive far cells of a periodic box: monopoles in FP64
            eval_cells<ORDER, true>(A, W, s_off, E, K, L, nBig, lane);
            L += nBig;
        }
        eval_cells<ORDER, false>(A, W, s_off, E, K, L, nNewt, lane);
        L += nNewt;
        for (int i = 0; i < nSoft; i += 32) {
            const int cnt = min(32, nSoft - i);
            const unsigned e = lane < cnt ? L[i + lane] : 0u;
            const SoftTerm st = eval_soft(A.nodes, A.momq, A.imgBits, s_off, (double)K.sx + E.cenx, (double)K.sy + E.ceny,
                                          (double)K.sz + E.cenz, (double)K.ms, (double)K.hs, E.nS, e, cnt, lane);
            K.dax += st.ax; K.day += st.ay; K.daz += st.az; K.dap += st.pot;
            K.dtm = fmaxf(K.dtm, (float)st.dt);
        }
        L += nSoft;
        for (int i = 0; i < nLeaf; i += 32) {
            const int cnt = min(32, nLeaf - i);
            const unsigned e = lane < cnt ? L[i + lane] : 0u;
            eval_leaves(A, W, s_off, E, K, e, cnt, lane);
        }

        // ---- combine the G sub-groups (FP64, fixed order) and write out
        {
            double *red = reinterpret_cast<double *>(W.stage[0]); // [32][4] doubles, then [32] floats
            float *redt = reinterpret_cast<float *>(red + 128);
            red[4 * lane] = K.dax; red[4 * lane + 1] = K.day; red[4 * lane + 2] = K.daz; red[4 * lane + 3] = K.dap;
            redt[lane] = K.dtm;
            __syncwarp();
            if (lane < E.nS) {
                double vx = 0.0, vy = 0.0, vz = 0.0, vp = 0.0;
                float vd = 0.f;
                for (int g = 0; g < E.G; ++g) {
                    const int l = g * E.nS + lane;
                    vx += red[4 * l]; vy += red[4 * l + 1]; vz += red[4 * l + 2]; vp += red[4 * l + 3];
                    vd = fmaxf(vd, redt[l]);
                }
                // the Ewald correction and the comoving background term (pkd.c:2962-2991) were accumulated before this
                // kernel; the tree sum is added to them here (a + b == b + a bit for bit), so this store is final and
                // can go straight to the caller's mapped host arrays as well
                const size_t si = (size_t)K.sidx;
                vx += A.acc[3 * si]; vy += A.acc[3 * si + 1]; vz += A.acc[3 * si + 2]; vp += A.pot[si];
                A.acc[3 * si] = vx; A.acc[3 * si + 1] = vy; A.acc[3 * si + 2] = vz;
                A.pot[si] = vp;
                A.dtg[si] = (double)vd;
                if (A.hacc) {
                    A.hacc[3 * si] = vx; A.hacc[3 * si + 1] = vy; A.hacc[3 * si + 2] = vz;
                    A.hpot[si] = vp;
                    A.hdtg[si] = (double)vd;
                }
            }
        }
        __syncwarp();
    }
}

} // namespace

static size_t image_table_bytes(int nImages) { return ((size_t)nImages * 3 * sizeof(double) + 15) & ~(size_t)15; }
size_t gg_walk_kernel_smem(int nImages) { return GG_WALK_WARPS * walk_smem_bytes() + image_table_bytes(nImages); }

static int grid_for(const void *fn, int threads, size_t smem, int nSM, int warpsPerCta, int nTasks, cudaError_t *pe) {
    int perSM = 0;
    *pe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, fn, threads, smem);
    if (perSM < 1) perSM = 1;
    int grid = nSM * perSM;
    const int need = (nTasks + warpsPerCta - 1) / warpsPerCta;
    if (grid > need) grid = need > 0 ? need : 1;
    return grid;
}

cudaError_t gg_launch_walk_kernel(const TreeKernelArgs &a, int nSM, cudaStream_t st) {
    const size_t smem = gg_walk_kernel_smem(a.nImages);
    // more image roots than a frontier takes at once (nReplicas > 3): the instantiation that seeds them in batches
    void (*fn)(const TreeKernelArgs) = a.nImages > GG_SEED_IMAGES ? k_walk<true> : k_walk<false>;
    cudaError_t e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int nGroups = (a.nBuckets + GG_WALK_GB - 1) / GG_WALK_GB;
    const int grid = grid_for((const void *)fn, GG_WALK_WARPS * 32, smem, nSM, GG_WALK_WARPS, nGroups, &e);
    if (e != cudaSuccess) return e;
    fn<<<grid, GG_WALK_WARPS * 32, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t gg_launch_guard_kernel(const TreeKernelArgs &a, long long capListEntries, int *okFlag, cudaStream_t st) {
    k_guard<<<1, 1, 0, st>>>(a.errFlag, a.poolCursor, a.capBlocks, a.bucketOff, a.nBuckets, capListEntries, okFlag);
    return cudaGetLastError();
}

cudaError_t gg_launch_scatter_kernel(const TreeKernelArgs &a, int nSM, cudaStream_t st) {
    (void)nSM;
    const int nGroups = (a.nBuckets + GG_WALK_GB - 1) / GG_WALK_GB;
    if (nGroups <= 0) return cudaSuccess;
    k_scatter<<<(2 * GG_NLIST * nGroups + 7) / 8, 256, 0, st>>>(a);
    return cudaGetLastError();
}

size_t gg_eval_kernel_smem(int mono64, int nImages) {
    return GG_WARPS_PER_CTA * (mono64 ? sizeof(EvalSmemT<2>) : sizeof(EvalSmemT<1>)) + image_table_bytes(nImages);
}

cudaError_t gg_launch_eval_kernel(const TreeKernelArgs &a, int nSM, cudaStream_t st) {
    void (*fn)(const TreeKernelArgs) = nullptr;
    if (a.mono64) {
        switch (a.iOrder) {
        case 1: fn = k_eval<1, true>; break;
        case 2: fn = k_eval<2, true>; break;
        case 3: fn = k_eval<3, true>; break;
        default: fn = k_eval<4, true>; break;
        }
    } else {
        switch (a.iOrder) {
        case 1: fn = k_eval<1, false>; break;
        case 2: fn = k_eval<2, false>; break;
        case 3: fn = k_eval<3, false>; break;
        default: fn = k_eval<4, false>; break;
        }
    }
    const size_t smem = gg_eval_kernel_smem(a.mono64, a.nImages);
    cudaError_t e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = grid_for((const void *)fn, GG_WARPS_PER_CTA * 32, smem, nSM, GG_WARPS_PER_CTA, a.taskEnd - a.taskBegin, &e);
    if (e != cudaSuccess) return e;
    fn<<<grid, GG_WARPS_PER_CTA * 32, smem, st>>>(a);
    return cudaGetLastError();
}
