// gg_internal.h -- device data layout and kernel launch interfaces shared by the .cu files.
//
// HBM layout (see DESIGN.md "Data layout"):
//   NodeW  nodes[nNodesAll]   64 B  walk record: FP64 centre of mass, fOpen2, fSoft, fMass + 4 ints (children,
//                                   first particle, particle count).  One record = two 32 B sectors; every
//                                   opening decision reads exactly one.
//   float4 momf[nNodesAll][8] 128 B evaluation record: FP32 traceless quadrupole, reduced octopole and
//                                   hexadecapole (31 values + pad) = one cache line per accepted cell.
//   double momq[nNodesAll][6] 48 B  raw FP64 quadrupole, read only on the (rare) softened-cell path.
//   PartS  parts[nPartAll]    32 B  source record: FP64 position, FP32 mass and softening.
// Indices are GLOBAL: the local domain first, then every remote domain, then the top-tree cells.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gasoline_b200.h"

#ifndef GG_WARPS_PER_CTA
#define GG_WARPS_PER_CTA 4   // k_eval
#endif
#ifndef GG_MIN_CTAS
#define GG_MIN_CTAS 5        // resident CTAs per SM k_eval is compiled for (44 KB shared memory each; register cap 102)
#endif
#ifndef GG_MONO_MIN_CTAS
#define GG_MONO_MIN_CTAS GG_MIN_CTAS       // k_eval<.,true> (periodic boxes): the FP64 monopoles are confined to the big-cell loop
#endif
#ifndef GG_CELL_UNROLL
#define GG_CELL_UNROLL 1     // unroll factor of k_eval's (sink, cell) loop
#endif
#ifndef GG_EVAL_BSG
#define GG_EVAL_BSG 1       // k_eval stages blocks of the largest multiple of G cells (<= 32) instead of always 32
#endif
#ifndef GG_GATHER_PPL
#define GG_GATHER_PPL 1     // consecutive 16 B pieces of a 128 B moment record one lane copies in k_eval's gather (1: 8 lanes per record)
#endif
#define GG_WALK_WARPS 8      // k_walk
#define GG_WALK_MIN_CTAS 4   // 32 warps per SM (register cap 64)
#define GG_SLAB_BLOCKS 32    // list blocks a warp takes from the pool per atomic
#ifndef GG_WALK_GB
#define GG_WALK_GB 10        // sink buckets that share one tree traversal in k_walk (<= 16: one mask bit per bucket)
#endif
#if GG_WALK_GB <= 8
typedef unsigned char gg_mask_t;  // bucket mask of a frontier item / of a masked list entry
#else
typedef unsigned short gg_mask_t;
#endif
#define GG_MAX_SINKS 8      // sinks evaluated per warp pass (accumulators live in registers)
#define GG_NLIST 4          // list types: 0 leaves (opened buckets), 1 softened cells, 2 Newtonian cells, 3 big Newtonian cells
#ifndef GG_BIG_FRAC
#define GG_BIG_FRAC (1.0 / 512.0) // periodic boxes: cells holding >= this share of the whole mass get FP64 monopoles (eval_cells).
                                  // Measured on the 256^3 box, theta 0.5 (profiles/r02_big_frac.md): 1/64 -> k_eval 104.5 ms but
                                  // potential rms error 5e-6; 1/512 -> 108.1 ms, 1.5e-6; all cells -> 122 ms, 3e-7
#endif
#define GG_STACK_CAP 512    // walk frontier entries per warp
#define GG_STACK_DFS_MARGIN 128
#define GG_MAX_IMAGES 1331   // (2 nReps + 1)^3 images of the tree walk, nReps <= 5 (11 image bits in a list reference)
#define GG_SEED_IMAGES 343   // image roots put on a warp's walk frontier at once (all of them up to nReps = 3)
#define GG_MAX_TOP 128       // top-tree cells (2 x ranks, heap indexed) = spare node records behind the domains' nodes

// The FP32 evaluation record of one cell from the reference's reduced multipoles q[GG_NMOM] (pkdCalcCell order): the
// quadrupole made traceless like SETILIST (walk.c:41-48), and each order pre-multiplied by (2l-1)!! -- 3, 15, 105 -- so
// that k_eval needs only the bare powers 1/r^(2l+1) instead of gam[l] = (2l-1)!!/r^(2l+1) (grav.c:172-191).
#ifdef __CUDACC__
__host__ __device__
#endif
inline void gg_pack_momf(const double *q, float *f) {
    const double tr = (q[0] + q[1] + q[2]) / 3.0;
    f[0] = (float)(3.0 * (q[0] - tr)); f[1] = (float)(3.0 * (q[1] - tr)); f[2] = (float)(3.0 * (q[2] - tr));
    f[3] = (float)(3.0 * q[3]); f[4] = (float)(3.0 * q[4]); f[5] = (float)(3.0 * q[5]);
    for (int k = 6; k < 16; ++k) f[k] = (float)(15.0 * q[k]);
    for (int k = 16; k < 31; ++k) f[k] = (float)(105.0 * q[k]);
    f[31] = 0.f;
}

struct __align__(16) NodeW {
    double rx, ry, rz;
    double fMass;  // first 32 B (one sector) = everything k_eval needs of an accepted cell
    double fOpen2;
    double fSoft;
    int c0, c1;   // children (global node index), c0 == -1: bucket
    int pLower;   // first particle (global particle index)
    int nP;       // particle count; >= 2^30 for top-tree cells (exempt from the "< 4 particles" rule, walk.c:81)
};
static_assert(sizeof(NodeW) == 64, "NodeW must be 64 bytes");

struct __align__(16) PartS {
    double x, y, z;
    float m, h;
};
static_assert(sizeof(PartS) == 32, "PartS must be 32 bytes");

// One unit of work for a k_eval warp: a sink bucket (local node index), which pass of 8 active sinks to evaluate, and
// the bucket's ordinal among the sink buckets (tree order); ordinal / GG_WALK_GB is its walk group.
struct Task {
    int node;
    int pass;
    int ord;
    int pad;
};

struct TreeKernelArgs {
    const NodeW *nodes;
    const float4 *momf;
    const double *momq;
    const PartS *parts;
    const int *active;       // local particles, may be null (all active)
    const double *hsoft;     // local particles: FP64 softening (fSoftMax must be exact, walk.c:319-324)
    const Task *tasks;
    int nTasks;
    int *taskCounter;        // [0] k_walk's, [2] k_eval's ([1] is errFlag)
    const int *bucketNode;   // [nBuckets] sink buckets (local node index) in tree order
    int nBuckets;
    // interaction lists: 128 B blocks of 32 references chained per bucket / per walk group (see gg_tree_kernel.cu)
    unsigned *pool;          // [capBlocks][32]
    int *nextBlk;            // [capBlocks]
    int capBlocks;
    int *poolCursor;         // blocks handed out (may exceed capBlocks: the host then grows the pool and reruns)
    gg_mask_t *poolMask;     // [capBlocks][32] masked chains: which buckets of the walk group the entry belongs to
    int *groupHead;          // [nWalkGroups][GG_NLIST list types][2]: chain shared by every bucket of the group, masked chain
    int *groupCnt;           // [nWalkGroups][GG_NLIST][2] entries in each chain
    // per-bucket contiguous lists (k_scatter output, k_eval input), indexed by bucket ordinal
    int *bucketCnt;          // [nBuckets][GG_NLIST] entries of the bucket's leaf / softened-cell / Newtonian / big Newtonian list
    long long *bucketTot;    // [nBuckets + 1] k_walk: entries of the bucket over the three lists
    const long long *bucketOff; // [nBuckets + 1] exclusive scan of bucketTot: where the bucket's lists start
    unsigned *lists;         // [bucketOff[nBuckets]]: per bucket its Newtonian cells, softened cells, leaves
    int rootNode;            // global index where every image's walk starts
    int nImages, homeImage, imgBits;
    const double *imgOff;    // [nImages][3]
    int iOrder;
    int maxBucket;           // largest particle count of any bucket (sizes the particle buffer)
    int walkOnly;
    int mono64;              // k_eval: the big-cell list exists and its monopoles are evaluated in FP64 (periodic boxes)
    double bigMass;          // k_walk: a Newtonian cell is "big" when fMass >= bigMass (GG_BIG_FRAC x the mass of the walk's root
                             // cell in periodic boxes, DBL_MAX otherwise: never)
    int sunNode;             // bDoSun pass (pkd.c:3003-3041): the dummy sink bucket's node, whose box is +-sunBox; else -1
    double sunBox;
    // outputs (local particles, tree order)
    double *acc;             // [n][3]
    double *pot;
    double *dtg;
    int *counts;             // [nLocalNodes][3]
    int *errFlag;
    int taskBegin, taskEnd;  // k_eval: this launch evaluates tasks [taskBegin, taskEnd) (the whole list unless gg_gravity_chunked)
    const int *okFlag;       // non-null: k_scatter / k_eval run only if *okFlag != 0 (k_guard: the walk's output fits the
                             // buffers they were launched with -- no host round trip between the walk and the evaluation)
    // zero-copy result delivery: the caller's a / fPot / dtGrav arrays when they are mapped pinned host memory (device
    // aliases, else null) -- k_eval stores every finished sink there as well, so the download overlaps the evaluation
    double *hacc, *hpot, *hdtg;
};

struct EwaldKernelArgs {
    const PartS *parts;      // local particles
    const int *active;
    int first, n;            // particles [first, n)
    double root[GG_NROOT];
    double trQ4[7];          // Qxx,Qxy,Qxz,Qyy,Qyz,Qzz, Qtr of the hexadecapole traces (meval.h:36-42)
    double trQ3[3];          // Qx,Qy,Qz (meval.h:55-57)
    double trQ2;             // 0.5*(xx+yy+zz) (meval.h:68)
    const double *ewt;       // [nEwh][5]
    int nEwh;
    int nReps, nEwReps, iOrder;
    double L, fEwCut2, alpha, alpha2, k1, ka;
    double *acc, *pot;       // accumulated into
    int *nLoop;              // per particle: real-space terms evaluated
};

struct StatsKernelArgs {
    const NodeW *nodes;      // local nodes
    const Task *tasks;
    int nTasks;
    const int *active;
    const int *counts;
    const int *nLoop;        // may be null (no Ewald)
    int nEwh, iOrder, iEwOrder;
    double *fWeight;
    double *hfWeight;        // the caller's fWeight array (mapped pinned host memory) or null
    unsigned long long *sums; // [0] nActive [1] part [2] cell [3] soft [4] flopI [5] flopE [6..8] max lists
};

cudaError_t gg_launch_walk_kernel(const TreeKernelArgs &a, int nSM, cudaStream_t st);
cudaError_t gg_launch_guard_kernel(const TreeKernelArgs &a, long long capListEntries, int *okFlag, cudaStream_t st);
cudaError_t gg_launch_scatter_kernel(const TreeKernelArgs &a, int nSM, cudaStream_t st);
cudaError_t gg_launch_eval_kernel(const TreeKernelArgs &a, int nSM, cudaStream_t st);
cudaError_t gg_launch_ewald_kernel(const EwaldKernelArgs &a, cudaStream_t st);
cudaError_t gg_launch_stats_kernel(const StatsKernelArgs &a, cudaStream_t st);
// gg_moments.cu: multipole moments of a domain's cells from its particles, on the device (two kernels on st)
cudaError_t gg_launch_device_moments(int nn, const NodeW *nodes, int nodeBase, int partBase, int iRootLocal,
                                     const double *x, const double *y, const double *z, const double *m, int *parent,
                                     int *arrive, double *raw, float4 *momf, double *momq, cudaStream_t st);

// gg_tree_gpu.cu: the gravity tree built on the device.  The arrays are device pointers into the builder's workspace
// (valid until the next build): the tree in the reference's pre-order numbering, SoA like gg_tree, and the particles
// permuted into tree order; iorder[i] = input index of the particle now at position i.
struct GGBuiltDev {
    int nNodes, nPart, nLevels;
    const double *bnd, *r, *fMass, *fSoft, *fOpen2;
    const int *pLower, *pUpper, *iLower, *iUpper;
    const int *iDim;                // split axis, -1 for a bucket (KDN.iDim)
    const double *fSplit, *fBmax;   // split coordinate (KDN.fSplit), Bmax (pkdCalcCellStruct.Bmax)
    const double *x, *y, *z, *m, *h;
    const int *active; // null: all active
    const int *iorder;
};
int gg_builder_run(void **pBuilder, const gg_particles *pp, int nBucket, double dTheta, double c23, cudaStream_t st, GGBuiltDev *out,
                   int *pnLaunches, char *err, size_t errLen);
void gg_builder_free(void *builder);

// gg_state.cu: kick / drift / grav-step on the device-resident particle store
cudaError_t gg_launch_kick(int n, double *v, const double *a, const int *active, double f1, double f2, cudaStream_t st);
cudaError_t gg_launch_drift(int n, double *x, double *y, double *z, const double *v, double dDelta, const double c[3],
                            int bPeriodic, const double L[3], int *nOutside, cudaStream_t st);
cudaError_t gg_launch_gravstep(int n, double *dt, const double *dtGrav, const int *active, double dEta,
                               unsigned long long *dtMinBits, int *nBad, cudaStream_t st);
cudaError_t gg_launch_permute(int n, const int *iorder, const double *vIn, double *vOut, const int *idIn, int *idOut,
                              const double *dtIn, double *dtOut, cudaStream_t st);
cudaError_t gg_launch_state_init(int n, int *id, double *dt, double dt0, cudaStream_t st);
cudaError_t gg_launch_bmax_about(int n, const PartS *parts, const double c[3], unsigned long long *out, cudaStream_t st);
cudaError_t gg_launch_mom_reduce(int nn, const double *raw, double *mom, cudaStream_t st);
cudaError_t gg_launch_init_dt(int n, double *dt, const int *active, double dDelta, cudaStream_t st);
cudaError_t gg_launch_accelstep(int n, double *dt, const double *a, const double *pot, const double *fSoft, const int *active,
                                double dEta, double dAccFac, int bEpsAcc, int bSqrtPhi, cudaStream_t st);
cudaError_t gg_launch_dt_to_rung(int n, int *idr, const double *dt, int iRung, double dDelta, int iMaxRung, int bAll, int *hist,
                                 int *ideal, int *nBad, cudaStream_t st);
cudaError_t gg_launch_active_rung(int n, const int *idr, int *active, int iRung, int bGreater, int *count, cudaStream_t st);

// gg_orb.cu: the per-rank services of the ORB domain decomposition (pstCalcBound, pstWeight, the split's outcome) on
// the particles of one rank, all cells of one level of the rank tree (PST) per launch
#define GG_ORB_MAX_CELL 256  // PST heap indices (ROOT = 1): 2^(1 + ceil(log2 nThreads)) <= 256 for up to 128 ranks
#define GG_ORB_MAX_SLOTS 64  // PST cells asked about in one call
struct OrbQuery {
    int nSlots;
    int cell[GG_ORB_MAX_SLOTS]; // heap index of the PST cell
    int dim[GG_ORB_MAX_SLOTS];  // split axis
    double split[GG_ORB_MAX_SLOTS];
};
struct OrbWrap {
    double inactive[GG_ORB_MAX_SLOTS]; // fSplitInactive of every cell of the query (gg_orb_split_wrap)
};
// The root finder of one level of the rank tree with its state on the device (gg_orb_bisect): _pstRootSplit's bisection
// (pst.c:959-1034) for all cells of the level at once, no host round trip per trial.
struct OrbBisect {
    OrbQuery q;                       // the level's cells; q.split = the trial split (fm) of every cell
    double fl[GG_ORB_MAX_SLOTS], fu[GG_ORB_MAX_SLOTS], fmm[GG_ORB_MAX_SLOTS]; // bracket and its midpoint
    double nLower[GG_ORB_MAX_SLOTS], nUpper[GG_ORB_MAX_SLOTS];                // ranks below / above the split
    int ittr[GG_ORB_MAX_SLOTS], live[GG_ORB_MAX_SLOTS];
    int nLive, splitWork, maxIttr, hasSplit[GG_ORB_MAX_SLOTS];
};
// queue the whole bisection on st: B (device) initialised from the host copy h; cnt/part/sums as for gg_launch_orb_weight
cudaError_t gg_launch_orb_bisect(OrbBisect *B, const OrbBisect &h, int n, const double *x, const double *y, const double *z,
                                 const double *w, const int *cellOf, int *cnt, double *part, double *sums, cudaStream_t st);
// ... in pieces, for gg_orb_bisect_all (a collective between a trial's weighing and its decision).  One rank's answer to a
// trial is a record of GG_ORB_REC_BYTES: sums[MAX_SLOTS][2] doubles, then cnt[MAX_SLOTS][2] ints at GG_ORB_REC_CNT.
#define GG_ORB_CHUNK 16 // trials queued between two looks at the number of live cells (gg_orb_bisect_all)
#define GG_ORB_REC_CNT (2 * GG_ORB_MAX_SLOTS * 8)
#define GG_ORB_REC_BYTES (GG_ORB_REC_CNT + 2 * GG_ORB_MAX_SLOTS * 4)
cudaError_t gg_launch_orb_bisect_begin(OrbBisect *B, const OrbBisect &h, int *cnt, double *sums, int useW, cudaStream_t st);
cudaError_t gg_launch_orb_trial(OrbBisect *B, int nSlots, int n, const double *x, const double *y, const double *z,
                                const double *w, const int *cellOf, int *cnt, double *part, double *sums, cudaStream_t st);
cudaError_t gg_launch_orb_decide(OrbBisect *B, int *cnt, double *sums, int useW, int nRanks, const unsigned char *all,
                                 cudaStream_t st);
cudaError_t gg_launch_orb_init(int n, int *cellOf, cudaStream_t st);
cudaError_t gg_launch_orb_bounds(const OrbQuery &q, int n, const double *x, const double *y, const double *z, const int *cellOf,
                                 unsigned long long *out, int *cnt, cudaStream_t st);
cudaError_t gg_launch_orb_weight(const OrbQuery &q, int n, const double *x, const double *y, const double *z, const double *w,
                                 const int *cellOf, int *cnt, double *part, double *sums, cudaStream_t st);
cudaError_t gg_launch_orb_split(const OrbQuery &q, int n, const double *x, const double *y, const double *z, int *cellOf,
                                cudaStream_t st);
cudaError_t gg_launch_orb_split_wrap(const OrbQuery &q, const OrbWrap &w, int n, const double *x, const double *y, const double *z,
                                     int *cellOf, cudaStream_t st);
size_t gg_orb_part_bytes(int n);
