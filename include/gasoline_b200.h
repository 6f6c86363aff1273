/*
 * gasoline_b200.h -- C ABI of the B200-native tree-gravity force evaluation.
 *
 * This is the drop-in boundary for ONE path of Gasoline: the per-rank gravity driver
 *     pkdGravAll (pkd.c:2868, prototype pkd.h:797-801)
 *       -> pkdBucketWalk (walk.c:306, walk.h:32) -> pkdBucketInteract (grav.c:23, grav.h:100)
 *       -> pkdEwaldInit / pkdBucketEwald (ewald.c:182 / ewald.c:15, ewald.h:7-8)
 * The reference has no FFI or plugin table: pstGravity (pst.c:3310-3315) calls pkdGravAll directly.  A C host is
 * switched over by linking a replacement pkdGravAll that flattens its PKD into the plain arrays below and calls
 * gg_gravity (the shim is gasoline_b200/csrc/pkd_gravall_shim.c; INTEGRATION.md shows the link line).
 *
 * Conventions: plain pointers and sizes only; every function returns GG_OK (0) or a negative GG_ERR_* code and
 * never falls back to a CPU path -- without a usable CUDA device gg_create fails.  gg_last_error() returns a
 * human-readable message for the calling thread.  One context per rank/GPU; a context is not re-entrant (the
 * reference calls pkdGravAll once per rank per force evaluation from that rank's only thread).
 * All floating-point inputs are the host's doubles, bit for bit (FLOAT is double, floattype.h:18): the opening
 * decisions are made in FP64 with the reference's operation order so per-bucket list counts are identical.
 */
#ifndef GASOLINE_B200_H
#define GASOLINE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_OK 0
#define GG_ERR_CUDA (-1)       /* a CUDA runtime call failed (message has file:line and the CUDA error) */
#define GG_ERR_ARG (-2)        /* invalid argument / call order */
#define GG_ERR_UNSUPPORTED (-3) /* e.g. iOrder outside 1..4, bucket larger than GG_MAX_BUCKET, bDoSun with several domains */
#define GG_ERR_NOMEM (-4)

#define GG_NMOM 31       /* Q6 O10 H15 in the order of struct pkdCalcCellStruct, pkd.h:441-451 */
#define GG_NROOT 35      /* struct ilCellNewt, pkd.h:481-494: m,x,y,z,xx,yy,xy,xz,yz,zz, 10 octopole, 15 hexadecapole */
/* opening criteria of pkdCalcOpen (pkd.c:2228-2264), numbered as in opentype.h:5-9 */
#define GG_OPEN_JOSH 1
#define GG_OPEN_ABSPAR 2
#define GG_OPEN_RELPAR 3
#define GG_OPEN_ABSTOT 4
#define GG_OPEN_RELTOT 5
#define GG_MAX_BUCKET 64 /* most particles one bucket may hold (reference default nBucket=8, master.c:480) */

typedef struct gg_context gg_context;

/*
 * A k-d tree exactly as pkdBuildBinary + pkdThreadTree leave it in pkd->kdNodes (pkd.h:454-469), field by field
 * in SoA form.  iLower = first child or -1 for a bucket; iUpper = next cell of the threaded depth-first sweep
 * (pkd.c:2590-2620), -1 at the end.  pLower/pUpper index the particle arrays of the same domain.
 */
typedef struct gg_tree {
    int nNodes;
    int iRoot;            /* pkd->iRoot */
    const double *bnd;    /* [nNodes][6]  bnd.fMin[3], bnd.fMax[3] */
    const double *r;      /* [nNodes][3]  centre of mass */
    const double *fMass;  /* [nNodes] */
    const double *fSoft;  /* [nNodes]     mass-weighted softening */
    const double *fOpen2; /* [nNodes]     squared opening radius (pkdCalcOpen, pkd.c:2228) */
    const double *mom;    /* [nNodes][GG_NMOM] reduced multipoles about r (pkdCalcCell, pkd.c:2018); NULL: the device
                             forms them itself from the particles (same definition, summed bottom-up in FP64) and the
                             248 B per cell are not transferred -- r, fMass, fSoft, fOpen2 stay the host's */
    const int *pLower;    /* [nNodes] */
    const int *pUpper;    /* [nNodes] */
    const int *iLower;    /* [nNodes] */
    const int *iUpper;    /* [nNodes] */
} gg_tree;

/* pkd->pStore[0..nLocal) in tree order, the five fields the path reads (pkd.h:90-108) + the ACTIVE bit. */
typedef struct gg_particles {
    int n;
    const double *x, *y, *z; /* r[0..2] */
    const double *fMass;
    const double *fSoft;
    const int *active; /* TYPEQueryACTIVE(p) != 0 (pkd.h:349); NULL = all active */
} gg_particles;

/* The scalar arguments of pkdGravAll (pkd.h:797-801) that the GPU path honours. */
typedef struct gg_params {
    int nReps;       /* image shells for the tree walk (nReplicas) */
    int bPeriodic;
    int iOrder;      /* multipole order of the tree lists, 1..4 */
    int bEwald;
    int iEwOrder;    /* order of the Ewald root expansion, 1..4 */
    double fEwCut;   /* dEwCut  (master.c:669) */
    double fEwhCut;  /* dEwhCut (master.c:672) */
    int bComove;     /* with !bPeriodic: uniform background term, pkd.c:2967-2991 */
    double dRhoFac;
    double fPeriod[3]; /* pkd->fPeriod; >= DBL_MAX on an axis = not periodic (walk.c:326) */
    int accumulate;  /* 1: a, fPot += and dtGrav = max(old,new) like the reference; 0: overwrite */
    int flags;       /* GG_FLAG_* */
    int bDoSun;      /* pkd.c:3003-3041: after the particles, a dummy sink at the origin (softening dSunSoft, cell box
                        +-1e-14) walks the tree and is evaluated like a bucket; its acceleration comes back in
                        gg_stats.aSun.  Open boundaries, one domain (the reference asserts nReps == 0, !bPeriodic) */
    double dSunSoft;
} gg_params;

#define GG_FLAG_WALK_ONLY 1 /* build lists and count them, skip all force arithmetic (parity hook) */
#define GG_FLAG_NO_DOWNLOAD 2 /* leave results on the device (gg_device_results); host pointers may be NULL */

/* What pkdGravAll returns through its pointer arguments (+ device timings in milliseconds, CUDA events). */
typedef struct gg_stats {
    int nActive;
    double dPartSum, dCellSum, dSoftSum; /* pkd.c:2945-2949 */
    double dFlop;                        /* grav.c:246-247 + ewald.c:175-176, the reference's own scoring */
    double dFlopEwald;                   /* the Ewald share of dFlop */
    double msTree;                       /* walk + list-evaluation kernels */
    double msEwald;                      /* Ewald kernel */
    double msTotal;                      /* everything on the device between upload and download */
    int nKernelLaunches;                 /* kernels of this library launched by the call */
    int nMaxPart, nMaxCellSoft, nMaxCellNewt; /* per-bucket list maxima (the reference's diag line, pkd.c:3057) */
    double msWalk;                       /* the walk kernel's share of msTree */
    double msEval;                       /* the list-evaluation kernel's share of msTree (the dominant kernel) */
    double nListEntries;                 /* entries of the per-bucket interaction lists k_eval streamed (4 B each) */
    double aSun[3];                      /* bDoSun: the indirect acceleration at the origin (pkd.c:3037-3039), else 0 */
    int nSunPart, nSunCellSoft, nSunCellNewt; /* bDoSun: the dummy bucket's list lengths (parity hook) */
} gg_stats;

const char *gg_last_error(void);
int gg_version(void);

/* Number of CUDA devices this process sees (a multi-rank host maps rank r to device r modulo this). */
int gg_device_count(int *pn);
/* device < 0: use the current device.  Fails (GG_ERR_CUDA) when no CUDA device is usable. */
int gg_create(gg_context **pctx, int device);
void gg_destroy(gg_context *ctx);

/*
 * Ingest the caller's domain: replaces pkd->kdNodes / pkd->pStore (host memory; pinned memory from gg_host_alloc
 * copies fastest).  Must be called again whenever the host rebuilds its tree or moves particles (the reference
 * frees and re-allocates kdNodes at every build, pkd.c:2636-2642).  idSelf is this domain's rank (pkd->idSelf).
 * On return everything the tree walk needs is on the device; the transfer of tree->mom (60 % of the bytes, read only
 * by the list evaluation) may still be in flight on a second stream so that it overlaps the walk of the following
 * gg_gravity: tree->mom must stay valid and unmodified until the next call into this library returns.
 */
int gg_set_local(gg_context *ctx, int idSelf, const gg_tree *tree, const gg_particles *part);

/*
 * New ACTIVE flags (tree order, host or device pointer; NULL = all active) for the domain that is already loaded -- what
 * msrActiveRung changes between two force evaluations on the same tree (master.c:8403-8420).  Only the flags are
 * transferred and the sink-bucket task list rebuilt; a resident store (gg_state_*) takes the flags over as well.
 */
int gg_set_active(gg_context *ctx, const int *active);

/*
 * Multi-rank runs only.  The gathered top tree pkd->kdTop[1..nCell) (heap indexed, ROOT=1, pkd.h:77-86, filled by
 * pkdDistribCells pkd.c:4376): for each heap cell its pLower (-1 interior, else the rank owning the leaf), the
 * used flag (pUpper != 0), r, fMass, fSoft, fOpen2 and mom.  With one rank this call is not needed.
 */
int gg_set_top(gg_context *ctx, int nCell, const int *pLower, const int *bUsed, const double *r, const double *fMass,
               const double *fSoft, const double *fOpen2, const double *mom);

/*
 * Multi-rank runs only.  A remote domain's tree and particles (what pkdRemoteWalk reads through mdlAquire,
 * walk.c:181-304), or a pruned locally-essential subset of it with the same link semantics.  bDevice != 0: the
 * pointers are device pointers on this context's GPU (filled by an NCCL gather), else host pointers.
 */
int gg_set_remote(gg_context *ctx, int id, const gg_tree *tree, const gg_particles *part, int bDevice);
int gg_clear_remote(gg_context *ctx);

/*
 * Multi-rank runs, device-to-device form of the same exchange (what a host with NCCL uses; no host staging).
 * gg_export_size: bytes of this rank's domain in the DEVICE record layout (per node a 64 B walk record, a 128 B FP32
 * moment record and a 48 B raw quadrupole; per particle a 32 B source record).  gg_export_local: write those records,
 * section after section, to the device buffer dst (e.g. the send half of an NCCL all-gather) and wait for the copy.
 * gg_set_remote_packed: ingest a remote domain from such a buffer (device pointer, e.g. a slice of the all-gather's
 * receive buffer); links and particle indices are rebased on the fly.  hdr receives / supplies {nNodes, nPart, iRoot}.
 */
/*
 * The pruned form of the same export: the LOCALLY ESSENTIAL part of this rank's tree for each of nRemote other domains
 * whose particles lie inside bnd[r] = (fMin[3], fMax[3]) -- the root bounds every rank knows from the top tree.
 * This rank walks its own tree against each remote box, for every periodic image offset of prm, with the reference's
 * opening test (INTERSECTNP walk.h:12-30 against c.fOpen2, cells of < 4 particles always opened, walk.c:81): a cell
 * the box does not open can never be opened by a bucket inside the box (the point-to-box distance only grows, in
 * floating point too), so the remote rank's walks see exactly what they would see in the full tree.  Kept: every
 * visited cell with its moments; the particles of opened buckets; links re-indexed.  Output: one device buffer
 * (owned by the context, valid until the next call) holding nRemote domains in the gg_export_local layout; domain r
 * occupies [offsets[r], offsets[r+1]) and hdr[3r..] = {nNodes, nPart, iRoot} for gg_set_remote_packed.
 */
int gg_let_export(gg_context *ctx, int nRemote, const double *bnd, const gg_params *prm, void **pDev, size_t *offsets,
                  int *hdr);
int gg_export_size(gg_context *ctx, size_t *bytes, int hdr[3]);
int gg_export_local(gg_context *ctx, void *dst);
int gg_set_remote_packed(gg_context *ctx, int id, const int hdr[3], const void *src);

/*
 * The exchange BELOW the ABI (multi-rank hosts: one context = one rank = one GPU).  In the reference every rank's
 * pkdGravAll pulls remote cells and particles on demand through the MDL software cache (pkdRemoteWalk, walk.c:181-304:
 * mdlAquire(CID_CELL / CID_PARTICLE, ..., id)); here every rank pushes, once per force evaluation, the locally
 * essential part of its tree to every other rank and the library moves the bytes itself:
 *   gg_comm_unique_id + gg_comm_init   an NCCL communicator over the ranks' GPUs (NVLink / NVSwitch).  One rank creates
 *                                      the 128-byte id, the host distributes it by its own means (an MDL service,
 *                                      MPI_Bcast, the pthread MDL's shared memory) and every rank calls gg_comm_init --
 *                                      collectively, like ncclCommInitRank.  Ranks may be processes or threads.
 *   gg_group_create + gg_comm_init_local   the same for ranks that are threads of ONE process (pthread MDL): pointers
 *                                      are published in the group and the pieces move by peer-to-peer copies; several
 *                                      ranks may then share one GPU (how the one-GPU test box runs multi-rank hosts).
 *   gg_exchange                        COLLECTIVE, after gg_set_local / gg_build_local and before gg_gravity: forgets the
 *                                      remote domains of the previous step, prunes the local tree against every other
 *                                      rank's root bounds (gg_let_export's rule, all periodic images of prm), sends each
 *                                      rank its piece (sizes by an all-gather, trees by grouped send/receive on the
 *                                      context's stream) and ingests what arrives as remote domain r = rank r.
 *                                      bndAll[nRanks][6] = every rank's root bounds (fMin[3], fMax[3]) as the host's top
 *                                      tree knows them (pkd->kdTop leaves), or NULL: the library gathers them itself.
 *                                      The top tree itself (gg_set_top) and the Ewald root (gg_set_root_moments) are the
 *                                      host's (pkdDistribCells / pkdDistribRoot) and may be set before or after.
 *   gg_comm_allgather                  small host buffers (<= 8 KB per rank) over the same transport -- what a host
 *                                      without an MDL of its own uses to assemble the top tree.
 * Rank r must be the context whose local domain has idSelf == r.
 */
#define GG_UNIQUE_ID_BYTES 128
typedef struct gg_group gg_group;
typedef struct gg_exchange_stats {
    double msExport, msTransfer, msIngest, msTotal; /* device time of the phases (CUDA events on the context's stream) */
    double bytesSent, bytesReceived;                /* pruned trees this rank sent / received */
    double bytesWholeDomain;                        /* what sending the whole local domain to ONE rank would have moved */
    int nKernelLaunches;
} gg_exchange_stats;
int gg_comm_unique_id(void *id /* GG_UNIQUE_ID_BYTES */);
int gg_comm_init(gg_context *ctx, const void *id, int rank, int nRanks);
int gg_group_create(gg_group **pg, int nRanks);
void gg_group_destroy(gg_group *g);
int gg_comm_init_local(gg_context *ctx, gg_group *g, int rank);
int gg_comm_free(gg_context *ctx);
/* transport: 0 NCCL, 1 in-process group; ncclVersion as ncclGetVersion reports it (0 for the group transport) */
int gg_comm_info(gg_context *ctx, int *pRank, int *pnRanks, int *pTransport, int *pNcclVersion);
int gg_comm_allgather(gg_context *ctx, const void *mine, size_t bytes, void *all);
int gg_exchange(gg_context *ctx, const gg_params *prm, const double *bndAll, gg_exchange_stats *stats);

/*
 * gg_set_local in SLICES, for hosts whose tree and particles are not SoA arrays to begin with (pkd->kdNodes is an array of
 * 536-byte KDN, pkd->pStore of 184-byte PARTICLE: the shim has to flatten them): while the host flattens slice k + 1 the
 * copy engine already moves slice k, and with an announced Ewald evaluation (gg_announce) the correction of slice k runs
 * on the side stream.  Pointers are to the slice's first element (pinned host memory, valid until gg_local_end returns);
 * slices may arrive in any order but must cover [0, nPart) and [0, nNodes) exactly once.  The cells' moments are formed on
 * the device (as with gg_tree.mom == NULL).  rootBnd = (fMin[3], fMax[3]) of the root cell or NULL (see gg_exchange);
 * active == NULL in every slice = all particles are sinks.  Same result as gg_set_local, to the bit.
 */
int gg_local_begin(gg_context *ctx, int idSelf, int nNodes, int iRoot, int nPart, const double *rootBnd, int bActive);
int gg_local_particles(gg_context *ctx, int first, int count, const double *x, const double *y, const double *z,
                       const double *fMass, const double *fSoft, const int *active);
int gg_local_nodes(gg_context *ctx, int first, int count, const double *r, const double *fMass, const double *fSoft,
                   const double *fOpen2, const int *pLower, const int *pUpper, const int *iLower, const int *iUpper);
int gg_local_end(gg_context *ctx);

/* pkd->ilcnRoot (pkdCalcRoot/pkdDistribRoot, pkd.c:4395-4493): complete l<=4 moments of the whole box for Ewald. */
int gg_set_root_moments(gg_context *ctx, const double root[GG_NROOT]);

/*
 * Optional: announce the parameters of the NEXT gg_gravity before the domain goes up.  pkdGravAll (pkd.c:2868) receives
 * tree, particles and parameters in one call; a host that splits it into gg_set_local + gg_gravity can say here what is
 * coming.  With bPeriodic && bEwald (and the root expansion already set by gg_set_root_moments) the following
 * gg_set_local then uploads the particles FIRST, in slices, and launches the Ewald correction of each slice on a side
 * stream as soon as it has landed (pkdBucketEwald needs only the particles and pkd->ilcnRoot, ewald.c:15): the FP64
 * kernel runs while the copy engine is still busy with the rest of the particles and the tree.  gg_gravity with the
 * same Ewald parameters picks the finished correction up instead of launching it; any other sequence (different
 * parameters, gg_set_active, new root moments in between, a re-run) falls back to the ordinary order.  Results are
 * identical either way (same kernel, same inputs).  prm == NULL withdraws the announcement.
 */
int gg_announce(gg_context *ctx, const gg_params *prm);

/*
 * One force evaluation = pkdGravAll (pkd.c:2868).  Output arrays are indexed like the local particles (tree
 * order): a[3*i..], fPot[i], dtGrav[i] (running max of 1/dt^2, grav.c:100), fWeight[i] (flops of the particle's
 * bucket, written for active particles only, pkd.c:2851).  Inactive particles are left untouched.
 */
int gg_gravity(gg_context *ctx, const gg_params *prm, double *a, double *fPot, double *dtGrav, double *fWeight,
               gg_stats *stats);

/*
 * gg_gravity with the results handed over in pieces.  The list evaluation runs as nChunks launches over consecutive ranges
 * of sink buckets (tree order); as soon as a range has finished -- while the GPU evaluates the next one -- onChunk(user,
 * first, count) is called on the calling thread: a[3 i], fPot[i], dtGrav[i], fWeight[i] of the ACTIVE particles
 * first <= i < first + count are final.  The ranges are disjoint and cover [0, nLocal) in ascending order.  For hosts that
 * must fold the results into records of their own (pkdGravAll's in-place += on pStore, grav.c:192-195): that memory-bound
 * pass then runs beside the evaluation instead of after it.  Needs overwrite mode (prm->accumulate == 0) and output arrays
 * in mapped pinned memory (gg_host_alloc); otherwise, and for small task lists, it is gg_gravity followed by ONE callback
 * for [0, nLocal).  Results are those of gg_gravity, bit for bit (same kernels, same lists).
 */
typedef void (*gg_chunk_fn)(void *user, int first, int count);
int gg_gravity_chunked(gg_context *ctx, const gg_params *prm, double *a, double *fPot, double *dtGrav, double *fWeight,
                       gg_stats *stats, int nChunks, gg_chunk_fn onChunk, void *user);

/* After gg_gravity: per-node (nPart, nCellSoft, nCellNewt) of the local tree, -1 where the node is not a bucket
 * with an active sink -- the counters pkdBucketWalk leaves in pkd->nPart/nCellSoft/nCellNewt (walk.c:175-177). */
int gg_bucket_counts(gg_context *ctx, int *counts3);

/* Per-bucket debug seam mirroring pkdBucketWalk (walk.h:32): the lists of ONE bucket, as (global node index,
 * image) pairs for cells and (particle index, image) pairs for particles.  Returns counts in n3. */
int gg_bucket_walk(gg_context *ctx, const gg_params *prm, int iBucket, int n3[3]);

/* The two inner seams of pkdGravAll's bucket loop (pkd.c:2952-2964), for ONE bucket of the loaded local domain:
 *   gg_bucket_interact = pkdBucketWalk + pkdBucketInteract (walk.h:32, grav.h:100): the bucket's lists are built and
 *       evaluated; a[3 j], fPot[j], dtGrav[j] receive the results of the bucket's particle j (pLower + j; inactive
 *       particles get zeros), n3 (may be NULL) the list lengths;
 *   gg_bucket_ewald    = pkdBucketEwald (ewald.h:8) on the bucket's particles with the root moments of
 *       gg_set_root_moments: a, fPot of the correction alone, *pnFlop (may be NULL) its return value (ewald.c:175-176).
 * Parity / debugging hooks: every call launches whole kernels for a handful of particles. */
int gg_bucket_interact(gg_context *ctx, const gg_params *prm, int iBucket, int nMax, double *a, double *fPot,
                       double *dtGrav, int n3[3]);
int gg_bucket_ewald(gg_context *ctx, const gg_params *prm, int iBucket, int nMax, double *a, double *fPot, int *pnFlop);

/* pkdEwaldInit (ewald.c:182): the k-space table the device uses, 5 doubles per row (hx,hy,hz,hCfac,hSfac). */
int gg_ewald_table(gg_context *ctx, const gg_params *prm, double *ewt5, int nMax, int *pnEwh);

/* Device pointers to the last results (tree order): a (3 doubles per particle), fPot, dtGrav, fWeight. */
int gg_device_results(gg_context *ctx, void **a, void **fPot, void **dtGrav, void **fWeight);

/* Measurement aids for bench.py (no reference counterpart): the FP32 FMA-pipe peak of this GPU at its clocks under
 * load (dependent-FFMA microbenchmark on all SMs, TFLOP/s; kernel duration in ms) -- the denominator of the FP32
 * roofline -- and an L2 flush (writes 384 MB) to put between timed iterations. */
int gg_measure_fp32_peak(gg_context *ctx, double *pTflops, double *pMs);
/* CUDA events on the stream the library launches on: the device time of everything queued between the two calls (several
 * ABI calls, e.g. gg_set_top + gg_exchange + gg_gravity), host-induced gaps included. */
int gg_timer_start(gg_context *ctx);
int gg_timer_stop(gg_context *ctx, double *pMs);
int gg_flush_l2(gg_context *ctx);

/* Pinned host memory for the arrays handed to gg_set_local / gg_gravity. */
int gg_host_alloc(void **p, size_t bytes);
int gg_host_free(void *p);

/*
 * Host-side tree construction with the semantics of pkdBuildBinary (pkd.c:2627: midpoint split of the longest axis
 * of the squeezed box, buckets of <= nBucket, pkdCalcCell moments, OPEN_JOSH opening radius, threaded links) and
 * pkdCalcRoot.  Used by hosts that do not already own a Gasoline tree (bench.py, tests); a Gasoline host passes its
 * own kdNodes instead.  Particles are permuted in place into tree order; iOrder receives the permutation.
 */
typedef struct gg_built_tree gg_built_tree;
int gg_tree_build(int n, double *x, double *y, double *z, double *fMass, double *fSoft, int *active, int *iOrder,
                  int nBucket, double dTheta, int iOrder_mom, int nThreads, gg_built_tree **out);
/* ... with any of pkdCalcOpen's opening criteria (pkd.c:2228-2264; numbering of opentype.h:5-9).  dCrit is theta for
 * GG_OPEN_JOSH and the absolute error bound for GG_OPEN_ABSPAR (the distance at which the truncation-error estimate of
 * the order-iOrder_mom expansion falls to it, dRootBracket pkd.c:2182-2224); the other three criteria take the reference's
 * "minimal" radius Bmax.  GG_OPEN_ABSPAR returns GG_ERR_UNSUPPORTED when a cell has no extent (a one-particle bucket):
 * the reference's root finder does not terminate on such a cell. */
int gg_tree_build_open(int n, double *x, double *y, double *z, double *fMass, double *fSoft, int *active, int *iOrder,
                       int nBucket, int iOpenType, double dCrit, int iOrder_mom, int nThreads, gg_built_tree **out);
int gg_tree_view(const gg_built_tree *bt, gg_tree *view, double root[GG_NROOT]);
/* Bmax, B2..B6 of every cell (struct pkdCalcCellStruct, pkd.h:441-451; pkd.c:2080-2087): bmom[nNodes][6]. */
int gg_tree_bnumbers(const gg_built_tree *bt, double *bmom);
void gg_tree_free(gg_built_tree *bt);


/*
 * pkdBuildBinary ON THE DEVICE (reference: BuildBinary pkd.c:2437-2587, pkdUpperPart pkd.c:1106-1133,
 * pkdCombine pkd.c:1973, pkdCalcCell pkd.c:2018, pkdCalcOpen pkd.c:2228, pkdThreadTree pkd.c:2590).  Takes the rank's
 * particles in ANY order (host pointers; pinned copies fastest), builds the gravity tree on the GPU -- the same cells,
 * the same particle order inside every bucket and bit-identical r, fMass, fSoft, fOpen2, bnd as gg_tree_build / the
 * reference's host build -- forms the moments there (as with gg_tree.mom = NULL) and leaves the domain loaded exactly
 * as gg_set_local would: gg_gravity follows.  Results of gg_gravity are in TREE order; iOrder[i] (host, may be NULL)
 * receives the input index of the particle at tree position i.  root (may be NULL) receives pkd->ilcnRoot
 * (pkdCalcRoot); it is kept for Ewald either way.  Replaces gg_tree_build + gg_set_local for hosts that hand over
 * particles instead of a tree.
 */
int gg_build_local(gg_context *ctx, int idSelf, const gg_particles *part, int nBucket, double dTheta, int *iOrder,
                   int *pnNodes, double root[GG_NROOT]);
/* ... with pkdCalcOpen's criterion named (GG_OPEN_*, below): GG_OPEN_JOSH (dCrit = theta) and the three criteria whose
 * opening radius is Bmax (GG_OPEN_RELPAR, GG_OPEN_ABSTOT, GG_OPEN_RELTOT; pkd.c:2261-2264) are built on the device;
 * GG_OPEN_ABSPAR returns GG_ERR_UNSUPPORTED (gg_tree_build_open on the host has it). */
int gg_build_local_open(gg_context *ctx, int idSelf, const gg_particles *part, int nBucket, int iOpenType, double dCrit,
                        int *iOrder, int *pnNodes, double root[GG_NROOT]);
/* The construction by-products a host needs to fill the rest of its KDN records (pkd.h:454-469): split axis (-1 for a
 * bucket), split coordinate, and Bmax of pkdCalcCellStruct, per cell in the numbering of gg_tree_fetch. */
int gg_tree_fetch_build(gg_context *ctx, int *iDim, double *fSplit, double *fBmax);

/*
 * Multi-rank hosts with device-built trees: the root cell of this rank's tree (what pstColCells gathers into kdTop,
 * pkd.c:4349, and pkdCalcRoot's expansion, pkd.c:4395) and this rank's pkdCalcCell sums about the centre of an interior
 * top-tree cell (pstCalcCell, pst.c:3789: reduced moments + Bmax over ALL local particles about rcm) -- the two things
 * the top-tree assembly needs from a rank, without the tree ever leaving the device.  Any output pointer of
 * gg_domain_summary may be NULL.
 */
int gg_domain_summary(gg_context *ctx, double bnd[6], double r[3], double *fMass, double *fSoft, double *fOpen2,
                      double mom[GG_NMOM], double root[GG_NROOT]);
int gg_domain_moments_about(gg_context *ctx, const double rcm[3], double mom[GG_NMOM], double *pBmax);
/* nodes, tree levels and device time (ms, CUDA events) of the last gg_build_local */
int gg_build_info(gg_context *ctx, int *pnNodes, int *pnLevels, double *pmsBuild);
/* Copy the device-built tree (pre-order numbering, the layout of gg_tree; mom = the device-formed reduced multipoles)
 * and the particles in tree order to host arrays; any pointer may be NULL.  Checking aid and the way a host gets
 * kdNodes back. */
int gg_tree_fetch(gg_context *ctx, double *bnd, double *r, double *fMass, double *fSoft, double *fOpen2, double *mom,
                  int *pLower, int *pUpper, int *iLower, int *iUpper, double *x, double *y, double *z, double *fMass_p,
                  double *fSoft_p, int *active);

/*
 * pkd->pStore RESIDENT on the device across force evaluations, with the steps either side of the force path
 * (SURVEY 8f ranks 2, 3).  A time-stepping host keeps positions, velocities, masses, softenings, ACTIVE flags and time
 * steps in HBM and drives
 *     gg_state_load once;  per step:  gg_state_kick, gg_state_drift, gg_state_build, gg_gravity(GG_FLAG_NO_DOWNLOAD),
 *     gg_state_kick, [gg_state_gravstep];  gg_state_fetch when it wants output
 * with no per-step host<->device particle traffic.  The store is kept in TREE order (gg_state_build permutes it like
 * pkdBuildBinary permutes pStore); id[] carries the index each particle had at gg_state_load (PARTICLE.iOrder).
 *   gg_state_kick     = pkdKick (pkd.c:3780; the -DNBODY branch pkd.c:3956-3962): ACTIVE particles
 *                       v = v*dvFacOne + a*dvFacTwo; a = the device results of the last gg_gravity, or the caller's
 *                       array a[n][3] (host or device pointer, the store's current order) when not NULL
 *   gg_state_drift    = pkdDrift (pkd.c:3686-3777): ALL particles r += dDelta*v, then the reference's periodic wrap
 *                       about fCenter; fails if a particle is still outside the box (the reference asserts)
 *   gg_state_gravstep = pkdGravStep (pkd.c:4609-4623): ACTIVE particles dt = min(dt, dEta/sqrt(dtGrav)); *pdtMin = the
 *                       smallest dt of all particles (what a single-rung host steps with)
 * All three are bit-identical to the reference's functions on the same inputs (tests/test_gpu_state.py).
 */
int gg_state_load(gg_context *ctx, int n, const double *x, const double *y, const double *z, const double *vx,
                  const double *vy, const double *vz, const double *fMass, const double *fSoft, const int *active,
                  double dt0);
int gg_state_build(gg_context *ctx, int idSelf, int nBucket, double dTheta, int *pnNodes);
int gg_state_build_open(gg_context *ctx, int idSelf, int nBucket, int iOpenType, double dCrit, int *pnNodes);
int gg_state_kick(gg_context *ctx, double dvFacOne, double dvFacTwo, const double *a);
int gg_state_drift(gg_context *ctx, double dDelta, const double fCenter[3], int bPeriodic, const double fPeriod[3]);
int gg_state_gravstep(gg_context *ctx, double dEta, double *pdtMin);
int gg_state_fetch(gg_context *ctx, double *x, double *y, double *z, double *vx, double *vy, double *vz, int *id,
                   double *dt);
/*
 * Time-step selection and rungs on the resident store (multistepping hosts, msrTopStepKDK master.c:8242):
 *   gg_state_init_dt     = pkdInitDt (pkd.c:4818): ACTIVE particles dt = dDelta
 *   gg_state_accelstep   = pkdAccelStep (pkd.c:4625, gravity-only build): dt = min(dt, dEta sqrt(fSoft/|a| dAccFac)) and, with
 *                          bSqrtPhi, dEta 3.5 sqrt(dAccFac |fPot|)/(|a| dAccFac); a, fPot = the last gg_gravity's device results
 *   gg_state_dt_to_rung  = pkdDtToRung (pkd.c:4715) with pkdOneParticleDtToRung (pkd.c:4689); returns the reference's
 *                          three results (largest rung, how many particles sit on it, the ideal largest rung)
 *   gg_state_active_rung = pkdActiveRung (pkd.c:4569): ACTIVE := rung == iRung || (bGreater && rung > iRung); the flags
 *                          also become the active set of the loaded domain (like gg_set_active) when the store has not
 *                          moved since its last build, otherwise the next gg_state_build carries them
 * All bit-identical to the reference on the same inputs (tests/test_gpu_state.py, golden vectors from the compiled
 * reference).  gg_state_set_rungs / gg_state_fetch_rungs move PARTICLE.iRung (and the ACTIVE flags) in the store's order.
 */
int gg_state_init_dt(gg_context *ctx, double dDelta);
int gg_state_accelstep(gg_context *ctx, double dEta, double dVelFac, double dAccFac, int bEpsAcc, int bSqrtPhi);
int gg_state_dt_to_rung(gg_context *ctx, int iRung, double dDelta, int iMaxRung, int bAll, int *pnMaxRung,
                        int *piMaxRungIdeal, int *piMaxRungOut);
int gg_state_active_rung(gg_context *ctx, int iRung, int bGreater, int *pnActive);
int gg_state_set_rungs(gg_context *ctx, const int *rung);
int gg_state_fetch_rungs(gg_context *ctx, int *rung, int *active);

/*
 * ORB domain decomposition: the PER-RANK services the reference's pstDomainDecomp (pst.c:1854) asks of
 * every rank, on the device.  The host (Gasoline's PST, or gasoline_b200/domain.py: pst_domain_decomp for hosts that are
 * not Gasoline) runs _pstRootSplit's bisection (pst.c:959-1034) and adds the ranks' answers; every particle carries
 * the heap index of its cell of the rank tree (ROOT = 1, LOWER(i) = 2i, UPPER(i) = 2i+1, pkd.h:77-86), so all cells
 * of one level are served by one launch.
 *   gg_orb_load    the rank's particles for this decomposition (host or device pointers; x == y == z == NULL: the
 *                  resident store of gg_state_load, n = its size); fWeight NULL: 1 for every particle, as after
 *                  reading a file; every particle starts in ROOT.  n = 0 is allowed (a rank without particles)
 *   gg_orb_bounds  = pstCalcBound (pst.c:1937) leaf / pkdCalcBound: bnd[k][6] = fMin[3], fMax[3] of the rank's particles
 *                  in PST cell iCell[k] (+-FLOAT_MAXVAL when it has none), nIn[k] = how many
 *   gg_orb_weight  = pstWeight (pst.c:1405) leaf / pkdWeight (pkd.c:945): for the trial split fSplit[k] of axis iDim[k] in
 *                  cell iCell[k]: particles with r[d] < fSplit (nLow, fLow = their weight) and the others (nHigh,
 *                  fHigh).  Counts are exact; weights are summed in a fixed order (reproducible run to run)
 *   gg_orb_split   the split is final: particles of iCell[k] move to LOWER (r[d] < fSplit[k]) or UPPER
 *   gg_orb_fetch   iCellOfParticle[n] (host or device pointer): the PST cell of every particle, in gg_orb_load order;
 *                  after the last level these are the leaves = the ranks the particles go to
 */
int gg_orb_load(gg_context *ctx, int n, const double *x, const double *y, const double *z, const double *fWeight);
int gg_orb_bounds(gg_context *ctx, int nCells, const int *iCell, double *bnd, int *nIn);
int gg_orb_weight(gg_context *ctx, int nCells, const int *iCell, const int *iDim, const double *fSplit, int *nLow,
                  int *nHigh, double *fLow, double *fHigh);
int gg_orb_split(gg_context *ctx, int nCells, const int *iCell, const int *iDim, const double *fSplit);
/* The split with a second boundary per cell -- what pkdColRejects leaves when a side's particles do not fit its ranks'
 * stores (pst.c:1049-1270, pkd.c:1463-1485): the lower child takes the WRAPPED interval between fSplitInactive and fSplit
 * (pkdLowerPartWrap, pkd.c:1165-1211: r < fSplit or r >= fSplitInactive when fSplitInactive > fSplit, else
 * fSplitInactive <= r < fSplit), the upper child the rest. */
int gg_orb_split_wrap(gg_context *ctx, int nCells, const int *iCell, const int *iDim, const double *fSplit,
                      const double *fSplitInactive);
int gg_orb_fetch(gg_context *ctx, int *iCellOfParticle);
/*
 * _pstRootSplit's root finder (pst.c:959-1034) for ALL cells of one level of the rank tree with its state on the device:
 * what a host that owns all particles of the decomposition in one context (one GPU, or the first levels of a larger
 * run) does instead of up to 64 gg_orb_weight round trips.  Per cell k: split axis iDim[k], bracket [fLow[k], fUp[k]]
 * (the cell's bounds along that axis), bLive[k] = the bisection runs for this cell (bDoRootFind, or the previous split
 * has left the bounds), nLower[k] / nUpper[k] = ranks below / above the split, bSplitWork (master.c:964).  Every trial
 * is one weighing launch (the cells still bisected) and a one-block decide kernel; the launches are queued without a
 * host synchronisation and return at their first instruction once no cell is live.  Out: fSplit[k] (where bHasSplit[k];
 * a cell that never got a trial keeps the split the host had for it), ittr[k] as the reference counts them.  Branches,
 * midpoints and stopping rules are the reference's, so the splits are those of the host-driven loop bit for bit.
 */
int gg_orb_bisect(gg_context *ctx, int nCells, const int *iCell, const int *iDim, const double *fLow, const double *fUp,
                  const int *bLive, const double *nLower, const double *nUpper, int bSplitWork, double *fSplit,
                  int *bHasSplit, int *ittr);
/* COLLECTIVE form for ranks that each hold part of the particles (a communicator from gg_comm_init / gg_comm_init_local):
 * per trial every rank weighs its own particles, the ranks' answers are all-gathered between the devices (in-stream with
 * NCCL: no host round trip per trial) and added in rank order on every rank -- the same bits, branch and split everywhere,
 * as pst.c:1004-1030 adds the answers of the lower and upper sub-trees.  Every rank passes the same arguments. */
int gg_orb_bisect_all(gg_context *ctx, int nCells, const int *iCell, const int *iDim, const double *fLow, const double *fUp,
                  const int *bLive, const double *nLower, const double *nUpper, int bSplitWork, double *fSplit,
                  int *bHasSplit, int *ittr);

/* Every cell's reduced multipoles by the algorithm the DEVICE uses when gg_tree.mom is NULL (raw moments of the
 * buckets, children translated to the parent's centre and summed, then reduced as pkdCalcCell defines them), executed
 * on the host: mom[nNodes][GG_NMOM].  A checking aid for hosts and tests; no GPU needed. */
int gg_tree_moments_m2m(const gg_tree *tree, const gg_particles *part, double *mom);

/* pkdCalcCell (pkd.c:2018-2135) over all n particles of a domain about the centre rcm: reduced multipoles in GG_NMOM
 * order and Bmax -- one rank's contribution to an interior cell of the top tree, which pstCalcCell (pst.c:3789) sums
 * over the ranks below that cell.  Host code (no GPU needed); used by multi-GPU hosts that are not Gasoline. */
int gg_cell_moments(int n, const double *x, const double *y, const double *z, const double *fMass,
                    const double rcm[3], int iOrder, double mom[GG_NMOM], double *pBmax);

#ifdef __cplusplus
}
#endif
#endif
