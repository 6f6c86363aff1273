// gg_context.h -- the per-rank context behind the C ABI and the helpers shared by gg_api.cu and gg_comm.cu.
#pragma once
#include <string>
#include <vector>
#include "gg_internal.h"

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Domain {
    int id, nNodes, nPart, iRoot, nodeBase, partBase;
};

struct GGComm; // gg_comm.cu: the exchange backend of a multi-rank run (NCCL communicator or in-process group)

struct gg_context {
    int device = 0, nSM = 0;
    cudaStream_t st = nullptr;
    cudaStream_t st2 = nullptr; // the local domain's moments travel here while the walk already runs on st
    cudaEvent_t evMom = nullptr;
    cudaStream_t st3 = nullptr; // k_stats (bookkeeping + fWeight) runs here, beside the list scatter / evaluation
    cudaEvent_t evWalk = nullptr, evStats = nullptr, evPacked = nullptr;
    double *zc[4] = {nullptr, nullptr, nullptr, nullptr}; // device aliases of the caller's mapped a, fPot, dtGrav, fWeight
    double *zcHost[4] = {nullptr, nullptr, nullptr, nullptr};
    bool momPending = false;    // st2 work (moment upload + k_pack_mom) not yet known to be complete
    cudaEvent_t ev[8];
    // layout
    int idSelf = 0;
    std::vector<Domain> dom; // dom[0] = local
    int nNodesAll = 0, nPartAll = 0, maxBucket = 1;
    bool haveRoot = false;
    double root[GG_NROOT];
    // top tree (host copy, packed at gravity time)
    int nTop = 0;
    std::vector<int> topLower, topUsed;
    std::vector<double> topR, topMass, topSoft, topOpen2, topMom;
    std::vector<int> hActive; // host copy of the local ACTIVE flags (empty = all active)
    // device buffers
    DevBuf nodes, momf, momq, parts, active, hsoft, tasks, ngroups, goffs, counts, acc, pot, dtg, fweight, nloop, sums,
        misc, imgoff, ewt, raw, rawi, cubtmp, flush, pool, nextblk, poolmask, isb, boffs, bnode, ghead, gcnt, bcnt, btot, boff64, lists, letflag, letfront, letidx, letout, letmisc, momraw, mparent, dbgtask, momout;
    void *pinned = nullptr;
    size_t pinnedCap = 0;
    int nTasks = 0;
    int nTasksLocal = 0, nBucketsLocal = 0, nPartUpload = 0; // the task list gg_set_local built
    int nLaunches = 0;
    size_t capBlocks = 0; // list pool capacity (blocks of 32 references), kept at the high-water mark
    void *builder = nullptr; // gg_tree_gpu.cu workspace (gg_build_local)
    GGBuiltDev built{};      // the last device-built tree (all zero: none)
    bool rootLazy = false;   // the Ewald root expansion is to be read from the device-formed moments when first needed
    double msBuild = 0.0;
    // device-resident particle store (gg_state_*): positions / mass / softening / ACTIVE in tree order after every
    // gg_state_build, velocities SoA [3][n], persistent particle id, time step
    DevBuf sx, sy, sz, sm, sh, sact, svel, sid, sdt, svel2, sid2, sdt2, sacc, srhist;
    int stateN = 0;
    bool stateHasActive = false, stateDirty = true, stateForces = false;
    bool sunMode = false; // run_gravity is evaluating the bDoSun dummy bucket: the particles' results stay as they are
    // ORB domain decomposition services (gg_orb_*): the rank's particles for the decomposition and their PST cell
    DevBuf ox, oy, oz, ow, ocell, okeys, ocnt, opart, osums, obis, oans;
    int orbN = -1;          // -1: gg_orb_load not called
    bool orbState = false;  // positions are the resident store's (sx, sy, sz)
    bool orbWeights = false;
    // multi-rank exchange below the C ABI (gg_comm.cu)
    GGComm *comm = nullptr;
    double rootBnd[6] = {0, 0, 0, 0, 0, 0}; // bounds of the local root cell (the remote ranks prune their trees against it)
    bool haveRootBnd = false;
    DevBuf letrecv, commscratch;
    cudaEvent_t evx[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evt[2] = {nullptr, nullptr}; // gg_timer_start / gg_timer_stop
    // gg_announce: parameters of the next gg_gravity; the Ewald correction launched early by gg_set_local (side stream st4)
    bool annValid = false;
    gg_params ann;
    cudaStream_t st4 = nullptr;
    std::vector<cudaEvent_t> evSlice; // "slice k of the particles is packed"
    cudaEvent_t evEw[3] = {nullptr, nullptr, nullptr}; // early Ewald: start, end (timed), done (ordering)
    bool ewPending = false; // st4 work in flight or finished but not yet ordered before c->st
    bool ewValid = false;   // acc / pot / nloop hold the Ewald correction of the loaded domain for ewPrm / ewRoot
    gg_params ewPrm;
    double ewRoot[GG_NROOT];
    int ewNEwh = 0, ewN = 0;
    // gg_gravity_chunked: the list evaluation in nChunk launches over consecutive task ranges, the caller told as each lands
    int nChunk = 1;
    void (*chunkFn)(void *, int, int) = nullptr;
    void *chunkUser = nullptr;
    std::vector<cudaEvent_t> evChunk;
    DevBuf chunkb;
    // gg_local_begin .. gg_local_end: the local domain arriving in slices
    struct Sliced {
        bool open = false, early = false, active = false;
        int nn = 0, np = 0, iRoot = 0, idSelf = 0, gotP = 0, gotN = 0, nEv = 0;
        double *dr = nullptr, *dM = nullptr, *dS = nullptr, *dO = nullptr, *dx = nullptr, *dy = nullptr, *dz = nullptr,
               *dm = nullptr, *dh = nullptr;
        int *di = nullptr;
        EwaldKernelArgs ea;
    } sl;
};

// error reporting: formats into the calling thread's message buffer (gg_last_error), prints it, returns code
int gg_fail(int code, const char *fmt, ...);

#define CK(call)                                                                                            \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return gg_fail(GG_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

// grow a device buffer to >= bytes (keeping the first `preserve` bytes); drains the context's streams before freeing
int gg_ensure(gg_context *c, DevBuf &b, size_t bytes, size_t preserve = 0);
// wait for the asynchronous half of the last gg_set_local (moments on st2)
int gg_finish_mom(gg_context *c);

// The pruned (locally essential) copies of the local domain for nRemote remote boxes: see gg_let_export in the header.
// nOut / nOutP receive the kept nodes / particles per remote; the records are in c->letout at offsets[].
int gg_let_export_impl(gg_context *c, int nRemote, const double *bnd, const gg_params *prm, size_t *offsets, int *hdr);
// ingest one remote domain from device records at src (stream-ordered on c->st; no host synchronisation)
int gg_ingest_packed(gg_context *c, int id, const int hdr[3], const void *src);
void gg_comm_release(gg_context *c);
// ranks of the context's communicator (1 without one), and a small all-gather between DEVICE buffers: rank r's `bytes` at
// sendDev land at recvDev + r * bytes on every rank -- stream-ordered on c->st with NCCL (no host synchronisation), through
// the meeting point with the in-process group
int gg_comm_ranks(const gg_context *c);
void gg_comm_abort(gg_context *c); // a rank leaves a collective early (error): in-process peers stop waiting for it
int gg_comm_allgather_dev(gg_context *c, const void *sendDev, void *recvDev, size_t bytes);
// start the Ewald correction of the resident local domain on the side stream (no-op unless prm asks for one; gg_api.cu)
int gg_early_ewald(gg_context *c, const gg_params *prm);
