// gg_comm.cu -- the multi-rank exchange BELOW the C ABI: what pkdRemoteWalk pulls cell by cell through the MDL software
// cache in the reference (walk.c:181-304; mdlAquire(CID_CELL / CID_PARTICLE)) is pushed here, once per force evaluation,
// as pruned locally-essential trees: every rank exports the part of its tree each other domain can reach
// (gg_let_export_impl), the pieces travel device to device, and each rank ingests what it received behind its own
// domain.  Two transports with the same semantics:
//   * NCCL (one GPU per rank; ranks may be processes or threads): sizes by ncclAllGather, trees by grouped
//     ncclSend/ncclRecv on the context's stream, over NVLink / NVSwitch.  NCCL is bound at run time (dlopen) so the
//     library has no link-time dependency and shares the NCCL a host process already loaded.
//   * an in-process group (ranks = threads of one process, e.g. a pthread MDL; several ranks may share one GPU):
//     pointers are published in a shared table and every rank copies its pieces with cudaMemcpyPeerAsync.
#include <dlfcn.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <condition_variable>
#include <mutex>
#include "gg_context.h"

#define GG_MAX_RANKS 16
#define GG_SMALL_BYTES 8192

struct gg_group {
    int n = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long gen = 0;
    bool failed = false;
    // what every rank publishes for the current collective
    unsigned char small[GG_MAX_RANKS][GG_SMALL_BYTES];
    const char *sendBase[GG_MAX_RANKS];
    size_t sendOff[GG_MAX_RANKS][GG_MAX_RANKS], sendBytes[GG_MAX_RANKS][GG_MAX_RANKS];
    int dev[GG_MAX_RANKS];
    // false: a rank of the group gave up (fail()) -- the collective is void and nobody waits for the missing rank
    bool barrier() {
        std::unique_lock<std::mutex> lk(m);
        if (failed) return false;
        const unsigned long g = gen;
        if (++arrived == n) {
            arrived = 0;
            ++gen;
            cv.notify_all();
        } else cv.wait(lk, [&] { return gen != g || failed; });
        return !failed;
    }
    // a rank that leaves a collective early (an error between two barriers) releases the ranks waiting for it; the group
    // stays failed: its ranks are out of step and a new group has to be made
    void fail() {
        std::lock_guard<std::mutex> lk(m);
        failed = true;
        cv.notify_all();
    }
};

namespace {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

std::mutex g_ncclMutex;
NcclApi g_nccl;

// libnccl.so.2: the copy the process already holds (a PyTorch host brings its own) or the system's
int load_nccl() {
    std::lock_guard<std::mutex> lk(g_ncclMutex);
    if (g_nccl.h) return GG_OK;
    const char *env = getenv("GG_NCCL_LIB");
    void *h = nullptr;
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return gg_fail(GG_ERR_UNSUPPORTED, "gg_comm: libnccl.so.2 cannot be loaded (%s); set GG_NCCL_LIB", dlerror());
    NcclApi a;
    a.h = h;
#define SYM(field, name)                                                                         \
    *(void **)(&a.field) = dlsym(h, name);                                                       \
    if (!a.field) return gg_fail(GG_ERR_UNSUPPORTED, "gg_comm: %s missing from libnccl", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
    SYM(GetVersion, "ncclGetVersion");
#undef SYM
    g_nccl = a;
    return GG_OK;
}

#define NCK(call)                                                                                                \
    do {                                                                                                         \
        ncclResult_t r_ = (call);                                                                                \
        if (r_ != ncclSuccess)                                                                                   \
            return gg_fail(GG_ERR_CUDA, "%s:%d %s -> NCCL: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
    } while (0)

} // namespace

struct GGComm {
    int rank = 0, n = 1;
    ncclComm_t nccl = nullptr; // NCCL transport
    gg_group *grp = nullptr;   // in-process transport
};

void gg_comm_release(gg_context *c) {
    if (!c->comm) return;
    if (c->comm->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm->nccl);
    delete c->comm;
    c->comm = nullptr;
}

int gg_comm_ranks(const gg_context *c) { return c && c->comm ? c->comm->n : 1; }

void gg_comm_abort(gg_context *c) {
    if (c && c->comm && c->comm->grp && c->comm->n > 1) c->comm->grp->fail();
}

int gg_comm_allgather_dev(gg_context *c, const void *sendDev, void *recvDev, size_t bytes) {
    GGComm *m = c->comm;
    if (!m) return gg_fail(GG_ERR_ARG, "gg_comm_allgather_dev: no communicator");
    if (m->grp) {
        gg_group *g = m->grp;
        cudaError_t e = cudaStreamSynchronize(c->st); // this rank's record must be complete before a peer reads it
        g->sendBase[m->rank] = (const char *)sendDev;
        g->dev[m->rank] = c->device;
        if (e != cudaSuccess) g->fail();
        if (!g->barrier()) return gg_fail(GG_ERR_CUDA, "gg_comm_allgather_dev: a rank of the in-process group failed");
        for (int r = 0; r < m->n && e == cudaSuccess; ++r) {
            char *dst = (char *)recvDev + (size_t)r * bytes;
            if (g->dev[r] == c->device) e = cudaMemcpyAsync(dst, g->sendBase[r], bytes, cudaMemcpyDeviceToDevice, c->st);
            else e = cudaMemcpyPeerAsync(dst, c->device, g->sendBase[r], g->dev[r], bytes, c->st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
        if (e != cudaSuccess) g->fail();
        if (!g->barrier()) // (every rank has read the records: they may be overwritten)
            return gg_fail(GG_ERR_CUDA, "gg_comm_allgather_dev: a rank of the in-process group failed (%s)", cudaGetErrorString(e));
        return GG_OK;
    }
    if (!g_nccl.AllGather) return gg_fail(GG_ERR_UNSUPPORTED, "gg_comm_allgather_dev: NCCL is not loaded");
    ncclResult_t r_ = g_nccl.AllGather(sendDev, recvDev, bytes, ncclChar, m->nccl, c->st);
    if (r_ != ncclSuccess) return gg_fail(GG_ERR_CUDA, "gg_comm_allgather_dev: NCCL: %s", g_nccl.GetErrorString(r_));
    return GG_OK;
}

namespace {

// every rank contributes `bytes` of host memory; all[r * bytes ..] receives rank r's
int allgather_host(gg_context *c, const void *mine, size_t bytes, void *all) {
    GGComm *m = c->comm;
    if (m->grp) {
        gg_group *g = m->grp;
        if (bytes > GG_SMALL_BYTES) return gg_fail(GG_ERR_ARG, "gg_comm_allgather: %zu bytes per rank (limit %d)", bytes, GG_SMALL_BYTES);
        memcpy(g->small[m->rank], mine, bytes);
        if (!g->barrier()) return gg_fail(GG_ERR_ARG, "gg_comm: a rank of the in-process group failed");
        for (int r = 0; r < m->n; ++r) memcpy((char *)all + (size_t)r * bytes, g->small[r], bytes);
        if (!g->barrier()) return gg_fail(GG_ERR_ARG, "gg_comm: a rank of the in-process group failed"); // (nobody overwrites its slot before everyone has read it)
        return GG_OK;
    }
    int rc;
    const size_t slot = (bytes + 15) & ~(size_t)15;
    if ((rc = gg_ensure(c, c->commscratch, slot * m->n + 16))) return rc;
    char *d = (char *)c->commscratch.p;
    CK(cudaMemcpyAsync(d + slot * m->rank, mine, bytes, cudaMemcpyHostToDevice, c->st));
    NCK(g_nccl.AllGather(d + slot * m->rank, d, slot, ncclChar, m->nccl, c->st));
    if (slot == bytes) CK(cudaMemcpyAsync(all, d, bytes * m->n, cudaMemcpyDeviceToHost, c->st));
    else
        CK(cudaMemcpy2DAsync(all, bytes, d, slot, bytes, (size_t)m->n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

// rank r's piece for rank p lies at sendBase + sendOff[p] (sendBytes[p] bytes) and lands at recvBase + recvOff[r'] on p
int alltoallv_device(gg_context *c, const char *sendBase, const size_t *sendOff, const size_t *sendBytes, char *recvBase,
                     const size_t *recvOff, const size_t *recvBytes) {
    GGComm *m = c->comm;
    if (m->grp) {
        gg_group *g = m->grp;
        CK(cudaStreamSynchronize(c->st)); // the pieces must be complete before a peer's copy engine reads them
        g->sendBase[m->rank] = sendBase;
        g->dev[m->rank] = c->device;
        for (int p = 0; p < m->n; ++p) { g->sendOff[m->rank][p] = sendOff[p]; g->sendBytes[m->rank][p] = sendBytes[p]; }
        if (!g->barrier()) return gg_fail(GG_ERR_ARG, "gg_exchange: a rank of the in-process group failed");
        cudaError_t e = cudaSuccess;
        for (int p = 0; p < m->n && e == cudaSuccess; ++p) {
            if (p == m->rank || recvBytes[p] == 0) continue;
            const char *src = g->sendBase[p] + g->sendOff[p][m->rank];
            if (g->dev[p] == c->device) e = cudaMemcpyAsync(recvBase + recvOff[p], src, recvBytes[p], cudaMemcpyDeviceToDevice, c->st);
            else e = cudaMemcpyPeerAsync(recvBase + recvOff[p], c->device, src, g->dev[p], recvBytes[p], c->st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
        if (e != cudaSuccess) {
            g->fail();
            return gg_fail(GG_ERR_CUDA, "gg_exchange: peer copy failed: %s", cudaGetErrorString(e));
        }
        if (!g->barrier()) return gg_fail(GG_ERR_ARG, "gg_exchange: a rank of the in-process group failed"); // (every rank has pulled its pieces: the send buffers may be reused)
        return GG_OK;
    }
    NCK(g_nccl.GroupStart());
    for (int p = 0; p < m->n; ++p) {
        if (p == m->rank) continue;
        if (sendBytes[p]) NCK(g_nccl.Send(sendBase + sendOff[p], sendBytes[p], ncclChar, p, m->nccl, c->st));
        if (recvBytes[p]) NCK(g_nccl.Recv(recvBase + recvOff[p], recvBytes[p], ncclChar, p, m->nccl, c->st));
    }
    NCK(g_nccl.GroupEnd());
    return GG_OK;
}

} // namespace

extern "C" {

int gg_comm_unique_id(void *id) {
    if (!id) return gg_fail(GG_ERR_ARG, "gg_comm_unique_id: null");
    int rc = load_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == GG_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    NCK(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return GG_OK;
}

int gg_comm_init(gg_context *c, const void *id, int rank, int nRanks) {
    if (!c || !id || nRanks < 1 || nRanks > GG_MAX_RANKS || rank < 0 || rank >= nRanks)
        return gg_fail(GG_ERR_ARG, "gg_comm_init: rank %d of %d (at most %d ranks)", rank, nRanks, GG_MAX_RANKS);
    int rc = load_nccl();
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    gg_comm_release(c);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    NCK(g_nccl.CommInitRank(&comm, nRanks, u, rank));
    c->comm = new GGComm();
    c->comm->rank = rank; c->comm->n = nRanks; c->comm->nccl = comm;
    return GG_OK;
}

int gg_group_create(gg_group **pg, int nRanks) {
    if (!pg || nRanks < 1 || nRanks > GG_MAX_RANKS) return gg_fail(GG_ERR_ARG, "gg_group_create: %d ranks (1..%d)", nRanks, GG_MAX_RANKS);
    gg_group *g = new gg_group();
    g->n = nRanks;
    *pg = g;
    return GG_OK;
}

void gg_group_destroy(gg_group *g) { delete g; }

int gg_comm_init_local(gg_context *c, gg_group *g, int rank) {
    if (!c || !g || rank < 0 || rank >= g->n) return gg_fail(GG_ERR_ARG, "gg_comm_init_local: bad argument");
    CK(cudaSetDevice(c->device));
    gg_comm_release(c);
    c->comm = new GGComm();
    c->comm->rank = rank; c->comm->n = g->n; c->comm->grp = g;
    // direct loads over NVLink when the ranks of the group sit on different GPUs (ignored where not possible)
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) == cudaSuccess)
        for (int d = 0; d < nDev; ++d)
            if (d != c->device) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, c->device, d) == cudaSuccess && can)
                    if (cudaDeviceEnablePeerAccess(d, 0) != cudaSuccess) cudaGetLastError();
            }
    cudaGetLastError();
    return GG_OK;
}

int gg_comm_free(gg_context *c) {
    if (!c) return gg_fail(GG_ERR_ARG, "gg_comm_free: null");
    cudaSetDevice(c->device);
    gg_comm_release(c);
    return GG_OK;
}

int gg_comm_info(gg_context *c, int *pRank, int *pnRanks, int *pTransport, int *pNcclVersion) {
    if (!c || !c->comm) return gg_fail(GG_ERR_ARG, "gg_comm_info: no communicator (gg_comm_init / gg_comm_init_local)");
    if (pRank) *pRank = c->comm->rank;
    if (pnRanks) *pnRanks = c->comm->n;
    if (pTransport) *pTransport = c->comm->grp ? 1 : 0;
    if (pNcclVersion) {
        *pNcclVersion = 0;
        if (!c->comm->grp && g_nccl.GetVersion) g_nccl.GetVersion(pNcclVersion);
    }
    return GG_OK;
}

int gg_comm_allgather(gg_context *c, const void *mine, size_t bytes, void *all) {
    if (!c || !c->comm || !mine || !all || bytes == 0) return gg_fail(GG_ERR_ARG, "gg_comm_allgather: bad argument / no communicator");
    CK(cudaSetDevice(c->device));
    return allgather_host(c, mine, bytes, all);
}

static int exchange_impl(gg_context *c, const gg_params *prm, const double *bndAll, gg_exchange_stats *stats);

int gg_exchange(gg_context *c, const gg_params *prm, const double *bndAll, gg_exchange_stats *stats) {
    const int rc = exchange_impl(c, prm, bndAll, stats);
    // in-process group: the other ranks are (or will be) waiting in this collective's barriers -- let them go
    if (rc != GG_OK) gg_comm_abort(c);
    return rc;
}

static int exchange_impl(gg_context *c, const gg_params *prm, const double *bndAll, gg_exchange_stats *stats) {
    if (!c || !prm) return gg_fail(GG_ERR_ARG, "gg_exchange: null argument");
    if (!c->comm) return gg_fail(GG_ERR_ARG, "gg_exchange: no communicator (gg_comm_init / gg_comm_init_local)");
    if (c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_exchange: no local domain (gg_set_local / gg_build_local)");
    GGComm *m = c->comm;
    if (c->idSelf != m->rank) return gg_fail(GG_ERR_ARG, "gg_exchange: the local domain is %d but this is rank %d", c->idSelf, m->rank);
    CK(cudaSetDevice(c->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    // forget the previous step's remote domains; the top tree (gg_set_top) stays
    c->dom.resize(1);
    c->nNodesAll = c->dom[0].nNodes;
    c->nPartAll = c->dom[0].nPart;
    if (m->n == 1) return GG_OK;
    int rc;
    const int n = m->n, me = m->rank;
    if ((rc = gg_early_ewald(c, prm))) return rc; // (beside the exchange; gg_gravity picks it up)
    CK(cudaEventRecord(c->evx[0], c->st));
    // ---- the other domains' root bounds: given by the host (it has them in its top tree) or gathered here
    double bnd[GG_MAX_RANKS][6];
    if (bndAll) memcpy(bnd, bndAll, sizeof(double) * 6 * n);
    else {
        if (!c->haveRootBnd) return gg_fail(GG_ERR_ARG, "gg_exchange: the bounds of the local root are unknown (gg_tree.bnd was NULL) and bndAll was not given");
        if ((rc = allgather_host(c, c->rootBnd, sizeof(c->rootBnd), bnd))) return rc;
    }
    double others[GG_MAX_RANKS][6];
    int who[GG_MAX_RANKS], nR = 0;
    for (int r = 0; r < n; ++r)
        if (r != me) { memcpy(others[nR], bnd[r], sizeof(bnd[r])); who[nR++] = r; }
    // ---- prune the local tree against every other domain's box
    size_t offs[GG_MAX_RANKS + 1];
    int hdr[3 * GG_MAX_RANKS];
    const int launches0 = c->nLaunches;
    if ((rc = gg_let_export_impl(c, nR, &others[0][0], prm, offs, hdr))) return rc;
    CK(cudaEventRecord(c->evx[1], c->st));
    // ---- sizes: meta[p] = what this rank sends to rank p (bytes, nNodes, nPart, iRoot); everybody learns all rows
    long long meta[GG_MAX_RANKS][4], allMeta[GG_MAX_RANKS][GG_MAX_RANKS][4];
    memset(meta, 0, sizeof(meta));
    size_t sendOff[GG_MAX_RANKS] = {0}, sendBytes[GG_MAX_RANKS] = {0};
    for (int k = 0; k < nR; ++k) {
        const int p = who[k];
        const size_t bytes = (size_t)hdr[3 * k] * (sizeof(NodeW) + 128 + 48) + (size_t)hdr[3 * k + 1] * sizeof(PartS);
        meta[p][0] = (long long)bytes; meta[p][1] = hdr[3 * k]; meta[p][2] = hdr[3 * k + 1]; meta[p][3] = hdr[3 * k + 2];
        sendOff[p] = offs[k];
        sendBytes[p] = bytes;
    }
    {
        long long flat[GG_MAX_RANKS * GG_MAX_RANKS * 4];
        if ((rc = allgather_host(c, meta, sizeof(long long) * 4 * n, flat))) return rc;
        for (int r = 0; r < n; ++r)
            for (int p = 0; p < n; ++p) memcpy(allMeta[r][p], &flat[((size_t)r * n + p) * 4], sizeof(long long) * 4);
    }
    size_t recvOff[GG_MAX_RANKS] = {0}, recvBytes[GG_MAX_RANKS] = {0}, total = 0, sent = 0;
    for (int r = 0; r < n; ++r) {
        if (r == me) continue;
        recvOff[r] = total;
        recvBytes[r] = (size_t)allMeta[r][me][0];
        total += (recvBytes[r] + 255) & ~(size_t)255;
        sent += sendBytes[r];
    }
    if ((rc = gg_ensure(c, c->letrecv, total + 256))) return rc;
    // room for everything that arrives, reserved in one go (gg_ingest_packed then never re-allocates mid-way)
    {
        size_t addN = 0, addP = 0;
        for (int r = 0; r < n; ++r)
            if (r != me) { addN += (size_t)allMeta[r][me][1]; addP += (size_t)allMeta[r][me][2]; }
        const size_t keepN = (size_t)c->nNodesAll, keepP = (size_t)c->nPartAll;
        if ((rc = gg_finish_mom(c))) return rc;
        if ((rc = gg_ensure(c, c->nodes, (keepN + addN + GG_MAX_TOP) * sizeof(NodeW), keepN * sizeof(NodeW)))) return rc;
        if ((rc = gg_ensure(c, c->momf, (keepN + addN + GG_MAX_TOP) * 128, keepN * 128))) return rc;
        if ((rc = gg_ensure(c, c->momq, (keepN + addN + GG_MAX_TOP) * 48, keepN * 48))) return rc;
        if ((rc = gg_ensure(c, c->parts, (keepP + addP + 1) * sizeof(PartS), keepP * sizeof(PartS)))) return rc;
    }
    // ---- the trees travel
    if ((rc = alltoallv_device(c, (const char *)c->letout.p, sendOff, sendBytes, (char *)c->letrecv.p, recvOff, recvBytes))) return rc;
    CK(cudaEventRecord(c->evx[2], c->st));
    // ---- ingest behind the local domain, in rank order (stream-ordered; the receive buffer is the context's own)
    for (int r = 0; r < n; ++r) {
        if (r == me) continue;
        const int h3[3] = {(int)allMeta[r][me][1], (int)allMeta[r][me][2], (int)allMeta[r][me][3]};
        if ((rc = gg_ingest_packed(c, r, h3, (const char *)c->letrecv.p + recvOff[r]))) return rc;
        c->nLaunches += 1;
    }
    CK(cudaEventRecord(c->evx[3], c->st));
    if (stats) {
        CK(cudaStreamSynchronize(c->st));
        float ms;
        CK(cudaEventElapsedTime(&ms, c->evx[0], c->evx[1])); stats->msExport = ms;
        CK(cudaEventElapsedTime(&ms, c->evx[1], c->evx[2])); stats->msTransfer = ms;
        CK(cudaEventElapsedTime(&ms, c->evx[2], c->evx[3])); stats->msIngest = ms;
        CK(cudaEventElapsedTime(&ms, c->evx[0], c->evx[3])); stats->msTotal = ms;
        stats->bytesSent = (double)sent;
        double got = 0;
        for (int r = 0; r < n; ++r) got += (double)recvBytes[r];
        stats->bytesReceived = got;
        stats->nKernelLaunches = c->nLaunches - launches0;
        size_t whole = (size_t)c->dom[0].nNodes * (sizeof(NodeW) + 128 + 48) + (size_t)c->dom[0].nPart * sizeof(PartS);
        stats->bytesWholeDomain = (double)whole;
    }
    return GG_OK;
}

} // extern "C"
