// gg_api.cu -- the C ABI (include/gasoline_b200.h): context, ingestion of the host's tree/particles into the
// device layout, and the orchestration of one force evaluation (= pkdGravAll, pkd.c:2868-3060).
#include <cub/device/device_scan.cuh>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <time.h>
#include <string>
#include <thread>
#include <vector>
#include "gg_context.h"
#include "gg_m2m.h"

void gg_ewald_table_host(const double *root, double L, double fhCut, int iOrder, std::vector<double> &ewt);

namespace {

thread_local char g_err[512] = "";

} // namespace

int gg_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    fprintf(stderr, "gasoline_b200: %s\n", g_err);
    return code;
}

namespace {

// GG_TRACE=1: host-side phase timings on stderr (development aid)
struct Trace {
    bool on;
    double t0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    Trace() : on(getenv("GG_TRACE") != nullptr), t0(now()) {}
    void mark(const char *what) {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[gg trace] %-28s %8.3f ms\n", what, t - t0);
        t0 = t;
    }
};

} // namespace

int gg_ensure(gg_context *c, DevBuf &b, size_t bytes, size_t preserve) {
    if (bytes <= b.cap) return GG_OK;
    size_t cap = bytes + bytes / 4 + 256;
    void *np = nullptr;
    CK(cudaMalloc(&np, cap));
    if (preserve && b.p) CK(cudaMemcpyAsync(np, b.p, preserve, cudaMemcpyDeviceToDevice, c->st));
    if (b.p) { // work reading the old buffer may be in flight on any of the context's streams
        CK(cudaStreamSynchronize(c->st));
        if (c->st2) CK(cudaStreamSynchronize(c->st2));
        if (c->st3) CK(cudaStreamSynchronize(c->st3));
        if (c->st4) CK(cudaStreamSynchronize(c->st4));
        CK(cudaFree(b.p));
    }
    b.p = np;
    b.cap = cap;
    return GG_OK;
}

namespace {

int ensure_pinned(gg_context *c, size_t bytes) {
    if (bytes <= c->pinnedCap) return GG_OK;
    if (c->pinned) CK(cudaFreeHost(c->pinned));
    c->pinned = nullptr;
    c->pinnedCap = 0;
    CK(cudaMallocHost(&c->pinned, bytes + bytes / 4));
    c->pinnedCap = bytes + bytes / 4;
    return GG_OK;
}

struct Images {
    std::vector<double> off;
    int n = 0, home = 0, bits = 5;
};

// Image offsets in the reference's loop order (walk.c:325-337): ix outermost, non-periodic axes not replicated.
Images make_images(const gg_params *prm) {
    Images im;
    const int nR = prm->nReps;
    for (int ix = -nR; ix <= nR; ++ix) {
        if (ix && prm->fPeriod[0] >= DBL_MAX) continue;
        for (int iy = -nR; iy <= nR; ++iy) {
            if (iy && prm->fPeriod[1] >= DBL_MAX) continue;
            for (int iz = -nR; iz <= nR; ++iz) {
                if (iz && prm->fPeriod[2] >= DBL_MAX) continue;
                if (!ix && !iy && !iz) im.home = im.n;
                im.off.push_back(ix * prm->fPeriod[0]);
                im.off.push_back(iy * prm->fPeriod[1]);
                im.off.push_back(iz * prm->fPeriod[2]);
                ++im.n;
            }
        }
    }
    im.bits = im.n <= 32 ? 5 : (im.n <= 128 ? 7 : (im.n <= 512 ? 9 : (im.n <= 1024 ? 10 : 11)));
    return im;
}

} // namespace

// Wait for the asynchronous half of the last gg_set_local (moments) before touching what it reads or writes.
int gg_finish_mom(gg_context *c) {
    if (c->momPending) {
        CK(cudaStreamSynchronize(c->st2));
        c->momPending = false;
    }
    return GG_OK;
}

namespace {

// raw staging layout for one domain (doubles): r[3n] fMass[n] fSoft[n] fOpen2[n] mom[31n]; ints: pLower pUpper
// iLower iUpper [n each]
__global__ void k_pack_nodes(int n, const double *r, const double *fMass, const double *fSoft, const double *fOpen2,
                             const int *pLower, const int *pUpper, const int *iLower, const int *iUpper, int nodeBase,
                             int partBase, NodeW *nodes, int nPartDomain, int *chk) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    // validation the host used to do in a serial loop (0.8 ms per 357 k cells): chk[0] = largest bucket,
    // chk[1] = buckets whose particle range leaves [0, nPartDomain)
    int nb = 0, bad = 0;
    if (i < n && iLower[i] == -1) {
        nb = pUpper[i] - pLower[i] + 1;
        bad = pLower[i] < 0 || pUpper[i] >= nPartDomain;
    }
    nb = __reduce_max_sync(0xffffffffu, nb);
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (nb > 0) atomicMax(&chk[0], nb);
        if (bad) atomicAdd(&chk[1], bad);
    }
    if (i >= n) return;
    NodeW w;
    w.rx = r[3 * (size_t)i]; w.ry = r[3 * (size_t)i + 1]; w.rz = r[3 * (size_t)i + 2];
    w.fOpen2 = fOpen2[i]; w.fSoft = fSoft[i]; w.fMass = fMass[i];
    int c0 = iLower[i], c1 = -1;
    if (c0 >= 0) { // second child = the first child's "next", unless that is already this cell's own next
        int cand = iUpper[c0];
        if (cand >= 0 && cand != iUpper[i]) c1 = cand;
    }
    w.c0 = c0 >= 0 ? c0 + nodeBase : -1;
    w.c1 = c1 >= 0 ? c1 + nodeBase : -1;
    w.pLower = pLower[i] + partBase;
    w.nP = pUpper[i] - pLower[i] + 1;
    if (w.nP < 0) w.nP = 0;
    nodes[nodeBase + i] = w;
}

// The evaluation records: FP32 moments (quadrupole made traceless like SETILIST, walk.c:41-48) and the raw FP64
// quadrupole of the softened-cell path.
__global__ void k_pack_mom(int n, const double *mom, int nodeBase, float4 *momf, double *momq) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *q = &mom[(size_t)GG_NMOM * i];
    float f[32];
    gg_pack_momf(q, f);
    float4 *o = &momf[(size_t)(nodeBase + i) * 8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
    double *oq = &momq[(size_t)(nodeBase + i) * 6];
#pragma unroll
    for (int k = 0; k < 6; ++k) oq[k] = q[k];
}

__global__ void k_pack_parts(int n, const double *x, const double *y, const double *z, const double *m,
                             const double *h, int partBase, PartS *parts, double *hsoft) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PartS p;
    p.x = x[i]; p.y = y[i]; p.z = z[i];
    p.m = (float)m[i]; p.h = (float)h[i];
    parts[partBase + i] = p;
    if (hsoft) hsoft[i] = h[i];
}

// gg_gravity_chunked: task index and first particle of each of nChunk + 1 boundaries -- boundary k starts at task
// k nTasks / nChunk, moved forward to the first pass of a bucket (the passes of one bucket stay in one chunk).
__global__ void k_chunk_bounds(const Task *tasks, int nTasks, const NodeW *nodes, int nChunk, int nPart, int *out) {
    const int k = threadIdx.x;
    if (k > nChunk) return;
    int t = (int)((long long)k * nTasks / nChunk);
    while (t < nTasks && tasks[t].pass != 0) ++t;
    out[2 * k] = t;
    out[2 * k + 1] = k == 0 ? 0 : (t < nTasks ? nodes[tasks[t].node].pLower : nPart);
}

// A remote domain arriving in device record layout: copy the walk records with links / particle indices rebased
// from the owner's numbering (base 0) to this rank's global numbering.
__global__ void k_rebase_nodes(int n, const NodeW *src, int nodeBase, int partBase, NodeW *dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    NodeW w = src[i];
    if (w.c0 >= 0) w.c0 += nodeBase;
    if (w.c1 >= 0) w.c1 += nodeBase;
    w.pLower += partBase;
    dst[nodeBase + i] = w;
}

// ---------------------------------------------------------------------------------------------- LET export
// flags per (remote, node): bit 0 = the remote rank's walks can reach the cell, bit 1 = ... and may open it
struct LetArgs {
    const NodeW *nodes;
    int nNodes, nRemote, nImages;
    const double *imgOff;  // [nImages][3]
    const double *bnd;     // [nRemote][6]
    unsigned char *flag;   // [nRemote][nNodes]
    const unsigned *front; // this level: (remote << 28) | node
    unsigned *next;
    int *count;            // [0] this level, [1] next level
};

// One level of the marking walk: every frontier item tests its cell against the remote box under all image offsets
// (opened under ANY offset = opened in the pruned tree, which all images share) and pushes the children.
__global__ void k_let_level(const LetArgs A, int level) {
    const int n = A.count[level & 1];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const unsigned it = A.front[t];
        const int r = (int)(it >> 28), i = (int)(it & 0x0fffffffu);
        const NodeW w = A.nodes[i];
        bool open = w.nP < 4; // walk.c:81
        if (!open) {
            const double *b = &A.bnd[6 * r];
            for (int k = 0; k < A.nImages && !open; ++k) {
                const double x = w.rx + A.imgOff[3 * k], y = w.ry + A.imgOff[3 * k + 1], z = w.rz + A.imgOff[3 * k + 2];
                double dx = b[0] - x, dy = b[1] - y, dz = b[2] - z;
                const double ex = x - b[3], ey = y - b[4], ez = z - b[5];
                dx = dx > ex ? dx : ex; dy = dy > ey ? dy : ey; dz = dz > ez ? dz : ez;
                dx = dx > 0.0 ? dx : 0.0; dy = dy > 0.0 ? dy : 0.0; dz = dz > 0.0 ? dz : 0.0;
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                open = d2 <= w.fOpen2;
            }
        }
        A.flag[(size_t)r * A.nNodes + i] = open ? 3 : 1;
        if (open && w.c0 >= 0) {
            const int nc = w.c1 >= 0 ? 2 : 1;
            const int pos = atomicAdd(&A.count[(level + 1) & 1], nc);
            A.next[pos] = ((unsigned)r << 28) | (unsigned)w.c0;
            if (nc == 2) A.next[pos + 1] = ((unsigned)r << 28) | (unsigned)w.c1;
        }
    }
}

__global__ void k_let_seed(unsigned *front, int *count, int nRemote, int iRoot) {
    const int r = threadIdx.x;
    if (r < nRemote) front[r] = ((unsigned)r << 28) | (unsigned)iRoot;
    if (r == 0) { count[0] = nRemote; count[1] = 0; }
}

__global__ void k_let_reset(int *count, int level) { count[(level + 1) & 1] = 0; }

// per (remote, node): 1 if kept / the particles it contributes (opened buckets only).  All remotes in one launch; the
// per-remote segments are nn + 1 long (one trailing zero), so ONE exclusive scan over the concatenation numbers every
// remote's kept nodes and travelling particles, and segment r's base is the scan value at r * (nn + 1).
__global__ void k_let_counts(int nn, int nRemote, const NodeW *nodes, const unsigned char *flag, int *keep, int *npart) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t seg = (size_t)nn + 1;
    if (t >= seg * nRemote + 1) return;
    const int r = (int)(t / seg), i = (int)(t - (size_t)r * seg);
    int k = 0, np = 0;
    if (r < nRemote && i < nn) {
        const unsigned char f = flag[(size_t)r * nn + i];
        k = f & 1;
        if (f == 3) {
            const int4 d = __ldg(reinterpret_cast<const int4 *>(&nodes[i]) + 3);
            if (d.x < 0) np = d.w;
        }
    }
    keep[t] = k;
    npart[t] = np;
}

// the scan values at the segment boundaries -> a small array for one copy to the host
__global__ void k_let_bounds(int nn, int nRemote, const int *newIdx, const int *newPart, int *out) {
    const int r = threadIdx.x;
    if (r > nRemote) return;
    const size_t at = (size_t)r * ((size_t)nn + 1);
    out[2 * r] = newIdx[at];
    out[2 * r + 1] = newPart[at];
}

struct LetPackArgs {
    int nn, nRemote;
    const NodeW *nodes;
    const float4 *momf;
    const double *momq;
    const PartS *parts;
    const unsigned char *flag;
    const int *newIdx, *newPart;
    int baseIdx[16], basePart[16];
    NodeW *oNodes[16];
    float4 *oMomf[16];
    double *oMomq[16];
    PartS *oParts[16];
};

// 8 lanes per kept node: the 64 B walk record (re-linked), the 128 B moment record and the 48 B quadrupole move as 16 B
// pieces; the particles of an opened bucket are copied by the same 8 lanes
__global__ void __launch_bounds__(256) k_let_pack(const LetPackArgs A) {
    const size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int piece = threadIdx.x & 7;
    if (t >= (size_t)A.nn * A.nRemote) return;
    const int r = (int)(t / A.nn), i = (int)(t - (size_t)r * A.nn);
    const unsigned char f = A.flag[t];
    if (!(f & 1)) return;
    const size_t seg = (size_t)r * ((size_t)A.nn + 1);
    const int *nIdx = A.newIdx + seg;
    const int o = nIdx[i] - A.baseIdx[r];
    const bool open = f == 3;
    const NodeW *src = &A.nodes[i];
    const int4 links = __ldg(reinterpret_cast<const int4 *>(src) + 3);
    const int p0 = A.newPart[seg + i] - A.basePart[r];
    if (piece < 3) reinterpret_cast<uint4 *>(&A.oNodes[r][o])[piece] = __ldg(reinterpret_cast<const uint4 *>(src) + piece);
    else if (piece == 3) {
        int4 w = links;
        if (open && w.x >= 0) { // opened cell: children re-numbered
            w.x = nIdx[links.x] - A.baseIdx[r];
            w.y = links.y >= 0 ? nIdx[links.y] - A.baseIdx[r] : -1;
            w.z = 0;
        } else if (open) w.z = p0; // opened bucket: its particles travel
        else { w.x = w.y = -1; w.z = 0; } // never opened over there: a childless cell (nP >= 4 keeps it clear of walk.c:81)
        reinterpret_cast<int4 *>(&A.oNodes[r][o])[3] = w;
    }
    A.oMomf[r][(size_t)o * 8 + piece] = __ldg(&A.momf[(size_t)i * 8 + piece]);
    if (piece < 3)
        reinterpret_cast<uint4 *>(&A.oMomq[r][(size_t)o * 6])[piece] = __ldg(reinterpret_cast<const uint4 *>(&A.momq[(size_t)i * 6]) + piece);
    if (open && links.x < 0) {
        const uint4 *ps = reinterpret_cast<const uint4 *>(&A.parts[links.z]);
        uint4 *pd = reinterpret_cast<uint4 *>(&A.oParts[r][p0]);
        for (int k = piece; k < 2 * links.w; k += 8) pd[k] = __ldg(ps + k);
    }
}

// number of 8-sink passes each local bucket needs (0 for cells and for buckets without an active sink)
__global__ void k_count_groups(int nNodes, const NodeW *nodes, const int *active, int *ngroups, int *isBucket) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nNodes) return;
    NodeW w = nodes[i];
    int g = 0;
    if (w.c0 < 0) {
        int n = 0;
        if (active) {
            for (int j = 0; j < w.nP; ++j) n += active[w.pLower + j] != 0;
        } else n = w.nP;
        g = (n + GG_MAX_SINKS - 1) / GG_MAX_SINKS;
    }
    ngroups[i] = g;
    isBucket[i] = g > 0;
}

__global__ void k_fill_tasks(int nNodes, const int *ngroups, const int *offs, const int *ords, Task *tasks,
                             int *bucketNode) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nNodes) return;
    for (int g = 0; g < ngroups[i]; ++g) tasks[offs[i] + g] = Task{i, g, ords[i], 0};
    if (ngroups[i] > 0) bucketNode[ords[i]] = i;
}

__global__ void k_max_bucket(int n, const int *pLower, const int *pUpper, const int *iLower, int *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int v = (i < n && iLower[i] == -1) ? pUpper[i] - pLower[i] + 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

// Uniform background for comoving, non-periodic runs (pkd.c:2967-2991).
__global__ void k_comove(int n, const PartS *parts, const int *active, double dRhoFac, double *acc, double *pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (active && !active[i])) return;
    PartS p = parts[i];
    double r2 = p.x * p.x + p.y * p.y + p.z * p.z;
    acc[3 * (size_t)i] += dRhoFac * p.x;
    acc[3 * (size_t)i + 1] += dRhoFac * p.y;
    acc[3 * (size_t)i + 2] += dRhoFac * p.z;
    pot[i] -= 0.5 * dRhoFac * r2;
}

// FP32 FMA-pipe peak: 8 independent dependent-FFMA chains per thread, all SMs, enough warps to hide the 4-cycle
// FFMA latency.  Gives the denominator of the FP32 roofline on THIS GPU at its clocks under load.
__global__ void __launch_bounds__(256) k_fma_peak(int iters, float *out) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
          a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f + blockIdx.x * 1e-9f, b = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, m, b); a1 = fmaf(a1, m, b); a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b);
            a4 = fmaf(a4, m, b); a5 = fmaf(a5, m, b); a6 = fmaf(a6, m, b); a7 = fmaf(a7, m, b);
        }
    }
    const float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456f) out[0] = r; // never true: keeps the chains alive
}

// The local domain's task list: sink buckets with >= 1 ACTIVE particle in tree order, one task per pass of 8 sinks.
// Depends only on the tree and the ACTIVE flags, so it is built at upload time; hCounts receives {nTasks, nBuckets}
// after the caller's next synchronisation of c->st.
int build_task_list(gg_context *c, int nn, const int *dActive, int *hCounts) {
    int rc;
    if ((rc = gg_ensure(c, c->ngroups, (size_t)(nn + 1) * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->goffs, (size_t)(nn + 1) * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->isb, (size_t)(nn + 1) * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->boffs, (size_t)(nn + 1) * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->bnode, (size_t)(nn + 1) * sizeof(int)))) return rc;
    // every bucket has >= 1 particle, a bucket of nP particles has <= ceil(nP/8) passes: nn + nPart/8 bounds the tasks
    if ((rc = gg_ensure(c, c->tasks, ((size_t)nn + (size_t)c->nPartUpload / GG_MAX_SINKS + 2) * sizeof(Task)))) return rc;
    k_count_groups<<<(nn + 255) / 256, 256, 0, c->st>>>(nn, (const NodeW *)c->nodes.p, dActive, (int *)c->ngroups.p,
                                                        (int *)c->isb.p);
    CK(cudaGetLastError());
    size_t tmpBytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, (int *)c->ngroups.p, (int *)c->goffs.p, nn + 1, c->st));
    if ((rc = gg_ensure(c, c->cubtmp, tmpBytes))) return rc;
    CK(cudaMemsetAsync((int *)c->ngroups.p + nn, 0, sizeof(int), c->st));
    CK(cudaMemsetAsync((int *)c->isb.p + nn, 0, sizeof(int), c->st));
    CK(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmpBytes, (int *)c->ngroups.p, (int *)c->goffs.p, nn + 1, c->st));
    CK(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmpBytes, (int *)c->isb.p, (int *)c->boffs.p, nn + 1, c->st));
    k_fill_tasks<<<(nn + 255) / 256, 256, 0, c->st>>>(nn, (const int *)c->ngroups.p, (const int *)c->goffs.p,
                                                      (const int *)c->boffs.p, (Task *)c->tasks.p, (int *)c->bnode.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&hCounts[0], (int *)c->goffs.p + nn, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(&hCounts[1], (int *)c->boffs.p + nn, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    c->nLaunches += 5;
    return GG_OK;
}

int make_ewald_args(gg_context *c, const gg_params *prm, EwaldKernelArgs &ea, cudaStream_t st);

// the parameters the Ewald correction depends on (ewald.c:15-178, 182-248)
bool same_ewald_params(const gg_params &a, const gg_params &b) {
    return a.bPeriodic == b.bPeriodic && a.bEwald == b.bEwald && a.iEwOrder == b.iEwOrder && a.nReps == b.nReps &&
           a.fEwCut == b.fEwCut && a.fEwhCut == b.fEwhCut && a.fPeriod[0] == b.fPeriod[0] && a.fPeriod[1] == b.fPeriod[1] &&
           a.fPeriod[2] == b.fPeriod[2];
}

// Before anything else touches acc / pot / nloop or the particle records: order the side stream's early Ewald (if any)
// before the main stream, and forget its results.
int drop_early_ewald(gg_context *c) {
    if (c->ewPending) {
        CK(cudaStreamWaitEvent(c->st, c->evEw[2], 0));
        c->ewPending = false;
    }
    c->ewValid = false;
    return GG_OK;
}

// Preconditions of an early Ewald correction for parameters p on the loaded / arriving local domain.
bool early_ewald_ok(const gg_context *c, const gg_params &p) {
    return p.bPeriodic && p.bEwald && p.iEwOrder > 0 && p.iEwOrder <= 4 && p.nReps >= 0 && p.nReps <= 3 && c->haveRoot &&
           !(p.flags & GG_FLAG_WALK_ONLY);
}

#ifndef GG_EARLY_SLICE
#define GG_EARLY_SLICE (1 << 20) // particles per slice of the early-Ewald upload (40 MB: ~0.7 ms of PCIe 5, ~1.7 ms of k_ewald)
#endif

int upload_domain(gg_context *c, const gg_tree *t, const gg_particles *pp, int nodeBase, int partBase, bool local,
                  bool onDevice) {
    const int nn = t->nNodes, np = pp->n;
    const cudaMemcpyKind kind = onDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const bool devMom = t->mom == nullptr; // the cells' moments are formed on the device from the particles (gg_moments.cu)
    const size_t nMomD = devMom ? 0 : (size_t)GG_NMOM * nn;
    size_t nd = (size_t)nn * (3 + 3) + nMomD + (size_t)np * 5;
    int rc;
    Trace tr;
    if ((rc = gg_finish_mom(c))) return rc;
    tr.mark("  upload: finish_mom");
    if ((rc = gg_ensure(c, c->raw, nd * sizeof(double)))) return rc;
    if ((rc = gg_ensure(c, c->rawi, (size_t)nn * 4 * sizeof(int)))) return rc;
    double *d = (double *)c->raw.p;
    double *dr = d, *dM = dr + 3 * (size_t)nn, *dS = dM + nn, *dO = dS + nn, *dmom = dO + nn;
    double *dx = dmom + nMomD, *dy = dx + np, *dz = dy + np, *dm = dz + np, *dh = dm + np;
    int *di = (int *)c->rawi.p;
    const size_t keepN = (size_t)nodeBase, keepP = (size_t)partBase;
    if ((rc = gg_ensure(c, c->nodes, (keepN + nn + GG_MAX_TOP) * sizeof(NodeW), keepN * sizeof(NodeW)))) return rc;
    if ((rc = gg_ensure(c, c->momf, (keepN + nn + GG_MAX_TOP) * 128, keepN * 128))) return rc;
    if ((rc = gg_ensure(c, c->momq, (keepN + nn + GG_MAX_TOP) * 48, keepN * 48))) return rc;
    if ((rc = gg_ensure(c, c->parts, (keepP + np + 1) * sizeof(PartS), keepP * sizeof(PartS)))) return rc;
    if (local && (rc = gg_ensure(c, c->hsoft, (size_t)(np + 1) * sizeof(double)))) return rc;
    // ---- announced periodic + Ewald evaluation (gg_announce): the particles go up FIRST, slice by slice, and every slice's
    //      Ewald correction starts on the side stream as soon as the slice is packed -- the FP64 kernel then runs beside
    //      the copies of the remaining particles and of the tree instead of after them
    const bool early = local && !onDevice && c->annValid && early_ewald_ok(c, c->ann) && np > 0 && partBase == 0;
    if (local && (rc = drop_early_ewald(c))) return rc;
    if (early) {
        if ((rc = gg_ensure(c, c->acc, (size_t)(np + 1) * 3 * sizeof(double)))) return rc;
        if ((rc = gg_ensure(c, c->pot, (size_t)(np + 1) * sizeof(double)))) return rc;
        if ((rc = gg_ensure(c, c->nloop, (size_t)(np + 1) * sizeof(int)))) return rc;
        const int *dActive = nullptr;
        if (pp->active) {
            if ((rc = gg_ensure(c, c->active, (size_t)(np + 1) * sizeof(int)))) return rc;
            CK(cudaMemcpyAsync(c->active.p, pp->active, sizeof(int) * np, cudaMemcpyDefault, c->st));
            dActive = (const int *)c->active.p;
        }
        EwaldKernelArgs ea;
        CK(cudaEventRecord(c->evPacked, c->st)); // (everything queued on the main stream so far, incl. the ACTIVE flags)
        CK(cudaStreamWaitEvent(c->st4, c->evPacked, 0));
        if ((rc = make_ewald_args(c, &c->ann, ea, c->st4))) return rc;
        ea.active = dActive;
        CK(cudaMemsetAsync(c->acc.p, 0, (size_t)np * 3 * sizeof(double), c->st4));
        CK(cudaMemsetAsync(c->pot.p, 0, (size_t)np * sizeof(double), c->st4));
        CK(cudaEventRecord(c->evEw[0], c->st4));
        const int nSlice = (np + GG_EARLY_SLICE - 1) / GG_EARLY_SLICE;
        while ((int)c->evSlice.size() < nSlice) {
            cudaEvent_t ev;
            CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            c->evSlice.push_back(ev);
        }
        for (int k = 0; k < nSlice; ++k) {
            const int lo = k * GG_EARLY_SLICE, hi = lo + GG_EARLY_SLICE < np ? lo + GG_EARLY_SLICE : np, m = hi - lo;
            CK(cudaMemcpyAsync(dx + lo, pp->x + lo, sizeof(double) * m, kind, c->st));
            CK(cudaMemcpyAsync(dy + lo, pp->y + lo, sizeof(double) * m, kind, c->st));
            CK(cudaMemcpyAsync(dz + lo, pp->z + lo, sizeof(double) * m, kind, c->st));
            CK(cudaMemcpyAsync(dm + lo, pp->fMass + lo, sizeof(double) * m, kind, c->st));
            CK(cudaMemcpyAsync(dh + lo, pp->fSoft + lo, sizeof(double) * m, kind, c->st));
            k_pack_parts<<<(m + 255) / 256, 256, 0, c->st>>>(m, dx + lo, dy + lo, dz + lo, dm + lo, dh + lo, partBase + lo,
                                                             (PartS *)c->parts.p, (double *)c->hsoft.p + lo);
            CK(cudaGetLastError());
            CK(cudaEventRecord(c->evSlice[k], c->st));
            CK(cudaStreamWaitEvent(c->st4, c->evSlice[k], 0));
            ea.first = lo;
            ea.n = hi;
            CK(gg_launch_ewald_kernel(ea, c->st4));
            c->nLaunches += 2;
        }
        CK(cudaEventRecord(c->evEw[1], c->st4));
        CK(cudaEventRecord(c->evEw[2], c->st4));
        c->ewPending = true;
        c->ewValid = true;
        c->ewPrm = c->ann;
        memcpy(c->ewRoot, c->root, sizeof(c->ewRoot));
        c->ewNEwh = ea.nEwh;
        c->ewN = np;
    }
    // ---- what the walk needs, on the main stream (issued first: the copy engine serves it first)
    if (nn > 0) {
        CK(cudaMemcpyAsync(dr, t->r, sizeof(double) * 3 * nn, kind, c->st));
        CK(cudaMemcpyAsync(dM, t->fMass, sizeof(double) * nn, kind, c->st));
        CK(cudaMemcpyAsync(dS, t->fSoft, sizeof(double) * nn, kind, c->st));
        CK(cudaMemcpyAsync(dO, t->fOpen2, sizeof(double) * nn, kind, c->st));
        CK(cudaMemcpyAsync(di, t->pLower, sizeof(int) * nn, kind, c->st));
        CK(cudaMemcpyAsync(di + nn, t->pUpper, sizeof(int) * nn, kind, c->st));
        CK(cudaMemcpyAsync(di + 2 * (size_t)nn, t->iLower, sizeof(int) * nn, kind, c->st));
        CK(cudaMemcpyAsync(di + 3 * (size_t)nn, t->iUpper, sizeof(int) * nn, kind, c->st));
    }
    if (np > 0 && !early) {
        CK(cudaMemcpyAsync(dx, pp->x, sizeof(double) * np, kind, c->st));
        CK(cudaMemcpyAsync(dy, pp->y, sizeof(double) * np, kind, c->st));
        CK(cudaMemcpyAsync(dz, pp->z, sizeof(double) * np, kind, c->st));
        CK(cudaMemcpyAsync(dm, pp->fMass, sizeof(double) * np, kind, c->st));
        CK(cudaMemcpyAsync(dh, pp->fSoft, sizeof(double) * np, kind, c->st));
    }
    // ---- the moments (60 % of the bytes; only k_eval reads them): the LOCAL domain's go on a second stream and are
    //      not waited for here, so that the walk of the following gg_gravity overlaps their transfer
    cudaStream_t ms = local ? c->st2 : c->st;
    if (nn > 0) {
        if (!devMom) {
            CK(cudaMemcpyAsync(dmom, t->mom, sizeof(double) * GG_NMOM * nn, kind, ms));
            k_pack_mom<<<(nn + 127) / 128, 128, 0, ms>>>(nn, dmom, nodeBase, (float4 *)c->momf.p, (double *)c->momq.p);
            CK(cudaGetLastError());
            ++c->nLaunches;
            if (local) {
                CK(cudaEventRecord(c->evMom, c->st2));
                c->momPending = true;
            }
        }
        if ((rc = gg_ensure(c, c->misc, 16 * sizeof(int)))) return rc;
        CK(cudaMemsetAsync((int *)c->misc.p + 8, 0, 2 * sizeof(int), c->st));
        k_pack_nodes<<<(nn + 127) / 128, 128, 0, c->st>>>(nn, dr, dM, dS, dO, di, di + nn, di + 2 * (size_t)nn,
                                                          di + 3 * (size_t)nn, nodeBase, partBase, (NodeW *)c->nodes.p,
                                                          np, (int *)c->misc.p + 8);
        CK(cudaGetLastError());
        ++c->nLaunches;
    }
    if (np > 0 && !early) {
        k_pack_parts<<<(np + 255) / 256, 256, 0, c->st>>>(np, dx, dy, dz, dm, dh, partBase, (PartS *)c->parts.p,
                                                          local ? (double *)c->hsoft.p : nullptr);
        CK(cudaGetLastError());
        ++c->nLaunches;
    }
    if (devMom && nn > 0) {
        // moments from the packed walk records + the FP64 particle columns still in the staging buffer; for the local
        // domain on the second stream, beside the walk of the following gg_gravity (k_eval waits for evMom)
        if ((rc = gg_ensure(c, c->momraw, (size_t)nn * 32 * sizeof(double)))) return rc;
        if ((rc = gg_ensure(c, c->mparent, (size_t)nn * 2 * sizeof(int)))) return rc;
        if (local) {
            CK(cudaEventRecord(c->evPacked, c->st));
            CK(cudaStreamWaitEvent(c->st2, c->evPacked, 0));
        }
        CK(gg_launch_device_moments(nn, (const NodeW *)c->nodes.p, nodeBase, partBase, t->iRoot, dx, dy, dz, dm,
                                    (int *)c->mparent.p, (int *)c->mparent.p + nn, (double *)c->momraw.p,
                                    (float4 *)c->momf.p, (double *)c->momq.p, ms));
        c->nLaunches += 2;
        if (local) {
            CK(cudaEventRecord(c->evMom, c->st2));
            c->momPending = true;
        }
    }
    // the local domain's sink-bucket task list depends only on the tree and the ACTIVE flags: built here, so that
    // gg_gravity starts its walk without a host round trip
    int hChk[2] = {0, 0}, hCounts[2] = {0, 0};
    if (local) {
        const int *dActive = nullptr;
        if (pp->active) {
            if ((rc = gg_ensure(c, c->active, (size_t)(np + 1) * sizeof(int)))) return rc;
            if (!early) CK(cudaMemcpyAsync(c->active.p, pp->active, sizeof(int) * np, cudaMemcpyDefault, c->st));
            dActive = (const int *)c->active.p;
        }
        c->nPartUpload = np;
        if ((rc = build_task_list(c, nn, dActive, hCounts))) return rc;
    }
    if (nn > 0) CK(cudaMemcpyAsync(hChk, (int *)c->misc.p + 8, sizeof(hChk), cudaMemcpyDeviceToHost, c->st));
    tr.mark("  upload: enqueue");
    // the staging buffer is reused by the next upload: finish the packing first (the moment half: finish_mom)
    CK(cudaStreamSynchronize(c->st));
    tr.mark("  upload: sync");
    if (hChk[1])
        return gg_fail(GG_ERR_ARG, "gg_set_%s: %d bucket(s) span particles outside [0,%d)", local ? "local" : "remote", hChk[1], np);
    if (hChk[0] > GG_MAX_BUCKET)
        return gg_fail(GG_ERR_UNSUPPORTED, "gg_set_%s: a bucket holds %d particles (limit GG_MAX_BUCKET=%d)",
                    local ? "local" : "remote", hChk[0], GG_MAX_BUCKET);
    if (local) {
        c->maxBucket = hChk[0] > 1 ? hChk[0] : 1;
        c->nTasksLocal = hCounts[0];
        c->nBucketsLocal = hCounts[1];
    } else if (hChk[0] > c->maxBucket) c->maxBucket = hChk[0];
    return GG_OK;
}

} // namespace

// The Ewald correction of the RESIDENT local domain, started on the side stream (gg_exchange calls this first: the FP64
// kernel then runs beside the pruning kernels, the size all-gather's host round trips, the NCCL transfer and the ingest).
// gg_gravity picks it up like the one gg_set_local starts for an announced upload.
int gg_early_ewald(gg_context *c, const gg_params *prm) {
    if (c->dom.empty() || c->ewValid || c->sunMode || !early_ewald_ok(c, *prm)) return GG_OK;
    if (c->rootLazy) return GG_OK; // (the root expansion would have to be fetched first: ordinary order)
    const int np = c->dom[0].nPart;
    if (np <= 0) return GG_OK;
    int rc;
    if ((rc = drop_early_ewald(c))) return rc;
    if ((rc = gg_ensure(c, c->acc, (size_t)(np + 1) * 3 * sizeof(double)))) return rc;
    if ((rc = gg_ensure(c, c->pot, (size_t)(np + 1) * sizeof(double)))) return rc;
    if ((rc = gg_ensure(c, c->nloop, (size_t)(np + 1) * sizeof(int)))) return rc;
    EwaldKernelArgs ea;
    CK(cudaEventRecord(c->evPacked, c->st)); // (whatever still reads or writes the result arrays on the main stream)
    CK(cudaStreamWaitEvent(c->st4, c->evPacked, 0));
    if ((rc = make_ewald_args(c, prm, ea, c->st4))) return rc;
    ea.active = c->hActive.empty() ? nullptr : (const int *)c->active.p;
    ea.first = 0;
    ea.n = np;
    CK(cudaMemsetAsync(c->acc.p, 0, (size_t)np * 3 * sizeof(double), c->st4));
    CK(cudaMemsetAsync(c->pot.p, 0, (size_t)np * sizeof(double), c->st4));
    CK(cudaEventRecord(c->evEw[0], c->st4));
    CK(gg_launch_ewald_kernel(ea, c->st4));
    ++c->nLaunches;
    CK(cudaEventRecord(c->evEw[1], c->st4));
    CK(cudaEventRecord(c->evEw[2], c->st4));
    c->ewPending = true;
    c->ewValid = true;
    c->ewPrm = *prm;
    memcpy(c->ewRoot, c->root, sizeof(c->ewRoot));
    c->ewNEwh = ea.nEwh;
    c->ewN = np;
    return GG_OK;
}

extern "C" {

const char *gg_last_error(void) { return g_err; }
int gg_version(void) { return 100; }

int gg_device_count(int *pn) {
    if (!pn) return gg_fail(GG_ERR_ARG, "gg_device_count: null");
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess) return gg_fail(GG_ERR_CUDA, "gg_device_count: %s", cudaGetErrorString(e));
    *pn = nDev;
    return GG_OK;
}

int gg_create(gg_context **pctx, int device) {
    if (!pctx) return gg_fail(GG_ERR_ARG, "gg_create: null out pointer");
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev == 0)
        return gg_fail(GG_ERR_CUDA, "gg_create: no usable CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= nDev) return gg_fail(GG_ERR_ARG, "gg_create: device %d of %d", device, nDev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return gg_fail(GG_ERR_UNSUPPORTED, "gg_create: device %s is sm_%d%d; this build targets sm_100a only", prop.name,
                    prop.major, prop.minor);
    gg_context *c = new gg_context();
    c->device = device;
    c->nSM = prop.multiProcessorCount;
    // the main stream outranks the side streams: kernels of the exchange / walk / evaluation get an SM slot as soon as one
    // frees up, an early Ewald correction (st4) only fills what they leave
    int prLeast = 0, prGreatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));
    CK(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, prGreatest));
    CK(cudaStreamCreateWithFlags(&c->st2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->evMom, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c->st3, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->evWalk, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evStats, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->evPacked, cudaEventDisableTiming));
    CK(cudaStreamCreateWithPriority(&c->st4, cudaStreamNonBlocking, prLeast));
    CK(cudaEventCreate(&c->evEw[0]));
    CK(cudaEventCreate(&c->evEw[1]));
    CK(cudaEventCreateWithFlags(&c->evEw[2], cudaEventDisableTiming));
    for (auto &ev : c->ev) CK(cudaEventCreate(&ev));
    for (auto &ev : c->evx) CK(cudaEventCreate(&ev));
    for (auto &ev : c->evt) CK(cudaEventCreate(&ev));
    *pctx = c;
    return GG_OK;
}

void gg_destroy(gg_context *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    cudaStreamSynchronize(c->st2);
    if (c->st3) cudaStreamSynchronize(c->st3);
    if (c->st4) cudaStreamSynchronize(c->st4);
    DevBuf *all[] = {&c->nodes, &c->momf, &c->momq, &c->parts, &c->active, &c->hsoft, &c->tasks, &c->ngroups,
                     &c->goffs, &c->counts, &c->acc, &c->pot, &c->dtg, &c->fweight, &c->nloop, &c->sums, &c->misc,
                     &c->imgoff, &c->ewt, &c->raw, &c->rawi, &c->cubtmp, &c->flush, &c->pool,
                     &c->nextblk, &c->poolmask, &c->isb, &c->boffs, &c->bnode, &c->ghead, &c->gcnt, &c->bcnt, &c->btot,
                     &c->boff64, &c->lists, &c->letflag, &c->letfront, &c->letidx, &c->letout, &c->letmisc, &c->momraw,
                     &c->mparent, &c->dbgtask, &c->momout, &c->sx, &c->sy, &c->sz, &c->sm, &c->sh, &c->sact, &c->svel, &c->sid, &c->sdt,
                     &c->svel2, &c->sid2, &c->sdt2, &c->sacc, &c->srhist, &c->ox, &c->oy, &c->oz, &c->ow, &c->ocell,
                     &c->okeys, &c->ocnt, &c->opart, &c->osums, &c->obis, &c->oans};
    for (DevBuf *b : all)
        if (b->p) cudaFree(b->p);
    gg_comm_release(c);
    if (c->letrecv.p) cudaFree(c->letrecv.p);
    if (c->commscratch.p) cudaFree(c->commscratch.p);
    for (auto &ev : c->evx)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->evt)
        if (ev) cudaEventDestroy(ev);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->builder) gg_builder_free(c->builder);
    for (auto &ev : c->ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->st);
    cudaStreamDestroy(c->st2);
    if (c->st3) cudaStreamDestroy(c->st3);
    if (c->st4) cudaStreamDestroy(c->st4);
    for (auto &ev : c->evSlice) cudaEventDestroy(ev);
    for (auto &ev : c->evEw)
        if (ev) cudaEventDestroy(ev);
    if (c->evWalk) cudaEventDestroy(c->evWalk);
    if (c->evStats) cudaEventDestroy(c->evStats);
    if (c->evPacked) cudaEventDestroy(c->evPacked);
    cudaEventDestroy(c->evMom);
    delete c;
}

int gg_host_alloc(void **p, size_t bytes) {
    if (!p) return gg_fail(GG_ERR_ARG, "gg_host_alloc: null");
    CK(cudaMallocHost(p, bytes ? bytes : 1));
    return GG_OK;
}
int gg_host_free(void *p) {
    if (p) CK(cudaFreeHost(p));
    return GG_OK;
}

int gg_set_local(gg_context *c, int idSelf, const gg_tree *t, const gg_particles *pp) {
    if (!c || !t || !pp) return gg_fail(GG_ERR_ARG, "gg_set_local: null argument");
    if (t->nNodes < 1 || pp->n < 0 || t->iRoot < 0 || t->iRoot >= t->nNodes)
        return gg_fail(GG_ERR_ARG, "gg_set_local: nNodes=%d n=%d iRoot=%d", t->nNodes, pp->n, t->iRoot);
    CK(cudaSetDevice(c->device));
    Trace tr;
    c->dom.clear();
    c->nTop = 0;
    c->idSelf = idSelf;
    c->built = GGBuiltDev{};
    c->rootLazy = false;
    c->stateN = 0; // a host-supplied domain replaces any resident store
    int rc = upload_domain(c, t, pp, 0, 0, true, false);
    if (rc) return rc;
    tr.mark("set_local: upload_domain");
    c->dom.push_back(Domain{idSelf, t->nNodes, pp->n, t->iRoot, 0, 0});
    c->nNodesAll = t->nNodes;
    c->nPartAll = pp->n;
    c->haveRootBnd = t->bnd != nullptr;
    if (t->bnd) memcpy(c->rootBnd, t->bnd + 6 * (size_t)t->iRoot, sizeof(c->rootBnd));
    if (pp->active) c->hActive.assign(pp->active, pp->active + pp->n); // (the device copy went up with the domain)
    else c->hActive.clear();
    return GG_OK;
}

// ---- gg_set_local in slices (see the header).  The staging layout and every kernel are upload_domain's.
int gg_local_begin(gg_context *c, int idSelf, int nNodes, int iRoot, int nPart, const double *rootBnd, int bActive) {
    if (!c) return gg_fail(GG_ERR_ARG, "gg_local_begin: null context");
    if (nNodes < 1 || nPart < 0 || iRoot < 0 || iRoot >= nNodes)
        return gg_fail(GG_ERR_ARG, "gg_local_begin: nNodes=%d n=%d iRoot=%d", nNodes, nPart, iRoot);
    CK(cudaSetDevice(c->device));
    c->dom.clear();
    c->nTop = 0;
    c->idSelf = idSelf;
    c->built = GGBuiltDev{};
    c->rootLazy = false;
    c->stateN = 0;
    gg_context::Sliced &S = c->sl;
    S = gg_context::Sliced{};
    int rc;
    if ((rc = gg_finish_mom(c))) return rc;
    if ((rc = drop_early_ewald(c))) return rc;
    const size_t nn = (size_t)nNodes, np = (size_t)nPart;
    if ((rc = gg_ensure(c, c->raw, (nn * 6 + np * 5) * sizeof(double)))) return rc;
    if ((rc = gg_ensure(c, c->rawi, nn * 4 * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->nodes, (nn + GG_MAX_TOP) * sizeof(NodeW)))) return rc;
    if ((rc = gg_ensure(c, c->momf, (nn + GG_MAX_TOP) * 128))) return rc;
    if ((rc = gg_ensure(c, c->momq, (nn + GG_MAX_TOP) * 48))) return rc;
    if ((rc = gg_ensure(c, c->parts, (np + 1) * sizeof(PartS)))) return rc;
    if ((rc = gg_ensure(c, c->hsoft, (np + 1) * sizeof(double)))) return rc;
    if ((rc = gg_ensure(c, c->misc, 16 * sizeof(int)))) return rc;
    if (bActive && (rc = gg_ensure(c, c->active, (np + 1) * sizeof(int)))) return rc;
    double *d = (double *)c->raw.p;
    S.dr = d; S.dM = S.dr + 3 * nn; S.dS = S.dM + nn; S.dO = S.dS + nn;
    S.dx = S.dO + nn; S.dy = S.dx + np; S.dz = S.dy + np; S.dm = S.dz + np; S.dh = S.dm + np;
    S.di = (int *)c->rawi.p;
    S.nn = nNodes; S.np = nPart; S.iRoot = iRoot; S.idSelf = idSelf; S.active = bActive != 0;
    c->haveRootBnd = rootBnd != nullptr;
    if (rootBnd) memcpy(c->rootBnd, rootBnd, sizeof(c->rootBnd));
    if (bActive) c->hActive.assign(np, 0);
    else c->hActive.clear();
    CK(cudaMemsetAsync((int *)c->misc.p + 8, 0, 2 * sizeof(int), c->st));
    S.early = c->annValid && early_ewald_ok(c, c->ann) && nPart > 0;
    if (S.early) {
        if ((rc = gg_ensure(c, c->acc, (np + 1) * 3 * sizeof(double)))) return rc;
        if ((rc = gg_ensure(c, c->pot, (np + 1) * sizeof(double)))) return rc;
        if ((rc = gg_ensure(c, c->nloop, (np + 1) * sizeof(int)))) return rc;
        CK(cudaEventRecord(c->evPacked, c->st));
        CK(cudaStreamWaitEvent(c->st4, c->evPacked, 0));
        if ((rc = make_ewald_args(c, &c->ann, S.ea, c->st4))) return rc;
        S.ea.active = bActive ? (const int *)c->active.p : nullptr;
        CK(cudaMemsetAsync(c->acc.p, 0, np * 3 * sizeof(double), c->st4));
        CK(cudaMemsetAsync(c->pot.p, 0, np * sizeof(double), c->st4));
        CK(cudaEventRecord(c->evEw[0], c->st4));
    }
    S.open = true;
    return GG_OK;
}

int gg_local_particles(gg_context *c, int first, int count, const double *x, const double *y, const double *z,
                       const double *fMass, const double *fSoft, const int *active) {
    if (!c || !c->sl.open) return gg_fail(GG_ERR_ARG, "gg_local_particles: no gg_local_begin");
    gg_context::Sliced &S = c->sl;
    if (count <= 0) return GG_OK;
    if (first < 0 || first + count > S.np || !x || !y || !z || !fMass || !fSoft || (S.active && !active))
        return gg_fail(GG_ERR_ARG, "gg_local_particles: slice [%d, %d) of %d particles / null array", first, first + count, S.np);
    CK(cudaSetDevice(c->device));
    const size_t lo = (size_t)first, nb = sizeof(double) * (size_t)count;
    if (S.active) {
        CK(cudaMemcpyAsync((int *)c->active.p + lo, active, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, c->st));
        memcpy(c->hActive.data() + lo, active, sizeof(int) * (size_t)count);
    }
    CK(cudaMemcpyAsync(S.dx + lo, x, nb, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.dy + lo, y, nb, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.dz + lo, z, nb, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.dm + lo, fMass, nb, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.dh + lo, fSoft, nb, cudaMemcpyHostToDevice, c->st));
    k_pack_parts<<<(count + 255) / 256, 256, 0, c->st>>>(count, S.dx + lo, S.dy + lo, S.dz + lo, S.dm + lo, S.dh + lo, first,
                                                         (PartS *)c->parts.p, (double *)c->hsoft.p + lo);
    CK(cudaGetLastError());
    ++c->nLaunches;
    if (S.early) {
        const size_t k = (size_t)S.nEv++;
        if (k >= c->evSlice.size()) {
            cudaEvent_t ev;
            CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            c->evSlice.push_back(ev);
        }
        CK(cudaEventRecord(c->evSlice[k], c->st));
        CK(cudaStreamWaitEvent(c->st4, c->evSlice[k], 0));
        S.ea.first = first;
        S.ea.n = first + count;
        CK(gg_launch_ewald_kernel(S.ea, c->st4));
        ++c->nLaunches;
    }
    S.gotP += count;
    return GG_OK;
}

int gg_local_nodes(gg_context *c, int first, int count, const double *r, const double *fMass, const double *fSoft,
                   const double *fOpen2, const int *pLower, const int *pUpper, const int *iLower, const int *iUpper) {
    if (!c || !c->sl.open) return gg_fail(GG_ERR_ARG, "gg_local_nodes: no gg_local_begin");
    gg_context::Sliced &S = c->sl;
    if (count <= 0) return GG_OK;
    if (first < 0 || first + count > S.nn || !r || !fMass || !fSoft || !fOpen2 || !pLower || !pUpper || !iLower || !iUpper)
        return gg_fail(GG_ERR_ARG, "gg_local_nodes: slice [%d, %d) of %d nodes / null array", first, first + count, S.nn);
    CK(cudaSetDevice(c->device));
    const size_t lo = (size_t)first, nn = (size_t)S.nn, nd = sizeof(double) * (size_t)count, ni = sizeof(int) * (size_t)count;
    CK(cudaMemcpyAsync(S.dr + 3 * lo, r, 3 * nd, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.dM + lo, fMass, nd, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.dS + lo, fSoft, nd, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.dO + lo, fOpen2, nd, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.di + lo, pLower, ni, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.di + nn + lo, pUpper, ni, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.di + 2 * nn + lo, iLower, ni, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(S.di + 3 * nn + lo, iUpper, ni, cudaMemcpyHostToDevice, c->st));
    S.gotN += count;
    return GG_OK;
}

int gg_local_end(gg_context *c) {
    if (!c || !c->sl.open) return gg_fail(GG_ERR_ARG, "gg_local_end: no gg_local_begin");
    gg_context::Sliced &S = c->sl;
    S.open = false;
    if (S.gotP != S.np || S.gotN != S.nn)
        return gg_fail(GG_ERR_ARG, "gg_local_end: %d of %d particles and %d of %d nodes arrived", S.gotP, S.np, S.gotN, S.nn);
    CK(cudaSetDevice(c->device));
    const int nn = S.nn, np = S.np;
    int rc;
    if (S.early) {
        CK(cudaEventRecord(c->evEw[1], c->st4));
        CK(cudaEventRecord(c->evEw[2], c->st4));
        c->ewPending = true;
        c->ewValid = true;
        c->ewPrm = c->ann;
        memcpy(c->ewRoot, c->root, sizeof(c->ewRoot));
        c->ewNEwh = S.ea.nEwh;
        c->ewN = np;
    }
    // links between cells cross slices (a cell reads its first child's iUpper): the node records are packed once all are here
    k_pack_nodes<<<(nn + 127) / 128, 128, 0, c->st>>>(nn, S.dr, S.dM, S.dS, S.dO, S.di, S.di + nn, S.di + 2 * (size_t)nn,
                                                      S.di + 3 * (size_t)nn, 0, 0, (NodeW *)c->nodes.p, np, (int *)c->misc.p + 8);
    CK(cudaGetLastError());
    ++c->nLaunches;
    if ((rc = gg_ensure(c, c->momraw, (size_t)nn * 32 * sizeof(double)))) return rc;
    if ((rc = gg_ensure(c, c->mparent, (size_t)nn * 2 * sizeof(int)))) return rc;
    CK(cudaEventRecord(c->evPacked, c->st));
    CK(cudaStreamWaitEvent(c->st2, c->evPacked, 0));
    CK(gg_launch_device_moments(nn, (const NodeW *)c->nodes.p, 0, 0, S.iRoot, S.dx, S.dy, S.dz, S.dm, (int *)c->mparent.p,
                                (int *)c->mparent.p + nn, (double *)c->momraw.p, (float4 *)c->momf.p, (double *)c->momq.p, c->st2));
    c->nLaunches += 2;
    CK(cudaEventRecord(c->evMom, c->st2));
    c->momPending = true;
    int hChk[2] = {0, 0}, hCounts[2] = {0, 0};
    c->nPartUpload = np;
    if ((rc = build_task_list(c, nn, S.active ? (const int *)c->active.p : nullptr, hCounts))) return rc;
    CK(cudaMemcpyAsync(hChk, (int *)c->misc.p + 8, sizeof(hChk), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (hChk[1]) return gg_fail(GG_ERR_ARG, "gg_local_end: %d bucket(s) span particles outside [0,%d)", hChk[1], np);
    if (hChk[0] > GG_MAX_BUCKET)
        return gg_fail(GG_ERR_UNSUPPORTED, "gg_local_end: a bucket holds %d particles (limit GG_MAX_BUCKET=%d)", hChk[0], GG_MAX_BUCKET);
    c->maxBucket = hChk[0] > 1 ? hChk[0] : 1;
    c->nTasksLocal = hCounts[0];
    c->nBucketsLocal = hCounts[1];
    c->dom.push_back(Domain{S.idSelf, nn, np, S.iRoot, 0, 0});
    c->nNodesAll = nn;
    c->nPartAll = np;
    return GG_OK;
}

// The Ewald root expansion (pkdCalcRoot, pkd.c:4395-4470: complete l <= 4 moments about the root's centre of mass) of a
// device-built tree: exactly the RAW moment record gg_moments.cu leaves for the root cell.
static int fetch_root_lazy(gg_context *c) {
    if (!c->rootLazy) return GG_OK;
    int rc = gg_finish_mom(c);
    if (rc) return rc;
    const int iRoot = c->dom[0].iRoot;
    double raw[32];
    NodeW w;
    CK(cudaMemcpyAsync(raw, (const double *)c->momraw.p + (size_t)iRoot * 32, sizeof(raw), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(&w, (const NodeW *)c->nodes.p + iRoot, sizeof(w), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    double *R = c->root;
    const double *q = raw + 1;
    R[0] = w.fMass; R[1] = w.rx; R[2] = w.ry; R[3] = w.rz;
    R[4] = q[0]; R[5] = q[1]; R[6] = q[3]; R[7] = q[4]; R[8] = q[5]; R[9] = q[2];
    for (int k = 6; k < 31; ++k) R[10 + (k - 6)] = q[k];
    c->haveRoot = true;
    c->rootLazy = false;
    return GG_OK;
}

// Build the tree of the particles pp (host or device pointers) on the device and load it as the local domain.
// pkdCalcOpen's criteria on the device: OPEN_JOSH, and the three that take Bmax itself (a zero factor in k_emit).
// OPEN_ABSPAR needs the radial moments B2..B6 summed particle by particle in the reference's order and its root finder
// (pkd.c:2182-2224): the host builder has them (gg_tree_build_open), the device builder does not.
static int open_factor(const char *who, int iOpenType, double *c23) {
    if (iOpenType == GG_OPEN_JOSH) *c23 = 2.0 / sqrt(3.0);
    else if (iOpenType == GG_OPEN_RELPAR || iOpenType == GG_OPEN_ABSTOT || iOpenType == GG_OPEN_RELTOT) *c23 = 0.0;
    else if (iOpenType == GG_OPEN_ABSPAR)
        return gg_fail(GG_ERR_UNSUPPORTED, "%s: OPEN_ABSPAR is built on the host (gg_tree_build_open), not on the device", who);
    else return gg_fail(GG_ERR_ARG, "%s: iOpenType=%d", who, iOpenType);
    return GG_OK;
}

static int build_and_load(gg_context *c, int idSelf, const gg_particles *pp, int nBucket, double dTheta, double c23,
                          int *iOrder, int *pnNodes, double *root) {
    int rc;
    if ((rc = gg_finish_mom(c))) return rc;
    if ((rc = drop_early_ewald(c))) return rc;
    char msg[400];
    int nl = 0;
    GGBuiltDev b{};
    CK(cudaEventRecord(c->ev[6], c->st));
    rc = gg_builder_run(&c->builder, pp, nBucket, dTheta, c23, c->st, &b, &nl, msg, sizeof(msg));
    if (rc) return gg_fail(rc, "%s", msg);
    CK(cudaEventRecord(c->ev[7], c->st));
    c->nLaunches += nl;
    c->dom.clear();
    c->nTop = 0;
    c->idSelf = idSelf;
    gg_tree t{};
    t.nNodes = b.nNodes; t.iRoot = 0;
    t.bnd = b.bnd; t.r = b.r; t.fMass = b.fMass; t.fSoft = b.fSoft; t.fOpen2 = b.fOpen2; t.mom = nullptr;
    t.pLower = b.pLower; t.pUpper = b.pUpper; t.iLower = b.iLower; t.iUpper = b.iUpper;
    gg_particles dp{};
    dp.n = b.nPart; dp.x = b.x; dp.y = b.y; dp.z = b.z; dp.fMass = b.m; dp.fSoft = b.h; dp.active = b.active;
    if ((rc = upload_domain(c, &t, &dp, 0, 0, true, true))) return rc;
    c->dom.push_back(Domain{idSelf, t.nNodes, dp.n, 0, 0, 0});
    c->nNodesAll = t.nNodes;
    c->nPartAll = dp.n;
    c->built = b;
    CK(cudaMemcpyAsync(c->rootBnd, b.bnd, sizeof(c->rootBnd), cudaMemcpyDeviceToHost, c->st)); // (pre-order: root = cell 0)
    c->haveRootBnd = true;
    if (b.active) {
        c->hActive.resize((size_t)dp.n);
        CK(cudaMemcpyAsync(c->hActive.data(), b.active, sizeof(int) * dp.n, cudaMemcpyDeviceToHost, c->st));
    } else c->hActive.clear();
    if (iOrder) CK(cudaMemcpyAsync(iOrder, b.iorder, sizeof(int) * dp.n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]));
    c->msBuild = ms;
    c->rootLazy = true;
    c->haveRoot = false;
    if (root) {
        if ((rc = fetch_root_lazy(c))) return rc;
        memcpy(root, c->root, sizeof(c->root));
    }
    if (pnNodes) *pnNodes = t.nNodes;
    return GG_OK;
}

int gg_build_local(gg_context *c, int idSelf, const gg_particles *pp, int nBucket, double dTheta, int *iOrder,
                   int *pnNodes, double *root) {
    return gg_build_local_open(c, idSelf, pp, nBucket, GG_OPEN_JOSH, dTheta, iOrder, pnNodes, root);
}

int gg_build_local_open(gg_context *c, int idSelf, const gg_particles *pp, int nBucket, int iOpenType, double dTheta,
                        int *iOrder, int *pnNodes, double *root) {
    if (!c || !pp) return gg_fail(GG_ERR_ARG, "gg_build_local: null argument");
    double c23;
    int rco = open_factor("gg_build_local", iOpenType, &c23);
    if (rco) return rco;
    if (pp->n < 1 || !pp->x || !pp->y || !pp->z || !pp->fMass || !pp->fSoft || nBucket < 1 || nBucket > GG_MAX_BUCKET ||
        !(dTheta > 0))
        return gg_fail(GG_ERR_ARG, "gg_build_local: n=%d nBucket=%d dTheta=%g", pp->n, nBucket, dTheta);
    CK(cudaSetDevice(c->device));
    c->stateN = 0; // a tree built from host particles replaces any resident store
    return build_and_load(c, idSelf, pp, nBucket, dTheta, c23, iOrder, pnNodes, root);
}


// ---- ORB domain decomposition: the per-rank services of pstDomainDecomp on the device ----

namespace {
int orb_query(const char *who, gg_context *c, int nCells, const int *iCell, const int *iDim, const double *fSplit, OrbQuery &q) {
    if (!c || c->orbN < 0) return gg_fail(GG_ERR_ARG, "%s: no particles loaded (gg_orb_load)", who);
    if (nCells < 1 || nCells > GG_ORB_MAX_SLOTS || !iCell) return gg_fail(GG_ERR_ARG, "%s: nCells=%d (1..%d)", who, nCells, GG_ORB_MAX_SLOTS);
    q.nSlots = nCells;
    for (int s = 0; s < nCells; ++s) {
        if (iCell[s] < 1 || iCell[s] >= GG_ORB_MAX_CELL) return gg_fail(GG_ERR_ARG, "%s: PST cell %d outside 1..%d", who, iCell[s], GG_ORB_MAX_CELL - 1);
        for (int t = 0; t < s; ++t)
            if (iCell[t] == iCell[s]) return gg_fail(GG_ERR_ARG, "%s: PST cell %d asked about twice", who, iCell[s]);
        q.cell[s] = iCell[s];
        q.dim[s] = iDim ? iDim[s] : 0;
        q.split[s] = fSplit ? fSplit[s] : 0.0;
        if (q.dim[s] < 0 || q.dim[s] > 2) return gg_fail(GG_ERR_ARG, "%s: split axis %d", who, q.dim[s]);
    }
    return GG_OK;
}
const double *orb_pos(gg_context *c, int k) {
    DevBuf *own[] = {&c->ox, &c->oy, &c->oz}, *st[] = {&c->sx, &c->sy, &c->sz};
    return (const double *)(c->orbState ? st[k]->p : own[k]->p);
}
} // namespace

int gg_orb_load(gg_context *c, int n, const double *x, const double *y, const double *z, const double *fWeight) {
    if (!c || n < 0) return gg_fail(GG_ERR_ARG, "gg_orb_load: bad argument (n=%d)", n);
    CK(cudaSetDevice(c->device));
    int rc;
    const bool fromState = !x && !y && !z;
    if (fromState) {
        if (c->stateN < 1 || n != c->stateN) return gg_fail(GG_ERR_ARG, "gg_orb_load: x == NULL means the resident store, which holds %d particles (n=%d)", c->stateN, n);
    } else if (n > 0 && (!x || !y || !z)) return gg_fail(GG_ERR_ARG, "gg_orb_load: x, y, z must all be given (or all NULL)");
    const size_t nb = sizeof(double) * (size_t)(n > 0 ? n : 1);
    if (!fromState) {
        const double *src[] = {x, y, z};
        DevBuf *dst[] = {&c->ox, &c->oy, &c->oz};
        for (int k = 0; k < 3; ++k) {
            if ((rc = gg_ensure(c, *dst[k], nb))) return rc;
            if (n > 0) CK(cudaMemcpyAsync(dst[k]->p, src[k], sizeof(double) * (size_t)n, cudaMemcpyDefault, c->st));
        }
    }
    if (fWeight) {
        if ((rc = gg_ensure(c, c->ow, nb))) return rc;
        if (n > 0) CK(cudaMemcpyAsync(c->ow.p, fWeight, sizeof(double) * (size_t)n, cudaMemcpyDefault, c->st));
    }
    if ((rc = gg_ensure(c, c->ocell, sizeof(int) * (size_t)(n > 0 ? n : 1))) || (rc = gg_ensure(c, c->okeys, sizeof(unsigned long long) * 6 * GG_ORB_MAX_SLOTS)) ||
        (rc = gg_ensure(c, c->ocnt, sizeof(int) * 2 * GG_ORB_MAX_SLOTS)) || (rc = gg_ensure(c, c->osums, sizeof(double) * 2 * GG_ORB_MAX_SLOTS)) ||
        (rc = gg_ensure(c, c->opart, gg_orb_part_bytes(n))))
        return rc;
    CK(gg_launch_orb_init(n, (int *)c->ocell.p, c->st));
    ++c->nLaunches;
    CK(cudaStreamSynchronize(c->st));
    c->orbN = n;
    c->orbState = fromState;
    c->orbWeights = fWeight != nullptr;
    return GG_OK;
}

int gg_orb_bounds(gg_context *c, int nCells, const int *iCell, double *bnd, int *nIn) {
    OrbQuery q;
    int rc;
    if ((rc = orb_query("gg_orb_bounds", c, nCells, iCell, nullptr, nullptr, q))) return rc;
    if (!bnd || !nIn) return gg_fail(GG_ERR_ARG, "gg_orb_bounds: bnd and nIn must be given");
    CK(cudaSetDevice(c->device));
    CK(gg_launch_orb_bounds(q, c->orbN, orb_pos(c, 0), orb_pos(c, 1), orb_pos(c, 2), (const int *)c->ocell.p,
                            (unsigned long long *)c->okeys.p, (int *)c->ocnt.p, c->st));
    ++c->nLaunches;
    unsigned long long keys[6 * GG_ORB_MAX_SLOTS];
    CK(cudaMemcpyAsync(keys, c->okeys.p, sizeof(unsigned long long) * 6 * nCells, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(nIn, c->ocnt.p, sizeof(int) * nCells, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    for (int i = 0; i < 6 * nCells; ++i) {
        const unsigned long long k = keys[i];
        const long long b = (k & 0x8000000000000000ull) ? (long long)(k & 0x7fffffffffffffffull) : (long long)~k;
        memcpy(&bnd[i], &b, 8);
    }
    // a cell without particles of this rank: fMin = +FLOAT_MAXVAL, fMax = -FLOAT_MAXVAL like pkdCalcBound (pkd.c:1304-1330)
    for (int s = 0; s < nCells; ++s)
        if (nIn[s] == 0)
            for (int k = 0; k < 6; ++k) bnd[6 * s + k] = k < 3 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return GG_OK;
}

int gg_orb_weight(gg_context *c, int nCells, const int *iCell, const int *iDim, const double *fSplit, int *nLow, int *nHigh,
                  double *fLow, double *fHigh) {
    OrbQuery q;
    int rc;
    if (!iDim || !fSplit || !nLow || !nHigh || !fLow || !fHigh) return gg_fail(GG_ERR_ARG, "gg_orb_weight: NULL argument");
    if ((rc = orb_query("gg_orb_weight", c, nCells, iCell, iDim, fSplit, q))) return rc;
    CK(cudaSetDevice(c->device));
    const double *w = c->orbWeights ? (const double *)c->ow.p : nullptr;
    CK(gg_launch_orb_weight(q, c->orbN, orb_pos(c, 0), orb_pos(c, 1), orb_pos(c, 2), w, (const int *)c->ocell.p, (int *)c->ocnt.p,
                            (double *)c->opart.p, (double *)c->osums.p, c->st));
    c->nLaunches += w ? 2 : 1;
    int cnt[2 * GG_ORB_MAX_SLOTS];
    double sums[2 * GG_ORB_MAX_SLOTS];
    CK(cudaMemcpyAsync(cnt, c->ocnt.p, sizeof(int) * 2 * nCells, cudaMemcpyDeviceToHost, c->st));
    if (w) CK(cudaMemcpyAsync(sums, c->osums.p, sizeof(double) * 2 * nCells, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    for (int s = 0; s < nCells; ++s) {
        nLow[s] = cnt[2 * s]; nHigh[s] = cnt[2 * s + 1];
        fLow[s] = w ? sums[2 * s] : (double)cnt[2 * s];       // fWeight = 1 for every particle (pkd.c:686: read sets it)
        fHigh[s] = w ? sums[2 * s + 1] : (double)cnt[2 * s + 1];
    }
    return GG_OK;
}

int gg_orb_bisect(gg_context *c, int nCells, const int *iCell, const int *iDim, const double *fLow, const double *fUp,
                  const int *bLive, const double *nLower, const double *nUpper, int bSplitWork, double *fSplit,
                  int *bHasSplit, int *ittr) {
    if (!iDim || !fLow || !fUp || !bLive || !nLower || !nUpper || !fSplit || !bHasSplit || !ittr)
        return gg_fail(GG_ERR_ARG, "gg_orb_bisect: NULL argument");
    OrbBisect h;
    memset(&h, 0, sizeof(h));
    int rc;
    if ((rc = orb_query("gg_orb_bisect", c, nCells, iCell, iDim, nullptr, h.q))) return rc;
    for (int s = 0; s < nCells; ++s) {
        if (!(nLower[s] > 0.0) || !(nUpper[s] > 0.0)) return gg_fail(GG_ERR_ARG, "gg_orb_bisect: cell %d has no ranks on one side", iCell[s]);
        h.fl[s] = fLow[s]; h.fu[s] = fUp[s];
        h.fmm[s] = (fLow[s] + fUp[s]) / 2;
        h.nLower[s] = nLower[s]; h.nUpper[s] = nUpper[s];
        h.live[s] = bLive[s] ? 1 : 0;
    }
    h.splitWork = bSplitWork ? 1 : 0;
    h.maxIttr = 64; // MAX_ITTR, pst.c:874
    CK(cudaSetDevice(c->device));
    if ((rc = gg_ensure(c, c->obis, sizeof(OrbBisect)))) return rc;
    const double *w = c->orbWeights ? (const double *)c->ow.p : nullptr;
    CK(gg_launch_orb_bisect((OrbBisect *)c->obis.p, h, c->orbN, orb_pos(c, 0), orb_pos(c, 1), orb_pos(c, 2), w,
                            (const int *)c->ocell.p, (int *)c->ocnt.p, (double *)c->opart.p, (double *)c->osums.p, c->st));
    c->nLaunches += 1 + (h.maxIttr + 1) * (w ? 3 : 2);
    CK(cudaMemcpyAsync(&h, c->obis.p, sizeof(OrbBisect), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    for (int s = 0; s < nCells; ++s) {
        fSplit[s] = h.q.split[s];
        bHasSplit[s] = h.hasSplit[s];
        ittr[s] = h.ittr[s];
    }
    return GG_OK;
}

// COLLECTIVE over the context's communicator: the same root finder when every rank holds PART of the particles.  Per trial
// each rank weighs its own particles, the ranks' answer records are all-gathered between the devices (in-stream with NCCL:
// no host round trip per trial either), and every rank's k_orb_decide adds them in rank order -- same bits, same branch,
// same split on all ranks (pst.c:1004-1030 adds the lower and upper sub-trees' answers the same way).  Every rank passes
// the same arguments; fSplit / bHasSplit / ittr come back identical everywhere.
static int orb_bisect_all(gg_context *c, int nCells, const int *iCell, const int *iDim, const double *fLow, const double *fUp,
                          const int *bLive, const double *nLower, const double *nUpper, int bSplitWork, double *fSplit,
                          int *bHasSplit, int *ittr);

int gg_orb_bisect_all(gg_context *c, int nCells, const int *iCell, const int *iDim, const double *fLow, const double *fUp,
                      const int *bLive, const double *nLower, const double *nUpper, int bSplitWork, double *fSplit,
                      int *bHasSplit, int *ittr) {
    const int rc = orb_bisect_all(c, nCells, iCell, iDim, fLow, fUp, bLive, nLower, nUpper, bSplitWork, fSplit, bHasSplit, ittr);
    if (rc != GG_OK) gg_comm_abort(c);
    return rc;
}

static int orb_bisect_all(gg_context *c, int nCells, const int *iCell, const int *iDim, const double *fLow, const double *fUp,
                          const int *bLive, const double *nLower, const double *nUpper, int bSplitWork, double *fSplit,
                          int *bHasSplit, int *ittr) {
    if (!c) return gg_fail(GG_ERR_ARG, "gg_orb_bisect_all: null context");
    const int nRanks = gg_comm_ranks(c);
    if (nRanks <= 1)
        return gg_orb_bisect(c, nCells, iCell, iDim, fLow, fUp, bLive, nLower, nUpper, bSplitWork, fSplit, bHasSplit, ittr);
    if (!iDim || !fLow || !fUp || !bLive || !nLower || !nUpper || !fSplit || !bHasSplit || !ittr)
        return gg_fail(GG_ERR_ARG, "gg_orb_bisect_all: NULL argument");
    OrbBisect h;
    memset(&h, 0, sizeof(h));
    int rc;
    if ((rc = orb_query("gg_orb_bisect_all", c, nCells, iCell, iDim, nullptr, h.q))) return rc;
    for (int s = 0; s < nCells; ++s) {
        if (!(nLower[s] > 0.0) || !(nUpper[s] > 0.0)) return gg_fail(GG_ERR_ARG, "gg_orb_bisect_all: cell %d has no ranks on one side", iCell[s]);
        h.fl[s] = fLow[s]; h.fu[s] = fUp[s];
        h.fmm[s] = (fLow[s] + fUp[s]) / 2;
        h.nLower[s] = nLower[s]; h.nUpper[s] = nUpper[s];
        h.live[s] = bLive[s] ? 1 : 0;
    }
    h.splitWork = bSplitWork ? 1 : 0;
    h.maxIttr = 64; // MAX_ITTR, pst.c:874
    CK(cudaSetDevice(c->device));
    if ((rc = gg_ensure(c, c->obis, sizeof(OrbBisect)))) return rc;
    // record 0: this rank's answer (the weighing kernels' sums / cnt point into it); records 1..nRanks: everybody's
    if ((rc = gg_ensure(c, c->oans, (size_t)GG_ORB_REC_BYTES * (nRanks + 1)))) return rc;
    unsigned char *mine = (unsigned char *)c->oans.p, *all = mine + GG_ORB_REC_BYTES;
    double *sums = (double *)mine;
    int *cnt = (int *)(mine + GG_ORB_REC_CNT);
    CK(cudaMemsetAsync(mine, 0, (size_t)GG_ORB_REC_BYTES * (nRanks + 1), c->st));
    const double *w = c->orbWeights ? (const double *)c->ow.p : nullptr;
    OrbBisect *B = (OrbBisect *)c->obis.p;
    CK(gg_launch_orb_bisect_begin(B, h, cnt, sums, w != nullptr, c->st));
    // Trials are queued GG_ORB_CHUNK at a time without a host synchronisation; between chunks the number of cells still
    // bisected is read back (the state is identical on every rank, so all ranks stop after the same chunk and have queued
    // the same collectives).  A trial past the last live cell costs a collective for nothing: the typical 20-50 trials of a
    // level end after 2-4 chunks instead of MAX_ITTR + 1 rounds.
    int rounds = 0;
    for (int t0 = 0; t0 <= h.maxIttr; t0 += GG_ORB_CHUNK) {
        for (int t = t0; t < t0 + GG_ORB_CHUNK && t <= h.maxIttr; ++t, ++rounds) {
            CK(gg_launch_orb_trial(B, h.q.nSlots, c->orbN, orb_pos(c, 0), orb_pos(c, 1), orb_pos(c, 2), w,
                                   (const int *)c->ocell.p, cnt, (double *)c->opart.p, sums, c->st));
            if ((rc = gg_comm_allgather_dev(c, mine, all, GG_ORB_REC_BYTES))) return rc;
            CK(gg_launch_orb_decide(B, cnt, sums, w != nullptr, nRanks, all, c->st));
        }
        int nLive = 0;
        CK(cudaMemcpyAsync(&nLive, &B->nLive, sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        if (nLive == 0) break;
    }
    c->nLaunches += 1 + rounds * (w ? 4 : 3);
    CK(cudaMemcpyAsync(&h, c->obis.p, sizeof(OrbBisect), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    for (int s = 0; s < nCells; ++s) {
        fSplit[s] = h.q.split[s];
        bHasSplit[s] = h.hasSplit[s];
        ittr[s] = h.ittr[s];
    }
    return GG_OK;
}

int gg_orb_split(gg_context *c, int nCells, const int *iCell, const int *iDim, const double *fSplit) {
    OrbQuery q;
    int rc;
    if (!iDim || !fSplit) return gg_fail(GG_ERR_ARG, "gg_orb_split: NULL argument");
    if ((rc = orb_query("gg_orb_split", c, nCells, iCell, iDim, fSplit, q))) return rc;
    for (int s = 0; s < nCells; ++s)
        if (2 * q.cell[s] + 1 >= GG_ORB_MAX_CELL) return gg_fail(GG_ERR_ARG, "gg_orb_split: children of PST cell %d exceed %d", q.cell[s], GG_ORB_MAX_CELL - 1);
    CK(cudaSetDevice(c->device));
    CK(gg_launch_orb_split(q, c->orbN, orb_pos(c, 0), orb_pos(c, 1), orb_pos(c, 2), (int *)c->ocell.p, c->st));
    ++c->nLaunches;
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_orb_split_wrap(gg_context *c, int nCells, const int *iCell, const int *iDim, const double *fSplit,
                      const double *fSplitInactive) {
    OrbQuery q;
    OrbWrap w;
    int rc;
    if (!iDim || !fSplit || !fSplitInactive) return gg_fail(GG_ERR_ARG, "gg_orb_split_wrap: NULL argument");
    if ((rc = orb_query("gg_orb_split_wrap", c, nCells, iCell, iDim, fSplit, q))) return rc;
    memset(&w, 0, sizeof(w));
    for (int s = 0; s < nCells; ++s) {
        if (2 * q.cell[s] + 1 >= GG_ORB_MAX_CELL) return gg_fail(GG_ERR_ARG, "gg_orb_split_wrap: children of PST cell %d exceed %d", q.cell[s], GG_ORB_MAX_CELL - 1);
        w.inactive[s] = fSplitInactive[s];
    }
    CK(cudaSetDevice(c->device));
    CK(gg_launch_orb_split_wrap(q, w, c->orbN, orb_pos(c, 0), orb_pos(c, 1), orb_pos(c, 2), (int *)c->ocell.p, c->st));
    ++c->nLaunches;
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_orb_fetch(gg_context *c, int *iCellOfParticle) {
    if (!c || c->orbN < 0 || !iCellOfParticle) return gg_fail(GG_ERR_ARG, "gg_orb_fetch: no particles loaded (gg_orb_load) or NULL argument");
    CK(cudaSetDevice(c->device));
    if (c->orbN > 0) CK(cudaMemcpyAsync(iCellOfParticle, c->ocell.p, sizeof(int) * (size_t)c->orbN, cudaMemcpyDefault, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

// ---- device-resident particle store: pkd->pStore kept in HBM across force evaluations ----

int gg_state_load(gg_context *c, int n, const double *x, const double *y, const double *z, const double *vx,
                  const double *vy, const double *vz, const double *fMass, const double *fSoft, const int *active,
                  double dt0) {
    if (!c || n < 1 || !x || !y || !z || !vx || !vy || !vz || !fMass || !fSoft)
        return gg_fail(GG_ERR_ARG, "gg_state_load: bad argument (n=%d)", n);
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = gg_finish_mom(c))) return rc;
    const size_t nb = sizeof(double) * (size_t)n;
    DevBuf *d8[] = {&c->sx, &c->sy, &c->sz, &c->sm, &c->sh, &c->sdt, &c->sdt2};
    for (DevBuf *b : d8)
        if ((rc = gg_ensure(c, *b, nb))) return rc;
    if ((rc = gg_ensure(c, c->svel, 3 * nb)) || (rc = gg_ensure(c, c->svel2, 3 * nb))) return rc;
    // sid: (persistent id, rung) pairs
    if ((rc = gg_ensure(c, c->sid, sizeof(int) * 2 * (size_t)n)) || (rc = gg_ensure(c, c->sid2, sizeof(int) * 2 * (size_t)n)) ||
        (rc = gg_ensure(c, c->sact, sizeof(int) * (size_t)n)))
        return rc;
    const double *src[] = {x, y, z, fMass, fSoft};
    DevBuf *dst[] = {&c->sx, &c->sy, &c->sz, &c->sm, &c->sh};
    for (int k = 0; k < 5; ++k) CK(cudaMemcpyAsync(dst[k]->p, src[k], nb, cudaMemcpyDefault, c->st));
    double *v = (double *)c->svel.p;
    CK(cudaMemcpyAsync(v, vx, nb, cudaMemcpyDefault, c->st));
    CK(cudaMemcpyAsync(v + n, vy, nb, cudaMemcpyDefault, c->st));
    CK(cudaMemcpyAsync(v + 2 * (size_t)n, vz, nb, cudaMemcpyDefault, c->st));
    if (active) CK(cudaMemcpyAsync(c->sact.p, active, sizeof(int) * (size_t)n, cudaMemcpyDefault, c->st));
    CK(gg_launch_state_init(n, (int *)c->sid.p, (double *)c->sdt.p, dt0, c->st));
    ++c->nLaunches;
    CK(cudaStreamSynchronize(c->st));
    c->stateN = n;
    c->stateHasActive = active != nullptr;
    c->stateDirty = true;
    c->stateForces = false;
    return GG_OK;
}

int gg_state_build(gg_context *c, int idSelf, int nBucket, double dTheta, int *pnNodes) {
    return gg_state_build_open(c, idSelf, nBucket, GG_OPEN_JOSH, dTheta, pnNodes);
}

int gg_state_build_open(gg_context *c, int idSelf, int nBucket, int iOpenType, double dTheta, int *pnNodes) {
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_build: no resident particles (gg_state_load)");
    double c23;
    int rco = open_factor("gg_state_build", iOpenType, &c23);
    if (rco) return rco;
    if (nBucket < 1 || nBucket > GG_MAX_BUCKET || !(dTheta > 0))
        return gg_fail(GG_ERR_ARG, "gg_state_build: nBucket=%d dTheta=%g", nBucket, dTheta);
    CK(cudaSetDevice(c->device));
    const int n = c->stateN;
    gg_particles pp{};
    pp.n = n;
    pp.x = (const double *)c->sx.p; pp.y = (const double *)c->sy.p; pp.z = (const double *)c->sz.p;
    pp.fMass = (const double *)c->sm.p; pp.fSoft = (const double *)c->sh.p;
    pp.active = c->stateHasActive ? (const int *)c->sact.p : nullptr;
    int rc = build_and_load(c, idSelf, &pp, nBucket, dTheta, c23, nullptr, pnNodes, nullptr);
    if (rc) return rc;
    // the store follows the tree order: velocities, ids and time steps by the build's permutation, the rest as built
    const GGBuiltDev &b = c->built;
    CK(gg_launch_permute(n, b.iorder, (const double *)c->svel.p, (double *)c->svel2.p, (const int *)c->sid.p,
                         (int *)c->sid2.p, (const double *)c->sdt.p, (double *)c->sdt2.p, c->st));
    ++c->nLaunches;
    std::swap(c->svel, c->svel2);
    std::swap(c->sid, c->sid2);
    std::swap(c->sdt, c->sdt2);
    const size_t nb = sizeof(double) * (size_t)n;
    CK(cudaMemcpyAsync(c->sx.p, b.x, nb, cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpyAsync(c->sy.p, b.y, nb, cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpyAsync(c->sz.p, b.z, nb, cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpyAsync(c->sm.p, b.m, nb, cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpyAsync(c->sh.p, b.h, nb, cudaMemcpyDeviceToDevice, c->st));
    if (b.active) CK(cudaMemcpyAsync(c->sact.p, b.active, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, c->st));
    c->stateDirty = false;
    c->stateForces = false;
    return GG_OK;
}

int gg_state_kick(gg_context *c, double dvFacOne, double dvFacTwo, const double *a) {
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_kick: no resident particles (gg_state_load)");
    CK(cudaSetDevice(c->device));
    const int n = c->stateN;
    const double *da = nullptr;
    if (a) { // the caller's accelerations, [n][3] in the store's current order
        int rc;
        if ((rc = gg_ensure(c, c->sacc, sizeof(double) * 3 * (size_t)n))) return rc;
        CK(cudaMemcpyAsync(c->sacc.p, a, sizeof(double) * 3 * (size_t)n, cudaMemcpyDefault, c->st));
        da = (const double *)c->sacc.p;
    } else {
        if (c->stateDirty || !c->stateForces)
            return gg_fail(GG_ERR_ARG, "gg_state_kick: no accelerations for the current particle order "
                                    "(sequence: gg_state_build, gg_gravity, gg_state_kick)");
        da = (const double *)c->acc.p;
    }
    CK(gg_launch_kick(n, (double *)c->svel.p, da, c->stateHasActive ? (const int *)c->sact.p : nullptr, dvFacOne,
                      dvFacTwo, c->st));
    ++c->nLaunches;
    if (a) CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_state_drift(gg_context *c, double dDelta, const double fCenter[3], int bPeriodic, const double fPeriod[3]) {
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_drift: no resident particles (gg_state_load)");
    if (bPeriodic && (!fCenter || !fPeriod)) return gg_fail(GG_ERR_ARG, "gg_state_drift: periodic drift needs fCenter, fPeriod");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = gg_finish_mom(c))) return rc; // the moment kernels of the last build read the positions
    if ((rc = gg_ensure(c, c->misc, 16 * sizeof(int)))) return rc;
    int *dOut = (int *)c->misc.p + 12;
    const double zero[3] = {0, 0, 0}, one[3] = {1, 1, 1};
    CK(cudaMemsetAsync(dOut, 0, sizeof(int), c->st));
    CK(gg_launch_drift(c->stateN, (double *)c->sx.p, (double *)c->sy.p, (double *)c->sz.p, (const double *)c->svel.p,
                       dDelta, bPeriodic ? fCenter : zero, bPeriodic, bPeriodic ? fPeriod : one, dOut, c->st));
    ++c->nLaunches;
    c->stateDirty = true;
    if (bPeriodic) { // the reference asserts every particle is back inside the box (pkd.c:3754-3756)
        int nOut = 0;
        CK(cudaMemcpyAsync(&nOut, dOut, sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        if (nOut) return gg_fail(GG_ERR_ARG, "gg_state_drift: %d particle(s) left the periodic box by more than one period", nOut);
    }
    return GG_OK;
}

int gg_state_gravstep(gg_context *c, double dEta, double *pdtMin) {
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_gravstep: no resident particles (gg_state_load)");
    if (c->stateDirty || !c->stateForces)
        return gg_fail(GG_ERR_ARG, "gg_state_gravstep: no dtGrav for the current particle order (gg_gravity first)");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = gg_ensure(c, c->misc, 16 * sizeof(int)))) return rc;
    unsigned long long *dMin = (unsigned long long *)((int *)c->misc.p + 14);
    int *dBad = (int *)c->misc.p + 13;
    CK(cudaMemsetAsync(dMin, 0xff, sizeof(unsigned long long), c->st));
    CK(cudaMemsetAsync(dBad, 0, sizeof(int), c->st));
    CK(gg_launch_gravstep(c->stateN, (double *)c->sdt.p, (const double *)c->dtg.p,
                          c->stateHasActive ? (const int *)c->sact.p : nullptr, dEta, dMin, dBad, c->st));
    ++c->nLaunches;
    unsigned long long bits = 0;
    int nBad = 0;
    CK(cudaMemcpyAsync(&bits, dMin, sizeof(bits), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(&nBad, dBad, sizeof(nBad), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (pdtMin) memcpy(pdtMin, &bits, sizeof(double));
    if (nBad) // the reference asserts (pkd.c:4616)
        return gg_fail(GG_ERR_ARG, "gg_state_gravstep: %d active particle(s) have dtGrav <= 0 (no interaction was evaluated for them)", nBad);
    return GG_OK;
}

int gg_state_fetch(gg_context *c, double *x, double *y, double *z, double *vx, double *vy, double *vz, int *id, double *dt) {
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_fetch: no resident particles (gg_state_load)");
    CK(cudaSetDevice(c->device));
    const int n = c->stateN;
    const size_t nb = sizeof(double) * (size_t)n;
    const double *v = (const double *)c->svel.p;
    struct { void *dst; const void *src; size_t bytes; } cp[] = {
        {x, c->sx.p, nb}, {y, c->sy.p, nb}, {z, c->sz.p, nb}, {vx, v, nb}, {vy, v + n, nb}, {vz, v + 2 * (size_t)n, nb},
        {dt, c->sdt.p, nb}};
    for (auto &e : cp)
        if (e.dst) CK(cudaMemcpyAsync(e.dst, e.src, e.bytes, cudaMemcpyDeviceToHost, c->st));
    if (id) CK(cudaMemcpy2DAsync(id, sizeof(int), c->sid.p, 2 * sizeof(int), sizeof(int), (size_t)n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_state_fetch_rungs(gg_context *c, int *rung, int *active) {
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_fetch_rungs: no resident particles (gg_state_load)");
    CK(cudaSetDevice(c->device));
    const int n = c->stateN;
    if (rung)
        CK(cudaMemcpy2DAsync(rung, sizeof(int), (const int *)c->sid.p + 1, 2 * sizeof(int), sizeof(int), (size_t)n,
                             cudaMemcpyDeviceToHost, c->st));
    if (active) {
        if (c->stateHasActive) CK(cudaMemcpyAsync(active, c->sact.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->st));
        else for (int i = 0; i < n; ++i) active[i] = 1;
    }
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_state_set_rungs(gg_context *c, const int *rung) {
    if (!c || c->stateN < 1 || !rung) return gg_fail(GG_ERR_ARG, "gg_state_set_rungs: bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy2DAsync((int *)c->sid.p + 1, 2 * sizeof(int), rung, sizeof(int), sizeof(int), (size_t)c->stateN,
                         cudaMemcpyDefault, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_state_init_dt(gg_context *c, double dDelta) {
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_init_dt: no resident particles (gg_state_load)");
    CK(cudaSetDevice(c->device));
    CK(gg_launch_init_dt(c->stateN, (double *)c->sdt.p, c->stateHasActive ? (const int *)c->sact.p : nullptr, dDelta, c->st));
    ++c->nLaunches;
    return GG_OK;
}

int gg_state_accelstep(gg_context *c, double dEta, double dVelFac, double dAccFac, int bEpsAcc, int bSqrtPhi) {
    (void)dVelFac; // the velocity enters only an assertion in the gravity-only build (pkd.c:4643)
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_accelstep: no resident particles (gg_state_load)");
    if (c->stateDirty || !c->stateForces)
        return gg_fail(GG_ERR_ARG, "gg_state_accelstep: no accelerations for the current particle order (gg_gravity first)");
    CK(cudaSetDevice(c->device));
    CK(gg_launch_accelstep(c->stateN, (double *)c->sdt.p, (const double *)c->acc.p, (const double *)c->pot.p,
                           (const double *)c->sh.p, c->stateHasActive ? (const int *)c->sact.p : nullptr, dEta, dAccFac,
                           bEpsAcc, bSqrtPhi, c->st));
    ++c->nLaunches;
    return GG_OK;
}

int gg_state_dt_to_rung(gg_context *c, int iRung, double dDelta, int iMaxRung, int bAll, int *pnMaxRung, int *piMaxRungIdeal,
                        int *piMaxRungOut) {
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_dt_to_rung: no resident particles (gg_state_load)");
    if (iRung < 0 || iMaxRung < 1 || iMaxRung > 127 || iRung + 1 > 127)
        return gg_fail(GG_ERR_UNSUPPORTED, "gg_state_dt_to_rung: iRung=%d iMaxRung=%d (supported < 128)", iRung, iMaxRung);
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = gg_ensure(c, c->srhist, 132 * sizeof(int)))) return rc;
    int *dh = (int *)c->srhist.p;
    CK(cudaMemsetAsync(dh, 0, 132 * sizeof(int), c->st));
    CK(gg_launch_dt_to_rung(c->stateN, (int *)c->sid.p, (const double *)c->sdt.p, iRung, dDelta, iMaxRung, bAll, dh, dh + 128,
                            dh + 130, c->st));
    ++c->nLaunches;
    int h[132];
    CK(cudaMemcpyAsync(h, dh, sizeof(h), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (h[130]) // the reference asserts both (pkd.c:4694, 4749)
        return gg_fail(GG_ERR_ARG, "gg_state_dt_to_rung: %d particle(s) have dt <= 0 or dDelta/dt >= 2.1e9", h[130]);
    int top = 0;
    for (int r = 127; r > 0; --r)
        if (h[r] > 0) { top = r; break; }
    if (piMaxRungOut) *piMaxRungOut = top;
    if (pnMaxRung) *pnMaxRung = h[top];
    if (piMaxRungIdeal) *piMaxRungIdeal = h[128];
    return GG_OK;
}

int gg_state_active_rung(gg_context *c, int iRung, int bGreater, int *pnActive) {
    if (!c || c->stateN < 1) return gg_fail(GG_ERR_ARG, "gg_state_active_rung: no resident particles (gg_state_load)");
    CK(cudaSetDevice(c->device));
    int rc;
    const int n = c->stateN;
    if ((rc = gg_ensure(c, c->sact, sizeof(int) * (size_t)n))) return rc;
    if ((rc = gg_ensure(c, c->srhist, 132 * sizeof(int)))) return rc;
    int *dCount = (int *)c->srhist.p + 129;
    CK(cudaMemsetAsync(dCount, 0, sizeof(int), c->st));
    CK(gg_launch_active_rung(n, (const int *)c->sid.p, (int *)c->sact.p, iRung, bGreater, dCount, c->st));
    ++c->nLaunches;
    c->stateHasActive = true;
    int nAct = 0;
    CK(cudaMemcpyAsync(&nAct, dCount, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (pnActive) *pnActive = nAct;
    // the loaded domain (same order as the store unless it moved since its build) evaluates the new active set
    if (!c->stateDirty && !c->dom.empty() && c->dom[0].nPart == n) {
        if ((rc = gg_set_active(c, (const int *)c->sact.p))) return rc;
    }
    return GG_OK;
}

// New ACTIVE flags for the domain that is already loaded (same tree, same particles): what msrActiveRung changes between
// two force evaluations on one tree (master.c:8403-8420).  Only the flags travel; the sink-bucket task list is rebuilt.
int gg_set_active(gg_context *c, const int *active) {
    if (!c || c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_set_active: no local domain (gg_set_local / gg_build_local)");
    CK(cudaSetDevice(c->device));
    const Domain &L = c->dom[0];
    const int np = L.nPart, nn = L.nNodes;
    int rc;
    const int *dActive = nullptr;
    c->stateForces = false; // a new sink set: the last evaluation's results do not cover the newly active particles
    if ((rc = drop_early_ewald(c))) return rc; // (an early Ewald correction was computed for the previous sink set)
    if (active) {
        if ((rc = gg_ensure(c, c->active, (size_t)(np + 1) * sizeof(int)))) return rc;
        CK(cudaMemcpyAsync(c->active.p, active, sizeof(int) * np, cudaMemcpyDefault, c->st));
        dActive = (const int *)c->active.p;
        c->hActive.resize((size_t)np);
        CK(cudaMemcpyAsync(c->hActive.data(), c->active.p, sizeof(int) * np, cudaMemcpyDeviceToHost, c->st));
    } else c->hActive.clear();
    int hCounts[2] = {0, 0};
    c->nPartUpload = np;
    if ((rc = build_task_list(c, nn, dActive, hCounts))) return rc;
    CK(cudaStreamSynchronize(c->st));
    c->nTasksLocal = hCounts[0];
    c->nBucketsLocal = hCounts[1];
    if (c->stateN > 0 && c->stateHasActive != (active != nullptr)) {
        // the resident store keeps its own copy of the flags (kick / grav-step read them)
        if (active) {
            if ((rc = gg_ensure(c, c->sact, sizeof(int) * (size_t)np))) return rc;
        }
        c->stateHasActive = active != nullptr;
    }
    if (c->stateN > 0 && active) {
        CK(cudaMemcpyAsync(c->sact.p, c->active.p, sizeof(int) * np, cudaMemcpyDeviceToDevice, c->st));
        CK(cudaStreamSynchronize(c->st));
    }
    return GG_OK;
}

// The root cell of the local (device-built) domain: what a multi-rank host hands to pstColCells / pstCalcRoot.
int gg_domain_summary(gg_context *c, double bnd[6], double r[3], double *fMass, double *fSoft, double *fOpen2,
                      double mom[GG_NMOM], double root[GG_NROOT]) {
    if (!c || !c->built.nNodes || c->dom.empty())
        return gg_fail(GG_ERR_ARG, "gg_domain_summary: no device-built local tree (gg_build_local)");
    CK(cudaSetDevice(c->device));
    int rc;
    c->rootLazy = true; // (re)read from the device records: gg_set_root_moments may have replaced c->root meanwhile
    double keep[GG_NROOT];
    const bool had = c->haveRoot;
    memcpy(keep, c->root, sizeof(keep));
    if ((rc = fetch_root_lazy(c))) return rc;
    if (root) memcpy(root, c->root, sizeof(c->root));
    if (had) memcpy(c->root, keep, sizeof(keep)); // a multi-rank root set by the host stays in force
    const int iRoot = c->dom[0].iRoot;
    double raw[32];
    NodeW w;
    CK(cudaMemcpyAsync(raw, (const double *)c->momraw.p + (size_t)iRoot * 32, sizeof(raw), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(&w, (const NodeW *)c->nodes.p + iRoot, sizeof(w), cudaMemcpyDeviceToHost, c->st));
    if (bnd) CK(cudaMemcpyAsync(bnd, c->built.bnd + 6 * (size_t)iRoot, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (r) { r[0] = w.rx; r[1] = w.ry; r[2] = w.rz; }
    if (fMass) *fMass = w.fMass;
    if (fSoft) *fSoft = w.fSoft;
    if (fOpen2) *fOpen2 = w.fOpen2;
    if (mom) {
        GGRawMom a;
        a.M = raw[0];
        for (int k = 0; k < 31; ++k) a.q[k] = raw[1 + k];
        gg_raw_reduce(a, mom);
    }
    return GG_OK;
}

// pkdCalcCell (pkd.c:2018-2135) over the whole local domain about the centre rcm -- one rank's contribution to an
// interior cell of the top tree (pstCalcCell, pst.c:3789): the moments by translating the root's raw moment record to
// rcm (exact binomial shift) and reducing it, Bmax by one pass over the particles.
int gg_domain_moments_about(gg_context *c, const double rcm[3], double mom[GG_NMOM], double *pBmax) {
    if (!c || !rcm || !mom || !pBmax || c->dom.empty() || !c->momraw.p)
        return gg_fail(GG_ERR_ARG, "gg_domain_moments_about: needs a local domain with device-formed moments");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = gg_finish_mom(c))) return rc;
    if ((rc = gg_ensure(c, c->misc, 16 * sizeof(int)))) return rc;
    unsigned long long *dMax = (unsigned long long *)((int *)c->misc.p + 14);
    const Domain &L = c->dom[0];
    CK(cudaMemsetAsync(dMax, 0, sizeof(unsigned long long), c->st));
    CK(gg_launch_bmax_about(L.nPart, (const PartS *)c->parts.p + L.partBase, rcm, dMax, c->st));
    ++c->nLaunches;
    double raw[32];
    NodeW w;
    unsigned long long bits = 0;
    CK(cudaMemcpyAsync(raw, (const double *)c->momraw.p + (size_t)L.iRoot * 32, sizeof(raw), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(&w, (const NodeW *)c->nodes.p + L.nodeBase + L.iRoot, sizeof(w), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(&bits, dMax, sizeof(bits), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    GGRawMom src, dst;
    src.M = raw[0];
    for (int k = 0; k < 31; ++k) src.q[k] = raw[1 + k];
    gg_raw_zero(dst);
    gg_raw_shift_add(dst, src, w.rx - rcm[0], w.ry - rcm[1], w.rz - rcm[2]);
    gg_raw_reduce(dst, mom);
    memcpy(pBmax, &bits, sizeof(double));
    return GG_OK;
}

int gg_build_info(gg_context *c, int *pnNodes, int *pnLevels, double *pmsBuild) {
    if (!c || !c->built.nNodes) return gg_fail(GG_ERR_ARG, "gg_build_info: no device-built tree");
    if (pnNodes) *pnNodes = c->built.nNodes;
    if (pnLevels) *pnLevels = c->built.nLevels;
    if (pmsBuild) *pmsBuild = c->msBuild;
    return GG_OK;
}

int gg_tree_fetch(gg_context *c, double *bnd, double *r, double *fMass, double *fSoft, double *fOpen2, double *mom,
                  int *pLower, int *pUpper, int *iLower, int *iUpper, double *x, double *y, double *z, double *m,
                  double *h, int *active) {
    if (!c || !c->built.nNodes) return gg_fail(GG_ERR_ARG, "gg_tree_fetch: no device-built tree (gg_build_local)");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = gg_finish_mom(c))) return rc;
    const GGBuiltDev &b = c->built;
    const size_t nn = (size_t)b.nNodes, np = (size_t)b.nPart;
    struct { void *dst; const void *src; size_t bytes; } cp[] = {
        {bnd, b.bnd, 48 * nn}, {r, b.r, 24 * nn}, {fMass, b.fMass, 8 * nn}, {fSoft, b.fSoft, 8 * nn},
        {fOpen2, b.fOpen2, 8 * nn}, {pLower, b.pLower, 4 * nn}, {pUpper, b.pUpper, 4 * nn}, {iLower, b.iLower, 4 * nn},
        {iUpper, b.iUpper, 4 * nn}, {x, b.x, 8 * np}, {y, b.y, 8 * np}, {z, b.z, 8 * np}, {m, b.m, 8 * np}, {h, b.h, 8 * np},
        {active, b.active, 4 * np}};
    for (auto &e : cp)
        if (e.dst && e.src) CK(cudaMemcpyAsync(e.dst, e.src, e.bytes, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (mom) { // the reduced multipoles the device formed: raw records reduced by a kernel, then one copy
        int rc2;
        if ((rc2 = gg_ensure(c, c->momout, sizeof(double) * GG_NMOM * nn))) return rc2;
        CK(gg_launch_mom_reduce((int)nn, (const double *)c->momraw.p, (double *)c->momout.p, c->st));
        ++c->nLaunches;
        CK(cudaMemcpyAsync(mom, c->momout.p, sizeof(double) * GG_NMOM * nn, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
    }
    return GG_OK;
}

int gg_tree_fetch_build(gg_context *c, int *iDim, double *fSplit, double *fBmax) {
    if (!c || !c->built.nNodes) return gg_fail(GG_ERR_ARG, "gg_tree_fetch_build: no device-built tree (gg_build_local)");
    CK(cudaSetDevice(c->device));
    const GGBuiltDev &b = c->built;
    const size_t nn = (size_t)b.nNodes;
    if (iDim) CK(cudaMemcpyAsync(iDim, b.iDim, 4 * nn, cudaMemcpyDeviceToHost, c->st));
    if (fSplit) CK(cudaMemcpyAsync(fSplit, b.fSplit, 8 * nn, cudaMemcpyDeviceToHost, c->st));
    if (fBmax) CK(cudaMemcpyAsync(fBmax, b.fBmax, 8 * nn, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_set_remote(gg_context *c, int id, const gg_tree *t, const gg_particles *pp, int bDevice) {
    if (!c || !t || !pp) return gg_fail(GG_ERR_ARG, "gg_set_remote: null argument");
    if (c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_set_remote: call gg_set_local first");
    if (id == c->idSelf) return gg_fail(GG_ERR_ARG, "gg_set_remote: id %d is the local domain", id);
    CK(cudaSetDevice(c->device));
    // (bucket sizes and particle ranges are validated on the device while the records are packed)
    int rc = upload_domain(c, t, pp, c->nNodesAll, c->nPartAll, false, bDevice != 0);
    if (rc) return rc;
    c->dom.push_back(Domain{id, t->nNodes, pp->n, t->iRoot, c->nNodesAll, c->nPartAll});
    c->nNodesAll += t->nNodes;
    c->nPartAll += pp->n;
    return GG_OK;
}

int gg_let_export(gg_context *c, int nRemote, const double *bnd, const gg_params *prm, void **pDev, size_t *offsets,
                  int *hdr) {
    if (!c || !bnd || !prm || !pDev || !offsets || !hdr || c->dom.empty())
        return gg_fail(GG_ERR_ARG, "gg_let_export: bad argument / no local domain");
    int rc = gg_let_export_impl(c, nRemote, bnd, prm, offsets, hdr);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->st));
    *pDev = c->letout.p;
    return GG_OK;
}

} // extern "C"

// The export proper; the packing kernel is left in flight on c->st (the sizes are already known to the host).
int gg_let_export_impl(gg_context *c, int nRemote, const double *bnd, const gg_params *prm, size_t *offsets, int *hdr) {
    if (nRemote < 1 || nRemote > 15) return gg_fail(GG_ERR_UNSUPPORTED, "gg_let_export: nRemote=%d (1..15)", nRemote);
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = gg_finish_mom(c))) return rc;
    const Domain &L = c->dom[0];
    const int nn = L.nNodes;
    if (nn >= (1 << 28)) return gg_fail(GG_ERR_UNSUPPORTED, "gg_let_export: %d nodes", nn);
    Images im = make_images(prm);
    if (im.n > GG_MAX_IMAGES) return gg_fail(GG_ERR_UNSUPPORTED, "gg_let_export: %d images", im.n);
    const size_t seg = (size_t)nn + 1, nAll = seg * nRemote + 1;
    if (nAll >= 0x7fffffffull) return gg_fail(GG_ERR_UNSUPPORTED, "gg_let_export: %d nodes x %d remotes", nn, nRemote);
    if ((rc = gg_ensure(c, c->imgoff, im.off.size() * sizeof(double)))) return rc;
    if ((rc = gg_ensure(c, c->letflag, (size_t)nRemote * nn))) return rc;
    if ((rc = gg_ensure(c, c->letfront, (size_t)2 * nRemote * nn * sizeof(unsigned)))) return rc;
    if ((rc = gg_ensure(c, c->letidx, 4 * nAll * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->letmisc, 64 * sizeof(double) * 2 + 64 + 40 * sizeof(int)))) return rc;
    double *dBnd = (double *)c->letmisc.p;
    int *dCount = (int *)(dBnd + 6 * 16);
    int *dBounds = dCount + 4;
    CK(cudaMemcpyAsync(c->imgoff.p, im.off.data(), im.off.size() * sizeof(double), cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(dBnd, bnd, sizeof(double) * 6 * nRemote, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemsetAsync(c->letflag.p, 0, (size_t)nRemote * nn, c->st));
    unsigned *f0 = (unsigned *)c->letfront.p, *f1 = f0 + (size_t)nRemote * nn;
    k_let_seed<<<1, 32, 0, c->st>>>(f0, dCount, nRemote, L.iRoot);
    CK(cudaGetLastError());
    LetArgs la;
    la.nodes = (const NodeW *)c->nodes.p; la.nNodes = nn; la.nRemote = nRemote; la.nImages = im.n;
    la.imgOff = (const double *)c->imgoff.p; la.bnd = dBnd; la.flag = (unsigned char *)c->letflag.p; la.count = dCount;
    const int grid = c->nSM * 8;
    for (int level = 0;; ++level) { // one launch per tree level; the frontier size is read back every 8 levels
        la.front = (level & 1) ? f1 : f0;
        la.next = (level & 1) ? f0 : f1;
        k_let_level<<<grid, 256, 0, c->st>>>(la, level);
        CK(cudaGetLastError());
        k_let_reset<<<1, 1, 0, c->st>>>(dCount, level + 1);
        c->nLaunches += 2;
        if ((level & 7) == 7) {
            int cnt[2];
            CK(cudaMemcpyAsync(cnt, dCount, sizeof(cnt), cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            if (cnt[(level + 1) & 1] == 0) break;
        }
        if (level > 4096) return gg_fail(GG_ERR_UNSUPPORTED, "gg_let_export: tree deeper than 4096 levels");
    }
    // ---- compaction of all remotes at once: new node numbers, new particle offsets, sizes
    int *keep = (int *)c->letidx.p, *npart = keep + nAll, *newIdx = npart + nAll, *newPart = newIdx + nAll;
    size_t tmpBytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, keep, newIdx, (int)nAll, c->st));
    if ((rc = gg_ensure(c, c->cubtmp, tmpBytes))) return rc;
    k_let_counts<<<(unsigned)((nAll + 255) / 256), 256, 0, c->st>>>(nn, nRemote, (const NodeW *)c->nodes.p,
                                                                  (const unsigned char *)c->letflag.p, keep, npart);
    CK(cudaGetLastError());
    CK(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmpBytes, keep, newIdx, (int)nAll, c->st));
    CK(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmpBytes, npart, newPart, (int)nAll, c->st));
    k_let_bounds<<<1, 32, 0, c->st>>>(nn, nRemote, newIdx, newPart, dBounds);
    CK(cudaGetLastError());
    int hb[34];
    CK(cudaMemcpyAsync(hb, dBounds, sizeof(int) * 2 * (nRemote + 1), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    c->nLaunches += 4;
    LetPackArgs pa;
    memset(&pa, 0, sizeof(pa));
    size_t off = 0;
    for (int r = 0; r < nRemote; ++r) {
        const int nOut = hb[2 * (r + 1)] - hb[2 * r], nOutP = hb[2 * (r + 1) + 1] - hb[2 * r + 1];
        offsets[r] = off;
        off += (size_t)nOut * (sizeof(NodeW) + 128 + 48) + (size_t)nOutP * sizeof(PartS);
        off = (off + 255) & ~(size_t)255;
        hdr[3 * r] = nOut; hdr[3 * r + 1] = nOutP; hdr[3 * r + 2] = 0; // pre-order: the root stays first
        pa.baseIdx[r] = hb[2 * r]; pa.basePart[r] = hb[2 * r + 1];
    }
    offsets[nRemote] = off;
    if ((rc = gg_ensure(c, c->letout, off + 256))) return rc;
    for (int r = 0; r < nRemote; ++r) {
        char *o = (char *)c->letout.p + offsets[r];
        const size_t nOut = (size_t)hdr[3 * r];
        pa.oNodes[r] = (NodeW *)o;
        pa.oMomf[r] = (float4 *)(o + nOut * sizeof(NodeW));
        pa.oMomq[r] = (double *)((char *)pa.oMomf[r] + nOut * 128);
        pa.oParts[r] = (PartS *)((char *)pa.oMomq[r] + nOut * 48);
    }
    pa.nn = nn; pa.nRemote = nRemote;
    pa.nodes = (const NodeW *)c->nodes.p; pa.momf = (const float4 *)c->momf.p; pa.momq = (const double *)c->momq.p;
    pa.parts = (const PartS *)c->parts.p; pa.flag = (const unsigned char *)c->letflag.p;
    pa.newIdx = newIdx; pa.newPart = newPart;
    const size_t nThreads = (size_t)nn * nRemote * 8;
    k_let_pack<<<(unsigned)((nThreads + 255) / 256), 256, 0, c->st>>>(pa);
    CK(cudaGetLastError());
    ++c->nLaunches;
    return GG_OK;
}

extern "C" {

int gg_export_size(gg_context *c, size_t *bytes, int hdr[3]) {
    if (!c || !bytes || !hdr || c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_export_size: no local domain");
    const Domain &L = c->dom[0];
    hdr[0] = L.nNodes; hdr[1] = L.nPart; hdr[2] = L.iRoot;
    *bytes = (size_t)L.nNodes * (sizeof(NodeW) + 128 + 48) + (size_t)L.nPart * sizeof(PartS);
    return GG_OK;
}

int gg_export_local(gg_context *c, void *dst) {
    if (!c || !dst || c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_export_local: no local domain");
    CK(cudaSetDevice(c->device));
    { int rc0 = gg_finish_mom(c); if (rc0) return rc0; }
    const Domain &L = c->dom[0];
    char *o = (char *)dst;
    const size_t nn = (size_t)L.nNodes, np = (size_t)L.nPart;
    CK(cudaMemcpyAsync(o, c->nodes.p, nn * sizeof(NodeW), cudaMemcpyDeviceToDevice, c->st)); o += nn * sizeof(NodeW);
    CK(cudaMemcpyAsync(o, c->momf.p, nn * 128, cudaMemcpyDeviceToDevice, c->st)); o += nn * 128;
    CK(cudaMemcpyAsync(o, c->momq.p, nn * 48, cudaMemcpyDeviceToDevice, c->st)); o += nn * 48;
    CK(cudaMemcpyAsync(o, c->parts.p, np * sizeof(PartS), cudaMemcpyDeviceToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_set_remote_packed(gg_context *c, int id, const int hdr[3], const void *src) {
    if (!c || !hdr || !src) return gg_fail(GG_ERR_ARG, "gg_set_remote_packed: null argument");
    int rc = gg_ingest_packed(c, id, hdr, src);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->st)); // the source buffer belongs to the caller (next all-gather may overwrite it)
    return GG_OK;
}

} // extern "C"

int gg_ingest_packed(gg_context *c, int id, const int hdr[3], const void *src) {
    if (c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_set_remote_packed: call gg_set_local first");
    if (id == c->idSelf) return gg_fail(GG_ERR_ARG, "gg_set_remote_packed: id %d is the local domain", id);
    const int nn = hdr[0], np = hdr[1], iRoot = hdr[2];
    if (nn < 1 || np < 0 || iRoot < 0 || iRoot >= nn) return gg_fail(GG_ERR_ARG, "gg_set_remote_packed: nNodes=%d n=%d iRoot=%d", nn, np, iRoot);
    CK(cudaSetDevice(c->device));
    const size_t keepN = (size_t)c->nNodesAll, keepP = (size_t)c->nPartAll;
    int rc;
    if ((rc = gg_finish_mom(c))) return rc;
    if ((rc = gg_ensure(c, c->nodes, (keepN + nn + GG_MAX_TOP) * sizeof(NodeW), keepN * sizeof(NodeW)))) return rc;
    if ((rc = gg_ensure(c, c->momf, (keepN + nn + GG_MAX_TOP) * 128, keepN * 128))) return rc;
    if ((rc = gg_ensure(c, c->momq, (keepN + nn + GG_MAX_TOP) * 48, keepN * 48))) return rc;
    if ((rc = gg_ensure(c, c->parts, (keepP + np + 1) * sizeof(PartS), keepP * sizeof(PartS)))) return rc;
    const char *i0 = (const char *)src;
    const char *i1 = i0 + (size_t)nn * sizeof(NodeW), *i2 = i1 + (size_t)nn * 128, *i3 = i2 + (size_t)nn * 48;
    k_rebase_nodes<<<(nn + 255) / 256, 256, 0, c->st>>>(nn, (const NodeW *)i0, (int)keepN, (int)keepP, (NodeW *)c->nodes.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync((char *)c->momf.p + keepN * 128, i1, (size_t)nn * 128, cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpyAsync((char *)c->momq.p + keepN * 48, i2, (size_t)nn * 48, cudaMemcpyDeviceToDevice, c->st));
    if (np > 0) CK(cudaMemcpyAsync((PartS *)c->parts.p + keepP, i3, (size_t)np * sizeof(PartS), cudaMemcpyDeviceToDevice, c->st));
    c->dom.push_back(Domain{id, nn, np, iRoot, (int)keepN, (int)keepP});
    c->nNodesAll += nn;
    c->nPartAll += np;
    // the device records do not say how large the owner's buckets are; GG_MAX_BUCKET bounds them on every rank
    return GG_OK;
}

extern "C" {

int gg_clear_remote(gg_context *c) {
    if (!c || c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_clear_remote: no local domain");
    c->dom.resize(1);
    c->nNodesAll = c->dom[0].nNodes;
    c->nPartAll = c->dom[0].nPart;
    c->nTop = 0;
    return GG_OK;
}

int gg_set_top(gg_context *c, int nCell, const int *pLower, const int *bUsed, const double *r, const double *fMass,
               const double *fSoft, const double *fOpen2, const double *mom) {
    if (!c || nCell < 2 || !pLower || !bUsed || !r || !fMass || !fSoft || !fOpen2 || !mom)
        return gg_fail(GG_ERR_ARG, "gg_set_top: bad argument");
    if (nCell > GG_MAX_TOP) return gg_fail(GG_ERR_UNSUPPORTED, "gg_set_top: nCell=%d > %d", nCell, GG_MAX_TOP);
    c->nTop = nCell;
    c->topLower.assign(pLower, pLower + nCell);
    c->topUsed.assign(bUsed, bUsed + nCell);
    c->topR.assign(r, r + 3 * (size_t)nCell);
    c->topMass.assign(fMass, fMass + nCell);
    c->topSoft.assign(fSoft, fSoft + nCell);
    c->topOpen2.assign(fOpen2, fOpen2 + nCell);
    c->topMom.assign(mom, mom + (size_t)GG_NMOM * nCell);
    return GG_OK;
}

int gg_announce(gg_context *c, const gg_params *prm) {
    if (!c) return gg_fail(GG_ERR_ARG, "gg_announce: null context");
    c->annValid = prm != nullptr;
    if (prm) c->ann = *prm;
    return GG_OK;
}

int gg_set_root_moments(gg_context *c, const double root[GG_NROOT]) {
    if (!c || !root) return gg_fail(GG_ERR_ARG, "gg_set_root_moments: null");
    memcpy(c->root, root, sizeof(c->root));
    c->haveRoot = true;
    c->rootLazy = false;
    return GG_OK;
}

} // extern "C"

namespace {

// Resolve a heap cell of the top tree to a global node index; interior cells live at topBase + i.
int map_top(const gg_context *c, int i, int topBase) {
    if (c->topLower[i] >= 0) {
        for (const Domain &d : c->dom)
            if (d.id == c->topLower[i]) return d.nodeBase + d.iRoot;
        return -2;
    }
    return topBase + i;
}

int pack_top(gg_context *c, int *pRoot) {
    const int topBase = c->nNodesAll;
    const int n = c->nTop;
    std::vector<NodeW> w(n);
    std::vector<float> mf((size_t)n * 32, 0.f);
    std::vector<double> mq((size_t)n * 6, 0.0);
    for (int i = 1; i < n; ++i) {
        if (!c->topUsed[i] || c->topLower[i] >= 0) continue;
        NodeW &o = w[i];
        o.rx = c->topR[3 * i]; o.ry = c->topR[3 * i + 1]; o.rz = c->topR[3 * i + 2];
        o.fOpen2 = c->topOpen2[i]; o.fSoft = c->topSoft[i]; o.fMass = c->topMass[i];
        if (2 * i + 1 >= n) return gg_fail(GG_ERR_ARG, "gg_set_top: interior cell %d has no children in the heap", i);
        o.c0 = map_top(c, 2 * i, topBase);
        o.c1 = map_top(c, 2 * i + 1, topBase);
        if (o.c0 < -1 || o.c1 < -1)
            return gg_fail(GG_ERR_ARG, "gg_set_top: a top leaf names a rank whose domain was not loaded");
        o.pLower = 0;
        o.nP = 1 << 30; // exempt from the "< 4 particles" rule: the top walk has no such test (walk.c:363-371)
        const double *q = &c->topMom[(size_t)GG_NMOM * i];
        gg_pack_momf(q, &mf[(size_t)i * 32]);
        for (int k = 0; k < 6; ++k) mq[(size_t)i * 6 + k] = q[k];
    }
    int rc;
    const size_t keep = (size_t)topBase;
    if ((rc = gg_ensure(c, c->nodes, (keep + n + 1) * sizeof(NodeW), keep * sizeof(NodeW)))) return rc;
    if ((rc = gg_ensure(c, c->momf, (keep + n + 1) * 128, keep * 128))) return rc;
    if ((rc = gg_ensure(c, c->momq, (keep + n + 1) * 48, keep * 48))) return rc;
    CK(cudaMemcpyAsync((NodeW *)c->nodes.p + topBase, w.data(), sizeof(NodeW) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync((char *)c->momf.p + keep * 128, mf.data(), 128 * (size_t)n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync((char *)c->momq.p + keep * 48, mq.data(), 48 * (size_t)n, cudaMemcpyHostToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    *pRoot = map_top(c, 1, topBase);
    if (*pRoot < 0) return gg_fail(GG_ERR_ARG, "gg_set_top: root of the top tree cannot be resolved");
    return GG_OK;
}

// pkdEwaldInit (ewald.c:182-248) + the constants of pkdBucketEwald (ewald.c:30-44) for the device kernel: the k-space
// table goes to c->ewt, everything else travels as kernel arguments.  parts / acc / pot / nLoop = the local domain's.
int make_ewald_args(gg_context *c, const gg_params *prm, EwaldKernelArgs &ea, cudaStream_t st) {
    std::vector<double> ewt;
    const double Lbox = prm->fPeriod[0];
    gg_ewald_table_host(c->root, Lbox, prm->fEwhCut, prm->iEwOrder, ewt);
    int rc;
    if ((rc = gg_ensure(c, c->ewt, (ewt.size() + 8) * sizeof(double)))) return rc;
    // (pageable source of a few KB: the runtime stages it before the call returns, so the local vector may go)
    CK(cudaMemcpyAsync(c->ewt.p, ewt.data(), ewt.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    memset(&ea, 0, sizeof(ea));
    ea.parts = (const PartS *)c->parts.p;
    memcpy(ea.root, c->root, sizeof(ea.root));
    const double *R = c->root;
    ea.trQ4[0] = R[20] + R[26] + R[29]; // Qxx = xxxx + xxyy + xxzz   (meval.h:36-41)
    ea.trQ4[1] = R[22] + R[21] + R[30]; // Qxy = xxxy + xyyy + xyzz
    ea.trQ4[2] = R[24] + R[28] + R[31]; // Qxz = xxxz + xyyz + xzzz
    ea.trQ4[3] = R[26] + R[23] + R[32]; // Qyy = xxyy + yyyy + yyzz
    ea.trQ4[4] = R[27] + R[25] + R[33]; // Qyz = xxyz + yyyz + yzzz
    ea.trQ4[5] = R[29] + R[32] + R[34]; // Qzz = xxzz + yyzz + zzzz
    ea.trQ4[6] = (1.0 / 8.0) * (ea.trQ4[0] + ea.trQ4[3] + ea.trQ4[5]);
    ea.trQ3[0] = 0.5 * (R[10] + R[11] + R[17]); // Qx = xxx + xyy + xzz   (meval.h:55-57)
    ea.trQ3[1] = 0.5 * (R[12] + R[13] + R[18]); // Qy = xxy + yyy + yzz
    ea.trQ3[2] = 0.5 * (R[14] + R[15] + R[19]); // Qz = xxz + yyz + zzz
    ea.trQ2 = 0.5 * (R[4] + R[5] + R[9]);
    ea.ewt = (const double *)c->ewt.p;
    ea.nEwh = (int)(ewt.size() / 5);
    ea.nReps = prm->nReps;
    ea.nEwReps = (int)ceil(prm->fEwCut);
    if (prm->nReps > ea.nEwReps) ea.nEwReps = prm->nReps;
    ea.iOrder = prm->iEwOrder;
    ea.L = Lbox;
    ea.fEwCut2 = prm->fEwCut * prm->fEwCut * Lbox * Lbox;
    ea.alpha = 2.0 / Lbox;
    ea.alpha2 = ea.alpha * ea.alpha;
    ea.k1 = M_PI / (ea.alpha2 * Lbox * Lbox * Lbox);
    ea.ka = 2.0 * ea.alpha / sqrt(M_PI);
    ea.acc = (double *)c->acc.p;
    ea.pot = (double *)c->pot.p;
    ea.nLoop = (int *)c->nloop.p;
    return GG_OK;
}

int run_gravity(gg_context *c, const gg_params *prm, const Task *singleTask, gg_stats *stats, int depth = 0) {
    if (c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_gravity: gg_set_local has not been called");
    if (prm->iOrder < 1 || prm->iOrder > 4 || prm->iEwOrder < 0 || prm->iEwOrder > 4)
        return gg_fail(GG_ERR_UNSUPPORTED, "gg_gravity: iOrder=%d iEwOrder=%d (supported 1..4)", prm->iOrder, prm->iEwOrder);
    if (prm->nReps < 0 || prm->nReps > 5) return gg_fail(GG_ERR_UNSUPPORTED, "gg_gravity: nReps=%d (supported 0..5)", prm->nReps);
    CK(cudaSetDevice(c->device));
    if (depth > 0) CK(cudaStreamSynchronize(c->st3)); // a re-run: the previous attempt's k_stats may still be in flight
    const Domain &L = c->dom[0];
    const int n = L.nPart, nn = L.nNodes;
    const bool doEwald = prm->bPeriodic && prm->bEwald && prm->iEwOrder > 0 && !(prm->flags & GG_FLAG_WALK_ONLY);
    if (doEwald && c->rootLazy) {
        int rcr = fetch_root_lazy(c);
        if (rcr) return rcr;
    }
    if (doEwald && !c->haveRoot) return gg_fail(GG_ERR_ARG, "gg_gravity: Ewald needs gg_set_root_moments");
    Trace tr;
    Images im = make_images(prm);
    if (im.n > GG_MAX_IMAGES) return gg_fail(GG_ERR_UNSUPPORTED, "gg_gravity: %d images", im.n);
    int rootNode = L.iRoot;
    int nNodesAll = c->nNodesAll;
    int rc;
    if (c->nTop > 0 || c->dom.size() > 1) { // several domains: buffers may be re-allocated below -- no overlap
        if ((rc = gg_finish_mom(c))) return rc;
    }
    if (c->nTop > 0) {
        if ((rc = pack_top(c, &rootNode))) return rc;
        nNodesAll += c->nTop;
    } else if (c->dom.size() > 1)
        return gg_fail(GG_ERR_ARG, "gg_gravity: remote domains are loaded but gg_set_top was not called");
    const unsigned cap = (1u << (32 - im.bits)) - 2u;
    if ((unsigned)nNodesAll > cap || (unsigned)c->nPartAll > cap)
        return gg_fail(GG_ERR_UNSUPPORTED, "gg_gravity: %d nodes / %d particles exceed the %u addressable with %d images",
                    nNodesAll, c->nPartAll, cap, im.n);
    c->nLaunches = 0;
    const int *dActive = c->hActive.empty() ? nullptr : (const int *)c->active.p;

    if ((rc = gg_ensure(c, c->imgoff, im.off.size() * sizeof(double)))) return rc;
    CK(cudaMemcpyAsync(c->imgoff.p, im.off.data(), im.off.size() * sizeof(double), cudaMemcpyHostToDevice, c->st));
    // bDoSun pass: the dummy sink lives in the spare slots BEHIND everything that is loaded -- particle sunP, node sunN
    // (behind the remote domains and the top cells when there are any) -- and the per-particle / per-node result
    // arrays, indexed by the same numbers, grow to cover them while keeping the main pass's results
    const int sunP = c->nPartAll, sunN = nNodesAll;
    const size_t nRes = c->sunMode ? (size_t)sunP + 1 : (size_t)n + 1, nCnt = c->sunMode ? (size_t)sunN + 1 : (size_t)nn + 1;
    const bool keepRes = c->sunMode;
    if ((rc = gg_ensure(c, c->counts, nCnt * 3 * sizeof(int), keepRes ? (size_t)nn * 3 * sizeof(int) : 0))) return rc;
    if ((rc = gg_ensure(c, c->acc, nRes * 3 * sizeof(double), keepRes ? (size_t)n * 3 * sizeof(double) : 0))) return rc;
    if ((rc = gg_ensure(c, c->pot, nRes * sizeof(double), keepRes ? (size_t)n * sizeof(double) : 0))) return rc;
    if ((rc = gg_ensure(c, c->dtg, nRes * sizeof(double), keepRes ? (size_t)n * sizeof(double) : 0))) return rc;
    if ((rc = gg_ensure(c, c->fweight, nRes * sizeof(double), keepRes ? (size_t)n * sizeof(double) : 0))) return rc;
    if ((rc = gg_ensure(c, c->nloop, (size_t)(n + 1) * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->sums, 16 * sizeof(unsigned long long)))) return rc;
    if ((rc = gg_ensure(c, c->misc, 16 * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->ngroups, (size_t)(nn + 1) * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->goffs, (size_t)(nn + 1) * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->isb, (size_t)(nn + 1) * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->boffs, (size_t)(nn + 1) * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->bnode, (size_t)(nn + 1) * sizeof(int)))) return rc;

    // the Ewald correction gg_set_local launched early (gg_announce): same parameters, same root expansion, same sink set,
    // first evaluation of this upload -> acc / pot / nloop already hold it (or will, on the side stream)
    const bool usePre = doEwald && !singleTask && !c->sunMode && depth == 0 && c->ewValid && c->ewN == n &&
                        same_ewald_params(c->ewPrm, *prm) && memcmp(c->ewRoot, c->root, sizeof(c->ewRoot)) == 0;
    if (c->ewPending) { // either way the side stream's work is ordered before what follows on the main stream
        CK(cudaStreamWaitEvent(c->st, c->evEw[2], 0));
        c->ewPending = false;
    }
    c->ewValid = false; // consumed (or stale): a second evaluation of the same upload computes the correction itself
    CK(cudaEventRecord(c->ev[0], c->st));
    if (c->sunMode) { // only the dummy sink's slot and its bucket's counters: everything else holds results
        CK(cudaMemsetAsync((int *)c->counts.p + 3 * (size_t)sunN, 0xff, 3 * sizeof(int), c->st));
        CK(cudaMemsetAsync((double *)c->acc.p + 3 * (size_t)sunP, 0, 3 * sizeof(double), c->st));
        CK(cudaMemsetAsync((double *)c->pot.p + sunP, 0, sizeof(double), c->st));
        CK(cudaMemsetAsync((double *)c->dtg.p + sunP, 0, sizeof(double), c->st));
        CK(cudaMemsetAsync((double *)c->fweight.p + sunP, 0, sizeof(double), c->st));
    } else {
        CK(cudaMemsetAsync(c->counts.p, 0xff, (size_t)nn * 3 * sizeof(int), c->st));
        if (!usePre) {
            CK(cudaMemsetAsync(c->acc.p, 0, (size_t)n * 3 * sizeof(double), c->st));
            CK(cudaMemsetAsync(c->pot.p, 0, (size_t)n * sizeof(double), c->st));
        }
        CK(cudaMemsetAsync(c->dtg.p, 0, (size_t)n * sizeof(double), c->st));
        CK(cudaMemsetAsync(c->fweight.p, 0, (size_t)n * sizeof(double), c->st));
    }
    CK(cudaMemsetAsync(c->sums.p, 0, 16 * sizeof(unsigned long long), c->st));
    CK(cudaMemsetAsync(c->misc.p, 0, 16 * sizeof(int), c->st));

    // ---- Ewald correction and comoving background FIRST: they need only the particles and the root moments, and
    //      with them already in acc/pot the list evaluation's store is the final one (zero-copy delivery below)
    CK(cudaEventRecord(c->ev[7], c->st));
    int nEwh = 0;
    if (usePre) nEwh = c->ewNEwh;
    else if (doEwald && !singleTask) {
        EwaldKernelArgs ea;
        if ((rc = make_ewald_args(c, prm, ea, c->st))) return rc;
        nEwh = ea.nEwh;
        ea.active = dActive;
        ea.n = n;
        CK(gg_launch_ewald_kernel(ea, c->st));
        ++c->nLaunches;
    }
    CK(cudaEventRecord(c->ev[3], c->st));
    // (not in the bDoSun / single-bucket passes: the reference adds the term once per bucket of the main loop,
    //  pkd.c:2967-2991, before the Sun pass)
    if (prm->bComove && !prm->bPeriodic && !(prm->flags & GG_FLAG_WALK_ONLY) && n > 0 && !singleTask) {
        k_comove<<<(n + 255) / 256, 256, 0, c->st>>>(n, (const PartS *)c->parts.p, dActive, prm->dRhoFac,
                                                     (double *)c->acc.p, (double *)c->pot.p);
        CK(cudaGetLastError());
        ++c->nLaunches;
    }
    // ---- task list: local buckets with an active sink, in tree order (+ their ordinals: walk groups)
    int nTasks = 0, nBuckets = 0;
    if (singleTask) {
        // (its own small buffer: the task list of gg_set_local stays intact for the next gg_gravity)
        if ((rc = gg_ensure(c, c->dbgtask, sizeof(Task) + sizeof(int)))) return rc;
        CK(cudaMemcpyAsync(c->dbgtask.p, singleTask, sizeof(Task), cudaMemcpyHostToDevice, c->st));
        CK(cudaMemcpyAsync((char *)c->dbgtask.p + sizeof(Task), &singleTask->node, sizeof(int), cudaMemcpyHostToDevice, c->st));
        nTasks = nBuckets = 1;
    } else {
        nTasks = c->nTasksLocal;   // built by gg_set_local (build_task_list)
        nBuckets = c->nBucketsLocal;
    }
    const int nWalkGroups = (nBuckets + GG_WALK_GB - 1) / GG_WALK_GB;
    if ((rc = gg_ensure(c, c->ghead, (size_t)(nWalkGroups + 1) * 2 * GG_NLIST * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->gcnt, (size_t)(nWalkGroups + 1) * 2 * GG_NLIST * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->bcnt, (size_t)(nBuckets + 1) * GG_NLIST * sizeof(int)))) return rc;
    if ((rc = gg_ensure(c, c->btot, (size_t)(nBuckets + 1) * sizeof(long long)))) return rc;
    if ((rc = gg_ensure(c, c->boff64, (size_t)(nBuckets + 1) * sizeof(long long)))) return rc;
    c->nTasks = nTasks;

    // ---- walk (lists -> HBM pool) + list evaluation
    const bool walkOnly = (prm->flags & GG_FLAG_WALK_ONLY) != 0;
    if (!walkOnly && c->capBlocks == 0) { // first guess: ~700 list entries per bucket, plus one slab per resident warp
        c->capBlocks = (size_t)nBuckets * 16 + (size_t)c->nSM * 64 * GG_SLAB_BLOCKS + 1024;
    }
    TreeKernelArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.nodes = (const NodeW *)c->nodes.p;
    ta.momf = (const float4 *)c->momf.p;
    ta.momq = (const double *)c->momq.p;
    ta.parts = (const PartS *)c->parts.p;
    ta.active = dActive;
    ta.hsoft = (const double *)c->hsoft.p;
    ta.tasks = singleTask ? (const Task *)c->dbgtask.p : (const Task *)c->tasks.p;
    ta.nTasks = nTasks;
    ta.bucketNode = singleTask ? (const int *)((const char *)c->dbgtask.p + sizeof(Task)) : (const int *)c->bnode.p;
    ta.nBuckets = nBuckets;
    ta.groupHead = (int *)c->ghead.p;
    ta.groupCnt = (int *)c->gcnt.p;
    ta.bucketCnt = (int *)c->bcnt.p;
    ta.bucketTot = (long long *)c->btot.p;
    ta.bucketOff = (const long long *)c->boff64.p;
    ta.taskCounter = (int *)c->misc.p;
    ta.errFlag = (int *)c->misc.p + 1;
    ta.poolCursor = (int *)c->misc.p + 3;
    ta.rootNode = rootNode;
    ta.nImages = im.n;
    ta.homeImage = im.home;
    ta.imgBits = im.bits;
    ta.imgOff = (const double *)c->imgoff.p;
    ta.iOrder = prm->iOrder;
    ta.maxBucket = c->maxBucket;
    ta.walkOnly = walkOnly ? 1 : 0;
    ta.mono64 = prm->bPeriodic ? 1 : 0;
    ta.bigMass = DBL_MAX;
    if (prm->bPeriodic && !walkOnly) { // the mass of the cell every image's walk starts from = everything there is
        double rootMass = 0.0;
        if (doEwald && c->haveRoot) rootMass = c->root[0]; // pkdCalcRoot / pkdDistribRoot: m of the whole box (pkd.c:4395-4493)
        else { // periodic without Ewald: read the root cell (one small synchronous copy)
            CK(cudaMemcpyAsync(&rootMass, &((const NodeW *)c->nodes.p)[rootNode].fMass, sizeof(double), cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
        }
        ta.bigMass = GG_BIG_FRAC * rootMass;
    }
    ta.sunNode = c->sunMode ? sunN : -1;
    ta.sunBox = 1e-14; // dTinyBox, pkd.c:3004
    ta.acc = (double *)c->acc.p;
    ta.pot = (double *)c->pot.p;
    ta.dtg = (double *)c->dtg.p;
    ta.counts = (int *)c->counts.p;
    ta.hacc = c->zc[0]; ta.hpot = c->zc[1]; ta.hdtg = c->zc[2];
    CK(cudaEventRecord(c->ev[1], c->st));
    if (nTasks > 0) {
        if (!walkOnly) {
            if (c->capBlocks > 0x7fffffffu / 32u)
                return gg_fail(GG_ERR_NOMEM, "gg_gravity: interaction lists need %zu blocks (> 2^31 references)", c->capBlocks);
            if ((rc = gg_ensure(c, c->pool, c->capBlocks * 32 * sizeof(unsigned)))) return rc;
            if ((rc = gg_ensure(c, c->nextblk, c->capBlocks * sizeof(int)))) return rc;
            if ((rc = gg_ensure(c, c->poolmask, c->capBlocks * 32 * sizeof(gg_mask_t)))) return rc;
        }
        ta.pool = (unsigned *)c->pool.p;
        ta.nextBlk = (int *)c->nextblk.p;
        ta.poolMask = (gg_mask_t *)c->poolmask.p;
        ta.capBlocks = (int)c->capBlocks;
        CK(gg_launch_walk_kernel(ta, c->nSM, c->st));
        ++c->nLaunches;
    }
    CK(cudaEventRecord(c->ev[5], c->st));
    // ---- bookkeeping (pkd.c:2945-2998): needs only the walk's counts (and Ewald's loop counts), so it runs on its own
    //      stream beside the list scatter and evaluation; fWeight goes straight to the caller's mapped array if any
    CK(cudaEventRecord(c->evWalk, c->st));
    CK(cudaStreamWaitEvent(c->st3, c->evWalk, 0));
    StatsKernelArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.nodes = (const NodeW *)c->nodes.p;
    sa.tasks = ta.tasks;
    sa.nTasks = nTasks;
    sa.active = dActive;
    sa.counts = (const int *)c->counts.p;
    sa.nLoop = (doEwald && !singleTask) ? (const int *)c->nloop.p : nullptr;
    sa.nEwh = nEwh;
    sa.iOrder = prm->iOrder;
    sa.iEwOrder = prm->iEwOrder;
    sa.fWeight = (double *)c->fweight.p;
    sa.hfWeight = (!walkOnly && dActive) ? c->zc[3] : nullptr; // partially active: only active entries may be written
    sa.sums = (unsigned long long *)c->sums.p;
    // every exit below must leave st3 idle: k_stats reads counts / writes fweight, sums and the caller's mapped fWeight
    struct St3Guard {
        cudaStream_t s;
        bool armed;
        ~St3Guard() { if (armed) cudaStreamSynchronize(s); }
    } st3Guard{c->st3, true};
    CK(gg_launch_stats_kernel(sa, c->st3));
    if (!walkOnly && c->zc[3] && !dActive && n > 0) // one coalesced copy (per-bucket 8-byte stores over PCIe cost ~1 ms per 1 M particles)
        CK(cudaMemcpyAsync(c->zcHost[3], c->fweight.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st3));
    CK(cudaEventRecord(c->evStats, c->st3));
    if (nTasks > 0) ++c->nLaunches;
    bool evalTimed = false, evalQueued = false, fastLists = false;
    long long nListEntries = 0, capListEntries = 0;
    int nCh = 1, hBounds[2 * 66];
    ta.taskBegin = 0;
    ta.taskEnd = nTasks;
    if (nTasks > 0 && !walkOnly) {
        // per-bucket list offsets = exclusive scan of the entry counts the walk left; the total sizes the list array
        size_t tmpBytes = 0;
        CK(cudaMemsetAsync((long long *)c->btot.p + nBuckets, 0, sizeof(long long), c->st));
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, (long long *)c->btot.p, (long long *)c->boff64.p, nBuckets + 1, c->st));
        if ((rc = gg_ensure(c, c->cubtmp, tmpBytes))) return rc;
        CK(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmpBytes, (long long *)c->btot.p, (long long *)c->boff64.p, nBuckets + 1, c->st));
        // Once the list array has a size from an earlier evaluation, scatter and evaluation are queued straight behind the
        // walk: a one-thread kernel checks on the device that the walk's output fits (k_guard), the host looks at the same
        // numbers after the evaluation and re-runs with larger buffers if not.  The first evaluation of a context (no size
        // known) and every re-run read the sizes back between the walk and the scatter (GG_SYNC_WALK=1 forces that).
        static const bool syncWalk = getenv("GG_SYNC_WALK") != nullptr && atoi(getenv("GG_SYNC_WALK")) != 0;
        // gg_gravity_chunked: the evaluation in nCh launches; the callbacks may only see final results, so this order reads
        // the walk's sizes back first (a failed k_guard would leave a chunk unwritten)
        if (c->chunkFn && c->nChunk > 1 && c->zc[0] && !singleTask && !c->sunMode && nTasks >= 64 * c->nChunk && c->nChunk <= 64) {
            nCh = c->nChunk;
            if ((rc = gg_ensure(c, c->chunkb, 2 * 66 * sizeof(int)))) return rc;
            k_chunk_bounds<<<1, 96, 0, c->st>>>(ta.tasks, nTasks, (const NodeW *)c->nodes.p, nCh, n, (int *)c->chunkb.p);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(hBounds, c->chunkb.p, sizeof(int) * 2 * (nCh + 1), cudaMemcpyDeviceToHost, c->st));
        }
        fastLists = depth == 0 && !syncWalk && nCh == 1 && c->lists.cap >= 64 * sizeof(unsigned);
        if (fastLists) {
            capListEntries = (long long)(c->lists.cap / sizeof(unsigned)) - 32;
            ta.okFlag = (const int *)c->misc.p + 4;
            CK(gg_launch_guard_kernel(ta, capListEntries, (int *)c->misc.p + 4, c->st));
        } else {
            long long nEntries = 0;
            int hm3[4];
            CK(cudaMemcpyAsync(&nEntries, (long long *)c->boff64.p + nBuckets, sizeof(long long), cudaMemcpyDeviceToHost, c->st));
            CK(cudaMemcpyAsync(hm3, c->misc.p, sizeof(hm3), cudaMemcpyDeviceToHost, c->st));
            CK(cudaStreamSynchronize(c->st));
            tr.mark("gravity: walk sync");
            if (hm3[1]) return gg_fail(GG_ERR_UNSUPPORTED, "gg_gravity: walk frontier overflowed %d entries (tree too deep)", GG_STACK_CAP);
            if ((size_t)hm3[3] > c->capBlocks) {
                // the chain pool was too small: the walk kept counting, so hm3[3] is what it needs -- grow and run again
                if (depth >= 2) return gg_fail(GG_ERR_NOMEM, "gg_gravity: list pool overflow persists (%d blocks)", hm3[3]);
                c->capBlocks = (size_t)hm3[3] + (size_t)hm3[3] / 8 + 1024;
                return run_gravity(c, prm, singleTask, stats, depth + 1);
            }
            nListEntries = nEntries;
            if ((rc = gg_ensure(c, c->lists, ((size_t)nEntries + 32) * sizeof(unsigned)))) return rc;
        }
        ta.lists = (unsigned *)c->lists.p;
        CK(gg_launch_scatter_kernel(ta, c->nSM, c->st));
        if (c->momPending) CK(cudaStreamWaitEvent(c->st, c->evMom, 0)); // the moments arrive on the second stream
        CK(cudaEventRecord(c->ev[6], c->st));
        if (nCh == 1) CK(gg_launch_eval_kernel(ta, c->nSM, c->st));
        else {
            while ((int)c->evChunk.size() < nCh) {
                cudaEvent_t ev;
                CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync));
                c->evChunk.push_back(ev);
            }
            for (int k = 0; k < nCh; ++k) {
                if (k > 0) CK(cudaMemsetAsync(ta.taskCounter + 2, 0, sizeof(int), c->st)); // (every launch counts from its own begin)
                ta.taskBegin = hBounds[2 * k];
                ta.taskEnd = hBounds[2 * k + 2];
                if (ta.taskEnd > ta.taskBegin) {
                    CK(gg_launch_eval_kernel(ta, c->nSM, c->st));
                    ++c->nLaunches;
                }
                CK(cudaEventRecord(c->evChunk[k], c->st));
            }
            ta.taskBegin = 0;
            ta.taskEnd = nTasks;
        }
        evalQueued = true;
        evalTimed = true;
        c->nLaunches += 4;
    }
    CK(cudaEventRecord(c->ev[2], c->st));

    CK(cudaStreamWaitEvent(c->st, c->evStats, 0));
    CK(cudaEventRecord(c->ev[4], c->st));
    if (nCh > 1) { // hand the finished ranges to the caller while the later ones are still being evaluated
        CK(cudaEventSynchronize(c->evStats)); // (fWeight is delivered by k_stats's stream)
        for (int k = 0; k < nCh; ++k) {
            CK(cudaEventSynchronize(c->evChunk[k]));
            const int p0 = hBounds[2 * k + 1], p1 = k + 1 < nCh ? hBounds[2 * k + 3] : n;
            if (p1 > p0) c->chunkFn(c->chunkUser, p0, p1 - p0);
        }
        c->chunkFn = nullptr; // delivered: gg_gravity_chunked does not call again
    }

    unsigned long long hs[16];
    int hm[16];
    long long nEntriesFast = 0;
    CK(cudaMemcpyAsync(hs, c->sums.p, sizeof(hs), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(hm, c->misc.p, sizeof(hm), cudaMemcpyDeviceToHost, c->st));
    if (fastLists) CK(cudaMemcpyAsync(&nEntriesFast, (long long *)c->boff64.p + nBuckets, sizeof(long long), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    st3Guard.armed = false; // c->st waited for evStats
    tr.mark("gravity: final sync");
    if (evalQueued) c->momPending = false; // k_eval waited for the moments and has finished
    if (hm[1]) return gg_fail(GG_ERR_UNSUPPORTED, "gg_gravity: walk frontier overflowed %d entries (tree too deep)", GG_STACK_CAP);
    if (!walkOnly && (size_t)hm[3] > c->capBlocks) {
        // the list pool was too small: the walk kept counting, so hm[3] is what it needs -- grow and run again
        if (depth >= 2) return gg_fail(GG_ERR_NOMEM, "gg_gravity: list pool overflow persists (%d blocks)", hm[3]);
        c->capBlocks = (size_t)hm[3] + (size_t)hm[3] / 8 + 1024;
        return run_gravity(c, prm, singleTask, stats, depth + 1);
    }
    if (fastLists) {
        nListEntries = nEntriesFast;
        if (nEntriesFast > capListEntries) { // k_guard held scatter and evaluation back: the list array was too small
            if ((rc = gg_ensure(c, c->lists, ((size_t)nEntriesFast + (size_t)nEntriesFast / 8 + 32) * sizeof(unsigned)))) return rc;
            return run_gravity(c, prm, singleTask, stats, depth + 1);
        }
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->nActive = (int)hs[0];
        stats->dPartSum = (double)hs[1];
        stats->dCellSum = (double)hs[2];
        stats->dSoftSum = (double)hs[3];
        stats->dFlop = (double)hs[4] + (double)hs[5];
        stats->dFlopEwald = (double)hs[5];
        stats->nMaxPart = (int)hs[6];
        stats->nMaxCellSoft = (int)hs[7];
        stats->nMaxCellNewt = (int)hs[8];
        float ms;
        CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); stats->msTree = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[5])); stats->msWalk = ms;
        if (evalTimed) { CK(cudaEventElapsedTime(&ms, c->ev[6], c->ev[2])); stats->msEval = ms; }
        stats->nListEntries = (double)nListEntries;
        if (usePre) CK(cudaEventElapsedTime(&ms, c->evEw[0], c->evEw[1])); // (on the side stream, beside the upload)
        else CK(cudaEventElapsedTime(&ms, c->ev[7], c->ev[3]));
        stats->msEwald = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[4])); stats->msTotal = ms;
        stats->nKernelLaunches = c->nLaunches;
    }
    return GG_OK;
}

// The bDoSun pass of pkdGravAll (pkd.c:3003-3041): a dummy ACTIVE sink at the origin with softening dSunSoft, in a bucket
// of its own whose cell box is +-1e-14, walks the tree (pkdBucketWalk) and is evaluated (pkdBucketInteract); its
// acceleration is the indirect term of solar-system runs.  The dummy lives in the spare slot behind the particles and
// the spare node behind the tree; the particles' results and counters are left as the main pass produced them, and --
// like the reference -- the dummy's interactions are not counted in dPartSum / dCellSum / dFlop.
int run_sun(gg_context *c, const gg_params *prm, gg_stats *stats) {
    if (prm->bPeriodic || prm->nReps != 0)
        return gg_fail(GG_ERR_ARG, "gg_gravity: bDoSun needs open boundaries (the reference asserts it, pkd.c:3013-3014)");
    int rc;
    if ((rc = gg_finish_mom(c))) return rc;
    // Several domains (pst.c:3258-3277 hands bDoSun to the ONE rank whose domain holds the origin; that rank's dummy sink
    // then walks top tree, local and remote trees like any bucket): the spare slots are the ones behind everything loaded.
    const int sunP = c->nPartAll, sunN = c->nNodesAll + c->nTop;
    if ((rc = gg_ensure(c, c->parts, ((size_t)sunP + 1) * sizeof(PartS), (size_t)sunP * sizeof(PartS)))) return rc;
    if ((rc = gg_ensure(c, c->nodes, ((size_t)sunN + 1) * sizeof(NodeW), (size_t)sunN * sizeof(NodeW)))) return rc;
    if ((rc = gg_ensure(c, c->momf, ((size_t)sunN + 1) * 128, (size_t)sunN * 128))) return rc;
    if ((rc = gg_ensure(c, c->momq, ((size_t)sunN + 1) * 48, (size_t)sunN * 48))) return rc;
    const int nLoc = c->dom[0].nPart;
    if ((rc = gg_ensure(c, c->hsoft, ((size_t)sunP + 1) * sizeof(double), ((size_t)nLoc + 1) * sizeof(double)))) return rc;
    if (!c->hActive.empty() && (rc = gg_ensure(c, c->active, ((size_t)sunP + 1) * sizeof(int), ((size_t)nLoc + 1) * sizeof(int)))) return rc;
    PartS ps;
    ps.x = ps.y = ps.z = 0.0; ps.m = 0.f; ps.h = (float)prm->dSunSoft;
    NodeW w;
    w.rx = w.ry = w.rz = 0.0; w.fMass = 0.0; w.fOpen2 = 0.0; w.fSoft = prm->dSunSoft;
    w.c0 = w.c1 = -1; w.pLower = sunP; w.nP = 1;
    const double hs = prm->dSunSoft;
    const int one = 1;
    CK(cudaMemcpyAsync((PartS *)c->parts.p + sunP, &ps, sizeof(ps), cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync((NodeW *)c->nodes.p + sunN, &w, sizeof(w), cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync((double *)c->hsoft.p + sunP, &hs, sizeof(hs), cudaMemcpyHostToDevice, c->st));
    if (!c->hActive.empty()) CK(cudaMemcpyAsync((int *)c->active.p + sunP, &one, sizeof(one), cudaMemcpyHostToDevice, c->st));
    CK(cudaStreamSynchronize(c->st)); // (the sources are stack variables)
    Task t;
    t.node = sunN; t.pass = 0; t.ord = 0; t.pad = 0;
    gg_params p2 = *prm;
    p2.flags = GG_FLAG_NO_DOWNLOAD;
    p2.bDoSun = 0;
    gg_stats st2;
    c->sunMode = true;
    rc = run_gravity(c, &p2, &t, &st2);
    c->sunMode = false;
    if (rc) return rc;
    double a3[3];
    int c3[3];
    CK(cudaMemcpyAsync(a3, (const double *)c->acc.p + 3 * (size_t)sunP, sizeof(a3), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(c3, (const int *)c->counts.p + 3 * (size_t)sunN, sizeof(c3), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (stats) {
        stats->aSun[0] = a3[0]; stats->aSun[1] = a3[1]; stats->aSun[2] = a3[2];
        stats->nSunPart = c3[0]; stats->nSunCellSoft = c3[1]; stats->nSunCellNewt = c3[2];
        stats->nKernelLaunches = c->nLaunches;
    }
    return GG_OK;
}

} // namespace

extern "C" {

int gg_gravity_chunked(gg_context *c, const gg_params *prm, double *a, double *fPot, double *dtGrav, double *fWeight,
                       gg_stats *stats, int nChunks, gg_chunk_fn onChunk, void *user) {
    if (!c || !prm) return gg_fail(GG_ERR_ARG, "gg_gravity_chunked: null argument");
    c->nChunk = nChunks > 64 ? 64 : nChunks;
    c->chunkFn = onChunk;
    c->chunkUser = user;
    const int rc = gg_gravity(c, prm, a, fPot, dtGrav, fWeight, stats);
    const bool pending = c->chunkFn != nullptr; // not delivered in pieces: one call for everything
    c->chunkFn = nullptr;
    c->nChunk = 1;
    if (rc == GG_OK && pending && onChunk && !c->dom.empty() && c->dom[0].nPart > 0 &&
        !(prm->flags & (GG_FLAG_NO_DOWNLOAD | GG_FLAG_WALK_ONLY)))
        onChunk(user, 0, c->dom[0].nPart);
    return rc;
}

int gg_gravity(gg_context *c, const gg_params *prm, double *a, double *fPot, double *dtGrav, double *fWeight,
               gg_stats *stats) {
    if (!c || !prm) return gg_fail(GG_ERR_ARG, "gg_gravity: null argument");
    const bool wantOut = !(prm->flags & (GG_FLAG_NO_DOWNLOAD | GG_FLAG_WALK_ONLY));
    if (wantOut && (!a || !fPot || !dtGrav || !fWeight)) return gg_fail(GG_ERR_ARG, "gg_gravity: null output array");
    // Zero-copy delivery: in overwrite mode, output arrays that are mapped pinned host memory (gg_host_alloc,
    // cudaHostAlloc, cudaHostRegister) are written by the kernels themselves as each sink bucket finishes, so the
    // device->host transfer overlaps the evaluation instead of following it.  Only ACTIVE particles are written
    // (the reference's contract, pkd.c:2851-2861).  Anything else takes the staged copy below.
    bool zeroCopy = false;
    c->zc[0] = c->zc[1] = c->zc[2] = c->zc[3] = nullptr;
    if (wantOut && !prm->accumulate && !c->dom.empty() && c->dom[0].nPart > 0) {
        double *hp[4] = {a, fPot, dtGrav, fWeight}, *dp[4] = {nullptr, nullptr, nullptr, nullptr};
        zeroCopy = true;
        CK(cudaSetDevice(c->device));
        for (int k = 0; k < 4 && zeroCopy; ++k) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, hp[k]) != cudaSuccess) { cudaGetLastError(); zeroCopy = false; break; }
            if (at.type != cudaMemoryTypeHost || !at.devicePointer) zeroCopy = false;
            else dp[k] = (double *)at.devicePointer;
        }
        if (zeroCopy) for (int k = 0; k < 4; ++k) { c->zc[k] = dp[k]; c->zcHost[k] = hp[k]; }
    }
    if (c->stateN > 0 && c->stateDirty)
        return gg_fail(GG_ERR_ARG, "gg_gravity: the resident particles moved since the last tree build (gg_state_build first)");
    int rc = run_gravity(c, prm, nullptr, stats);
    c->zc[0] = c->zc[1] = c->zc[2] = c->zc[3] = nullptr;
    if (rc) return rc;
    if (c->stateN > 0 && !(prm->flags & GG_FLAG_WALK_ONLY)) c->stateForces = true;
    if (prm->bDoSun && !(prm->flags & GG_FLAG_WALK_ONLY)) {
        if ((rc = run_sun(c, prm, stats))) return rc;
    }
    if (!wantOut || zeroCopy) return GG_OK;
    const int n = c->dom[0].nPart;
    if (n == 0) return GG_OK;
    const bool all = c->hActive.empty();
    if (!prm->accumulate && all) { // overwrite, every particle active: straight into the caller's arrays
        CK(cudaMemcpyAsync(a, c->acc.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(fPot, c->pot.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(dtGrav, c->dtg.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(fWeight, c->fweight.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        return GG_OK;
    }
    // staged: pinned bounce buffer, then += / max / = (accumulate, the reference's in-place semantics: pkd.c:2851-2861,
    // grav.c:100,192-195) or plain assignment (overwrite) on ACTIVE particles only -- inactive ones are never touched
    if ((rc = ensure_pinned(c, sizeof(double) * 6 * (size_t)n))) return rc;
    double *h = (double *)c->pinned, *ha = h, *hp = h + 3 * (size_t)n, *hd = hp + n, *hw = hd + n;
    CK(cudaMemcpyAsync(ha, c->acc.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(hp, c->pot.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(hd, c->dtg.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(hw, c->fweight.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const int *act = all ? nullptr : c->hActive.data();
    const bool acc = prm->accumulate != 0;
    auto merge = [=](int i0, int i1) {
        for (int i = i0; i < i1; ++i) {
            if (act && !act[i]) continue;
            if (acc) {
                a[3 * (size_t)i] += ha[3 * (size_t)i];
                a[3 * (size_t)i + 1] += ha[3 * (size_t)i + 1];
                a[3 * (size_t)i + 2] += ha[3 * (size_t)i + 2];
                fPot[i] += hp[i];
                if (hd[i] > dtGrav[i]) dtGrav[i] = hd[i];
            } else {
                a[3 * (size_t)i] = ha[3 * (size_t)i];
                a[3 * (size_t)i + 1] = ha[3 * (size_t)i + 1];
                a[3 * (size_t)i + 2] = ha[3 * (size_t)i + 2];
                fPot[i] = hp[i];
                dtGrav[i] = hd[i];
            }
            fWeight[i] = hw[i];
        }
    };
    const int nT = n >= (1 << 17) ? 4 : 1; // the merge is memory-bound: a few threads saturate it
    if (nT == 1) merge(0, n);
    else {
        std::vector<std::thread> th;
        const int chunk = (n + nT - 1) / nT;
        for (int t = 1; t < nT; ++t) th.emplace_back(merge, t * chunk, std::min(n, (t + 1) * chunk));
        merge(0, std::min(n, chunk));
        for (auto &t : th) t.join();
    }
    return GG_OK;
}

int gg_bucket_counts(gg_context *c, int *counts3) {
    if (!c || !counts3 || c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_bucket_counts: bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(counts3, c->counts.p, sizeof(int) * 3 * (size_t)c->dom[0].nNodes, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_bucket_walk(gg_context *c, const gg_params *prm, int iBucket, int n3[3]) {
    if (!c || !prm || !n3 || c->dom.empty()) return gg_fail(GG_ERR_ARG, "gg_bucket_walk: bad argument");
    if (iBucket < 0 || iBucket >= c->dom[0].nNodes) return gg_fail(GG_ERR_ARG, "gg_bucket_walk: iBucket=%d", iBucket);
    gg_params p = *prm;
    p.flags |= GG_FLAG_WALK_ONLY;
    Task t{iBucket, 0, 0, 0};
    c->stateForces = false; // run_gravity clears the device result arrays: a following kick must not consume them
    int rc = run_gravity(c, &p, &t, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(n3, (int *)c->counts.p + 3 * (size_t)iBucket, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

// The two inner seams of the reference's bucket loop as entry points (parity / debugging hooks).
static int bucket_range(gg_context *c, const char *who, int iBucket, int *pLo, int *pN) {
    if (c->dom.empty()) return gg_fail(GG_ERR_ARG, "%s: no local domain", who);
    if (iBucket < 0 || iBucket >= c->dom[0].nNodes) return gg_fail(GG_ERR_ARG, "%s: iBucket=%d", who, iBucket);
    NodeW w;
    CK(cudaMemcpyAsync(&w, (const NodeW *)c->nodes.p + iBucket, sizeof(w), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (w.c0 >= 0) return gg_fail(GG_ERR_ARG, "%s: node %d is a cell, not a bucket", who, iBucket);
    *pLo = w.pLower; *pN = w.nP;
    return GG_OK;
}

int gg_bucket_interact(gg_context *c, const gg_params *prm, int iBucket, int nMax, double *a, double *fPot, double *dtGrav,
                       int n3[3]) {
    if (!c || !prm || !a || !fPot || !dtGrav) return gg_fail(GG_ERR_ARG, "gg_bucket_interact: null argument");
    CK(cudaSetDevice(c->device));
    int lo = 0, np = 0, rc;
    if ((rc = bucket_range(c, "gg_bucket_interact", iBucket, &lo, &np))) return rc;
    if (np > nMax) return gg_fail(GG_ERR_ARG, "gg_bucket_interact: the bucket holds %d particles, room for %d", np, nMax);
    gg_params p = *prm;
    p.flags = GG_FLAG_NO_DOWNLOAD; p.bDoSun = 0;
    Task t{iBucket, 0, 0, 0};
    c->stateForces = false; // run_gravity clears the device result arrays
    const int nPass = (np + GG_MAX_SINKS - 1) / GG_MAX_SINKS;
    for (int pass = 0; pass < nPass; ++pass) { // a pass evaluates <= 8 active sinks of the bucket; results accumulate
        t.pass = pass;
        const bool keep = c->sunMode;
        if (pass > 0) c->sunMode = true; // (re-uses the "leave the other particles' results alone" mode of the bDoSun pass)
        rc = run_gravity(c, &p, &t, nullptr);
        c->sunMode = keep;
        if (rc) return rc;
    }
    CK(cudaMemcpyAsync(a, (const double *)c->acc.p + 3 * (size_t)lo, sizeof(double) * 3 * np, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(fPot, (const double *)c->pot.p + lo, sizeof(double) * np, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(dtGrav, (const double *)c->dtg.p + lo, sizeof(double) * np, cudaMemcpyDeviceToHost, c->st));
    if (n3) CK(cudaMemcpyAsync(n3, (const int *)c->counts.p + 3 * (size_t)iBucket, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_bucket_ewald(gg_context *c, const gg_params *prm, int iBucket, int nMax, double *a, double *fPot, int *pnFlop) {
    if (!c || !prm || !a || !fPot) return gg_fail(GG_ERR_ARG, "gg_bucket_ewald: null argument");
    CK(cudaSetDevice(c->device));
    int lo = 0, np = 0, rc;
    if ((rc = bucket_range(c, "gg_bucket_ewald", iBucket, &lo, &np))) return rc;
    if (np > nMax) return gg_fail(GG_ERR_ARG, "gg_bucket_ewald: the bucket holds %d particles, room for %d", np, nMax);
    if (c->rootLazy && (rc = fetch_root_lazy(c))) return rc;
    if (!c->haveRoot) return gg_fail(GG_ERR_ARG, "gg_bucket_ewald: gg_set_root_moments has not been called");
    const int n = c->dom[0].nPart;
    if ((rc = gg_ensure(c, c->acc, (size_t)(n + 1) * 3 * sizeof(double))) || (rc = gg_ensure(c, c->pot, (size_t)(n + 1) * sizeof(double))) ||
        (rc = gg_ensure(c, c->nloop, (size_t)(n + 1) * sizeof(int))))
        return rc;
    EwaldKernelArgs ea;
    if ((rc = make_ewald_args(c, prm, ea, c->st))) return rc;
    c->stateForces = false;
    // the kernel on the bucket's slice of the particle arrays (the correction is per particle: no neighbours involved)
    ea.parts = (const PartS *)c->parts.p + lo;
    ea.active = c->hActive.empty() ? nullptr : (const int *)c->active.p + lo;
    ea.n = np;
    ea.acc = (double *)c->acc.p + 3 * (size_t)lo;
    ea.pot = (double *)c->pot.p + lo;
    ea.nLoop = (int *)c->nloop.p + lo;
    CK(cudaMemsetAsync(ea.acc, 0, sizeof(double) * 3 * np, c->st));
    CK(cudaMemsetAsync(ea.pot, 0, sizeof(double) * np, c->st));
    CK(cudaMemsetAsync(ea.nLoop, 0, sizeof(int) * np, c->st));
    CK(gg_launch_ewald_kernel(ea, c->st));
    std::vector<int> nl((size_t)np);
    CK(cudaMemcpyAsync(a, ea.acc, sizeof(double) * 3 * np, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(fPot, ea.pot, sizeof(double) * np, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(nl.data(), ea.nLoop, sizeof(int) * np, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (pnFlop) { // ewald.c:175-176: nLoop real-space terms and nEwhLoop k-space rows per active particle
        static const int mflop[5] = {10, 10, 48, 151, 343};
        long long f = 0;
        for (int j = 0; j < np; ++j)
            if (c->hActive.empty() || c->hActive[(size_t)lo + j]) f += (long long)nl[j] * (104 + mflop[prm->iEwOrder]) + (long long)ea.nEwh * 58;
        *pnFlop = (int)f;
    }
    return GG_OK;
}

int gg_ewald_table(gg_context *c, const gg_params *prm, double *ewt5, int nMax, int *pnEwh) {
    if (!c || !prm || !pnEwh) return gg_fail(GG_ERR_ARG, "gg_ewald_table: bad argument");
    if (c->rootLazy) {
        int rcr = fetch_root_lazy(c);
        if (rcr) return rcr;
    }
    if (!c->haveRoot) return gg_fail(GG_ERR_ARG, "gg_ewald_table: gg_set_root_moments has not been called");
    std::vector<double> ewt;
    gg_ewald_table_host(c->root, prm->fPeriod[0], prm->fEwhCut, prm->iEwOrder, ewt);
    *pnEwh = (int)(ewt.size() / 5);
    if (ewt5) memcpy(ewt5, ewt.data(), sizeof(double) * 5 * (size_t)(*pnEwh < nMax ? *pnEwh : nMax));
    return GG_OK;
}

int gg_measure_fp32_peak(gg_context *c, double *pTflops, double *pMs) {
    if (!c || !pTflops) return gg_fail(GG_ERR_ARG, "gg_measure_fp32_peak: null");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = gg_ensure(c, c->misc, 16 * sizeof(int)))) return rc;
    const int iters = 4096, blocks = c->nSM * 8, threads = 256;
    double best = 0.0, bestMs = 0.0;
    for (int rep = 0; rep < 4; ++rep) { // first repetition warms the clocks up
        CK(cudaEventRecord(c->ev[0], c->st));
        k_fma_peak<<<blocks, threads, 0, c->st>>>(iters, (float *)c->misc.p);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->ev[1], c->st));
        CK(cudaStreamSynchronize(c->st));
        float ms;
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
        const double tf = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) { best = tf; bestMs = ms; }
    }
    *pTflops = best;
    if (pMs) *pMs = bestMs;
    return GG_OK;
}

int gg_timer_start(gg_context *c) {
    if (!c) return gg_fail(GG_ERR_ARG, "gg_timer_start: null");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->evt[0], c->st));
    return GG_OK;
}

int gg_timer_stop(gg_context *c, double *pMs) {
    if (!c || !pMs) return gg_fail(GG_ERR_ARG, "gg_timer_stop: null");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->evt[1], c->st));
    CK(cudaEventSynchronize(c->evt[1]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->evt[0], c->evt[1]));
    *pMs = ms;
    return GG_OK;
}

int gg_flush_l2(gg_context *c) {
    if (!c) return gg_fail(GG_ERR_ARG, "gg_flush_l2: null");
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)384 << 20; // 3x the 126 MB L2
    int rc;
    if ((rc = gg_ensure(c, c->flush, bytes))) return rc;
    CK(cudaMemsetAsync(c->flush.p, 0x5a, bytes, c->st));
    CK(cudaStreamSynchronize(c->st));
    return GG_OK;
}

int gg_device_results(gg_context *c, void **a, void **fPot, void **dtGrav, void **fWeight) {
    if (!c) return gg_fail(GG_ERR_ARG, "gg_device_results: null");
    if (a) *a = c->acc.p;
    if (fPot) *fPot = c->pot.p;
    if (dtGrav) *dtGrav = c->dtg.p;
    if (fWeight) *fWeight = c->fweight.p;
    return GG_OK;
}

} // extern "C"
