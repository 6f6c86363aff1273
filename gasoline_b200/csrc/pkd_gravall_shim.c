/*
 * pkd_gravall_shim.c -- the drop-in: a replacement for Gasoline's pkdGravAll (pkd.c:2868, prototype pkd.h:797-801)
 * that runs the force evaluation on a B200 through the C ABI of include/gasoline_b200.h.
 *
 * How a Gasoline host adopts it (INTEGRATION.md has the link lines): compile the host's pkd.c with
 * -DpkdGravAll=pkdGravAll_cpu, compile THIS file against the host's own pkd.h with the host's own -D flags (the
 * PARTICLE and KDN layouts change with them, so every field is copied by name -- never by offset), and link
 * libgasoline_b200.so.  pstGravity (pst.c:3310-3315) then calls this function unchanged.
 *
 * What it does, per call (= per rank per force evaluation):
 *   1. flattens pkd->kdNodes[0..nNodes) and pkd->pStore[0..nLocal) into the SoA views gg_set_local takes (pinned
 *      staging buffers kept between calls, grown by high-water mark; the reference frees and rebuilds kdNodes before
 *      every gravity call, pkd.c:2636-2642, so nothing can be assumed resident).  The AoS records are 536 B / 184 B
 *      wide, so this pass is memory-bound host work: it runs on GG_SHIM_THREADS threads per rank (default: the cores
 *      divided among the ranks of the process, at most 16).
 *      The cells' multipole moments are NOT copied unless GG_SHIM_HOST_MOMENTS=1: the device forms them from the
 *      particles (gg_tree.mom = NULL; same definition, FP64, forces identical to rounding of the FP32 records);
 *   2. hands over pkd->ilcnRoot (pkdDistribRoot, pkd.c:4472) when Ewald is on;
 *   3. gg_gravity in overwrite mode into pinned result arrays (the kernels deliver them zero-copy while they run), then
 *      ONE threaded pass applies the reference's in-place semantics to pStore: a, fPot += ; dtGrav = max ; fWeight =
 *      for ACTIVE particles only; nActive / dPartSum / dCellSum / dSoftSum / dFlop go back through the
 *      pointer arguments exactly as pkdGravAll returns them (pkd.c:2945-2949, grav.c:246-247, ewald.c:175-176).
 * Errors follow the host's convention: print and abort (the reference asserts; there is no CPU fallback here).
 *
 * pkdBuildBinary (pkd.c:2627) is substituted the same way (-DpkdBuildBinary=pkdBuildBinary_cpu on the host's pkd.c): with
 * GG_SHIM_DEVICE_TREE=1 in the environment the gravity tree is built on the GPU (gg_build_local: the same cells, the
 * same numbering, the same order of pStore, r / fMass / fSoft / fOpen2 / bnd bit for bit), pStore is permuted and
 * kdNodes filled from it, so that pstBuildTree and everything downstream see the tree they would have built -- in
 * ~30 ms instead of ~1.2 s for 1 M particles.  Without the variable, or for what the device build does not cover
 * (iOpenType = OPEN_ABSPAR, a tree over part of the particles, bGravity = 0), the host's own build runs.
 *
 * Several MDL ranks (mdlThreads > 1; pst.c:3248-3303 fans pstGravity out, every leaf calls this function at the same
 * time): rank r drives GPU r (GG_SHIM_DEVICE overrides; r modulo the device count otherwise).  On its first call every
 * rank joins the library's communicator -- the 128-byte NCCL id (or, with GG_SHIM_COMM=local / fewer GPUs than ranks,
 * the address of an in-process group for thread ranks) travels from rank 0 through the host's own MDL: a read-only
 * cache on CID_PARTICLE, the slot pkdGravAll itself opens at this point (pkd.c:2896-2899), read with mdlAquire.  Per
 * call: pkd->kdTop (pkdDistribCells, pkd.c:4376; walked as a heap from ROOT like walk.c:342-435 walks it) goes down
 * through gg_set_top, the ranks' root bounds (the kdTop leaves) to gg_exchange, which replaces pkdRemoteWalk's pulls
 * (walk.c:181-304) by one push of pruned trees over NCCL / NVLink; then gg_gravity as with one rank.
 * bDoSun (pkd.c:3003-3041, the indirect term of solar-system runs) is passed through: gg_gravity evaluates the dummy
 * sink at the origin after the particles and aSun comes back through gg_stats.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "pkd.h"
#include "opentype.h"
#include "gasoline_b200.h"

typedef struct {
    gg_context *ctx;
    size_t capNodes, capPart;
    /* pinned staging, SoA */
    double *bnd, *r, *fMass, *fSoft, *fOpen2, *mom;
    int *pLower, *pUpper, *iLower, *iUpper;
    double *x, *y, *z, *m, *h, *a, *pot, *dt, *w;
    int *active;
    /* pkdBuildBinary on the device: construction by-products per cell, permutation per particle */
    double *fSplit, *fBmax;
    int *iDim, *order;
    int bJoined; /* this rank has joined the library's communicator (multi-rank runs) */
    PARTICLE *tmp; /* pStore permutation scratch, kept between builds (fresh pages cost more than the copy) */
    size_t capTmp;
    const KDN *builtNodes; /* pkd->kdNodes as our pkdBuildBinary left it (NULL: the device holds no tree of this host) */
    int builtN;
} SHIM;

static SHIM g_shim[64]; /* one per MDL rank living in this process (pthread-MDL ranks are threads) */

/* Debug / parity seam for hosts and tests: the library context rank idSelf of this process drives (NULL before its first
 * force evaluation) -- e.g. for gg_bucket_counts after a pstGravity. */
gg_context *pkdGravAllContext(int idSelf) { return (idSelf >= 0 && idSelf < 64) ? g_shim[idSelf].ctx : NULL; }

/* ---- a minimal parallel-for over [0, n): the flatten / write-back passes are bandwidth-bound strided copies ---- */
typedef struct {
    void (*fn)(void *, size_t, size_t);
    void *arg;
    size_t lo, hi;
} JOB;

static void *job_main(void *p) {
    JOB *j = (JOB *)p;
    j->fn(j->arg, j->lo, j->hi);
    return NULL;
}

static int g_shimRanks = 1; /* ranks of the host inside this process (pthread MDL): they share the cores */

static int shim_threads(void) {
    static int nAll = 0;
    int nT;
    if (!nAll) {
        const char *e = getenv("GG_SHIM_THREADS");
        long nc = sysconf(_SC_NPROCESSORS_ONLN);
        nAll = e ? -atoi(e) : (int)nc; /* negative: set by hand, per rank */
        if (!nAll) nAll = 1;
    }
    nT = nAll < 0 ? -nAll : nAll / (g_shimRanks > 0 ? g_shimRanks : 1);
    if (nAll > 0 && nT > 16) nT = 16;
    if (nT < 1) nT = 1;
    if (nT > 64) nT = 64;
    return nT;
}

static void parallel_for(size_t n, void (*fn)(void *, size_t, size_t), void *arg) {
    int nT = shim_threads(), t, started = 0;
    pthread_t th[64];
    JOB job[64];
    size_t chunk;
    if (n < 65536) nT = 1;
    chunk = (n + nT - 1) / nT;
    for (t = 1; t < nT; ++t) {
        job[t].fn = fn; job[t].arg = arg;
        job[t].lo = (size_t)t * chunk; job[t].hi = job[t].lo + chunk < n ? job[t].lo + chunk : n;
        if (job[t].lo >= job[t].hi) break;
        if (pthread_create(&th[t], NULL, job_main, &job[t]) != 0) { /* run it here instead */
            fn(arg, job[t].lo, job[t].hi);
            th[t] = 0;
        }
        started = t;
    }
    fn(arg, 0, chunk < n ? chunk : n);
    for (t = 1; t <= started; ++t)
        if (th[t]) pthread_join(th[t], NULL);
}

/* the same over [lo, hi) */
typedef struct {
    void (*fn)(void *, size_t, size_t);
    void *arg;
    size_t base;
} SHIFT;
static void shifted(void *arg, size_t lo, size_t hi) {
    SHIFT *h = (SHIFT *)arg;
    h->fn(h->arg, h->base + lo, h->base + hi);
}
static void parallel_range(size_t lo, size_t hi, void (*fn)(void *, size_t, size_t), void *arg) {
    SHIFT h;
    h.fn = fn; h.arg = arg; h.base = lo;
    if (hi > lo) parallel_for(hi - lo, shifted, &h);
}

typedef struct {
    PKD pkd;
    SHIM *s;
    int bMom;
    int iOrder;      /* multipole order the host asked pkdBuildBinary for */
    PARTICLE *tmp;   /* pStore permutation scratch */
} PASS;

static void flatten_nodes(void *arg, size_t lo, size_t hi) {
    PASS *a = (PASS *)arg;
    SHIM *s = a->s;
    size_t i;
    int j;
    for (i = lo; i < hi; ++i) {
        const KDN *c = &a->pkd->kdNodes[i];
        for (j = 0; j < 3; ++j) {
            s->bnd[6 * i + j] = c->bnd.fMin[j];
            s->bnd[6 * i + 3 + j] = c->bnd.fMax[j];
            s->r[3 * i + j] = c->r[j];
        }
        s->fMass[i] = c->fMass; s->fSoft[i] = c->fSoft; s->fOpen2[i] = c->fOpen2;
        s->pLower[i] = c->pLower; s->pUpper[i] = c->pUpper; s->iLower[i] = c->iLower; s->iUpper[i] = c->iUpper;
        if (a->bMom) {
            const struct pkdCalcCellStruct *q = &c->mom;
            double *mo = &s->mom[(size_t)GG_NMOM * i];
            mo[0] = q->Qxx; mo[1] = q->Qyy; mo[2] = q->Qzz; mo[3] = q->Qxy; mo[4] = q->Qxz; mo[5] = q->Qyz;
            mo[6] = q->Oxxx; mo[7] = q->Oxyy; mo[8] = q->Oxxy; mo[9] = q->Oyyy; mo[10] = q->Oxxz; mo[11] = q->Oyyz;
            mo[12] = q->Oxyz; mo[13] = q->Oxzz; mo[14] = q->Oyzz; mo[15] = q->Ozzz;
            mo[16] = q->Hxxxx; mo[17] = q->Hxyyy; mo[18] = q->Hxxxy; mo[19] = q->Hyyyy; mo[20] = q->Hxxxz;
            mo[21] = q->Hyyyz; mo[22] = q->Hxxyy; mo[23] = q->Hxxyz; mo[24] = q->Hxyyz; mo[25] = q->Hxxzz;
            mo[26] = q->Hxyzz; mo[27] = q->Hxzzz; mo[28] = q->Hyyzz; mo[29] = q->Hyzzz; mo[30] = q->Hzzzz;
        }
    }
}

static void flatten_particles(void *arg, size_t lo, size_t hi) {
    PASS *a = (PASS *)arg;
    SHIM *s = a->s;
    size_t i;
    for (i = lo; i < hi; ++i) {
        const PARTICLE *p = &a->pkd->pStore[i];
        s->x[i] = p->r[0]; s->y[i] = p->r[1]; s->z[i] = p->r[2];
        s->m[i] = p->fMass; s->h[i] = p->fSoft;
        s->active[i] = TYPEQueryACTIVE(p) ? 1 : 0;
    }
}

static void flatten_active(void *arg, size_t lo, size_t hi) {
    PASS *a = (PASS *)arg;
    size_t i;
    for (i = lo; i < hi; ++i) a->s->active[i] = TYPEQueryACTIVE(&a->pkd->pStore[i]) ? 1 : 0;
}

/* a, fPot += ; dtGrav = max ; fWeight = : what pkdBucketInteract / pkdBucketEwald / pkdBucketWeight leave in pStore
 * (grav.c:100,192-195, ewald.c:166-170, pkd.c:2851-2861), ACTIVE particles only */
static void write_back(void *arg, size_t lo, size_t hi) {
    PASS *a = (PASS *)arg;
    SHIM *s = a->s;
    size_t i;
    for (i = lo; i < hi; ++i) {
        PARTICLE *p = &a->pkd->pStore[i];
        if (!s->active[i]) continue;
        p->a[0] += s->a[3 * i]; p->a[1] += s->a[3 * i + 1]; p->a[2] += s->a[3 * i + 2];
        p->fPot += s->pot[i];
        if (s->dt[i] > p->dtGrav) p->dtGrav = s->dt[i];
        p->fWeight = s->w[i];
    }
}

static void write_back_chunk(void *arg, int first, int count) {
    parallel_range((size_t)first, (size_t)first + (size_t)count, write_back, arg);
}

/* GG_SHIM_TRACE=1: phase timings of the two entry points on stderr */
static double lap(double *t0, const char *what) {
    struct timespec ts;
    double t, d;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    t = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    d = t - *t0;
    if (what && getenv("GG_SHIM_TRACE")) fprintf(stderr, "[gg shim] %-28s %8.2f ms\n", what, d);
    *t0 = t;
    return d;
}

static void die(const char *what) {
    fprintf(stderr, "pkdGravAll (gasoline_b200): %s failed: %s\n", what, gg_last_error());
    abort();
}

static void *pinned(size_t bytes) {
    void *p = NULL;
    if (gg_host_alloc(&p, bytes) != GG_OK) die("gg_host_alloc");
    return p;
}

static void reserve(SHIM *s, size_t nNodes, size_t nPart) {
    if (nNodes > s->capNodes) {
        size_t c = nNodes + nNodes / 4 + 16;
        gg_host_free(s->bnd); gg_host_free(s->r); gg_host_free(s->fMass); gg_host_free(s->fSoft);
        gg_host_free(s->fOpen2); gg_host_free(s->mom); gg_host_free(s->pLower); gg_host_free(s->pUpper);
        gg_host_free(s->iLower); gg_host_free(s->iUpper);
        gg_host_free(s->fSplit); gg_host_free(s->fBmax); gg_host_free(s->iDim);
        s->fSplit = pinned(c * sizeof(double)); s->fBmax = pinned(c * sizeof(double)); s->iDim = pinned(c * sizeof(int));
        s->bnd = pinned(c * 6 * sizeof(double)); s->r = pinned(c * 3 * sizeof(double));
        s->fMass = pinned(c * sizeof(double)); s->fSoft = pinned(c * sizeof(double));
        s->fOpen2 = pinned(c * sizeof(double)); s->mom = pinned(c * GG_NMOM * sizeof(double));
        s->pLower = pinned(c * sizeof(int)); s->pUpper = pinned(c * sizeof(int));
        s->iLower = pinned(c * sizeof(int)); s->iUpper = pinned(c * sizeof(int));
        s->capNodes = c;
    }
    if (nPart > s->capPart) {
        size_t c = nPart + nPart / 4 + 16;
        gg_host_free(s->x); gg_host_free(s->y); gg_host_free(s->z); gg_host_free(s->m); gg_host_free(s->h);
        gg_host_free(s->a); gg_host_free(s->pot); gg_host_free(s->dt); gg_host_free(s->w); gg_host_free(s->active);
        gg_host_free(s->order);
        s->order = pinned(c * sizeof(int));
        s->x = pinned(c * sizeof(double)); s->y = pinned(c * sizeof(double)); s->z = pinned(c * sizeof(double));
        s->m = pinned(c * sizeof(double)); s->h = pinned(c * sizeof(double)); s->a = pinned(c * 3 * sizeof(double));
        s->pot = pinned(c * sizeof(double)); s->dt = pinned(c * sizeof(double)); s->w = pinned(c * sizeof(double));
        s->active = pinned(c * sizeof(int));
        s->capPart = c;
    }
}

static void die(const char *what);

/* Which GPU this rank drives: GG_SHIM_DEVICE, else rank modulo the number of devices. */
static int shim_device(PKD pkd) {
    const char *e = getenv("GG_SHIM_DEVICE");
    int nDev = 0;
    if (e && *e) return atoi(e);
    if (mdlThreads(pkd->mdl) == 1) return -1; /* the current device, as before */
    if (gg_device_count(&nDev) != GG_OK || nDev < 1) die("gg_device_count");
    return pkd->idSelf % nDev;
}

/* First multi-rank call: every rank joins the communicator.  COLLECTIVE over the MDL ranks. */
typedef struct {
    char id[GG_UNIQUE_ID_BYTES]; /* NCCL unique id ... */
    gg_group *grp;               /* ... or the in-process group (thread ranks) */
    int bLocal;
} COMMSEED;

static void shim_join(PKD pkd, SHIM *s) {
    static COMMSEED seed; /* rank 0's copy is the one the others read */
    COMMSEED mine, *p0;
    const int nT = mdlThreads(pkd->mdl);
    memset(&mine, 0, sizeof(mine));
    if (pkd->idSelf == 0) {
        const char *e = getenv("GG_SHIM_COMM");
        int nDev = 0;
        if (gg_device_count(&nDev) != GG_OK) die("gg_device_count");
        seed.bLocal = e ? strcmp(e, "local") == 0 : nDev < nT;
        if (seed.bLocal) {
            if (gg_group_create(&seed.grp, nT) != GG_OK) die("gg_group_create");
        } else if (gg_comm_unique_id(seed.id) != GG_OK) die("gg_comm_unique_id");
    }
    /* rank 0's seed through the host's MDL (works for every MDL flavour: the cache open is the barrier) */
    mdlROcache(pkd->mdl, CID_PARTICLE, &seed, sizeof(COMMSEED), 1);
    p0 = (COMMSEED *)mdlAquire(pkd->mdl, CID_PARTICLE, 0, 0);
    mine = *p0;
    mdlRelease(pkd->mdl, CID_PARTICLE, p0);
    mdlFinishCache(pkd->mdl, CID_PARTICLE);
    if (mine.bLocal) {
        if (gg_comm_init_local(s->ctx, mine.grp, pkd->idSelf) != GG_OK) die("gg_comm_init_local");
    } else if (gg_comm_init(s->ctx, mine.id, pkd->idSelf, nT) != GG_OK) die("gg_comm_init");
    s->bJoined = 1;
}

/* pkd->kdTop -> gg_set_top, and the ranks' root bounds for gg_exchange.  kdTop is a heap (ROOT = 1, LOWER(i) = 2i,
 * UPPER(i) = 2i + 1, pkd.h:77-86) in malloc'ed memory of which only the used cells were written (pkd.c:4385-4391), so it
 * is walked from ROOT: pLower >= 0 is a leaf = that rank's root cell, else an interior cell with both children. */
static void shim_top(PKD pkd, SHIM *s, double *bndAll) {
    const int nT = mdlThreads(pkd->mdl);
    int nCell = 2, i, j, top, stack[128];
    static const int kMax = 128;
    int pLower[128], bUsed[128];
    double r[3 * 128], fMass[128], fSoft[128], fOpen2[128], *mom;
    while (nCell < 2 * nT) nCell *= 2; /* master.c:4293: 2^(1 + ceil(log2 nThreads)) */
    mdlassert(pkd->mdl, nCell <= kMax);
    mom = (double *)calloc((size_t)GG_NMOM * nCell, sizeof(double));
    mdlassert(pkd->mdl, mom != NULL);
    memset(bUsed, 0, sizeof(bUsed));
    memset(pLower, 0xff, sizeof(pLower));
    memset(r, 0, sizeof(r)); memset(fMass, 0, sizeof(fMass)); memset(fSoft, 0, sizeof(fSoft)); memset(fOpen2, 0, sizeof(fOpen2));
    top = 0;
    stack[top++] = ROOT;
    while (top > 0) {
        const KDN *c;
        const struct pkdCalcCellStruct *q;
        double *mo;
        i = stack[--top];
        mdlassert(pkd->mdl, i < nCell);
        c = &pkd->kdTop[i];
        q = &c->mom;
        bUsed[i] = 1;
        pLower[i] = c->pLower;
        for (j = 0; j < 3; ++j) r[3 * i + j] = c->r[j];
        fMass[i] = c->fMass; fSoft[i] = c->fSoft; fOpen2[i] = c->fOpen2;
        mo = &mom[(size_t)GG_NMOM * i];
        mo[0] = q->Qxx; mo[1] = q->Qyy; mo[2] = q->Qzz; mo[3] = q->Qxy; mo[4] = q->Qxz; mo[5] = q->Qyz;
        mo[6] = q->Oxxx; mo[7] = q->Oxyy; mo[8] = q->Oxxy; mo[9] = q->Oyyy; mo[10] = q->Oxxz; mo[11] = q->Oyyz;
        mo[12] = q->Oxyz; mo[13] = q->Oxzz; mo[14] = q->Oyzz; mo[15] = q->Ozzz;
        mo[16] = q->Hxxxx; mo[17] = q->Hxyyy; mo[18] = q->Hxxxy; mo[19] = q->Hyyyy; mo[20] = q->Hxxxz;
        mo[21] = q->Hyyyz; mo[22] = q->Hxxyy; mo[23] = q->Hxxyz; mo[24] = q->Hxyyz; mo[25] = q->Hxxzz;
        mo[26] = q->Hxyzz; mo[27] = q->Hxzzz; mo[28] = q->Hyyzz; mo[29] = q->Hyzzz; mo[30] = q->Hzzzz;
        if (c->pLower >= 0) { /* a rank's root cell */
            mdlassert(pkd->mdl, c->pLower < nT);
            for (j = 0; j < 3; ++j) {
                bndAll[6 * c->pLower + j] = c->bnd.fMin[j];
                bndAll[6 * c->pLower + 3 + j] = c->bnd.fMax[j];
            }
        } else {
            stack[top++] = UPPER(i);
            stack[top++] = LOWER(i);
        }
    }
    if (gg_set_top(s->ctx, nCell, pLower, bUsed, r, fMass, fSoft, fOpen2, mom) != GG_OK) die("gg_set_top");
    free(mom);
}

void pkdGravAll(PKD pkd, int nReps, int bPeriodic, int iOrder, int bEwald, int iEwOrder, double fEwCut,
                double fEwhCut, int bComove, double dRhoFac, int bDoSun, double dSunSoft, double *aSun, int *nActive,
                double *pdPartSum, double *pdCellSum, double *pdSoftSum, CASTAT *pcs, double *pdFlop) {
    const int nNodes = pkd->nNodes, n = pkdLocal(pkd);
    SHIM *s;
    gg_tree t;
    gg_particles pp;
    gg_params prm;
    gg_stats st;
    PASS pass;
    int j, bResident;

    mdlassert(pkd->mdl, pkd->idSelf >= 0 && pkd->idSelf < 64);
    s = &g_shim[pkd->idSelf];
    g_shimRanks = mdlThreads(pkd->mdl); /* (every rank writes the same value) */
    if (!s->ctx && gg_create(&s->ctx, shim_device(pkd)) != GG_OK) die("gg_create");
    if (mdlThreads(pkd->mdl) > 1 && !s->bJoined) shim_join(pkd, s);
    reserve(s, (size_t)nNodes, (size_t)n);

    /* the timers the caller reads back (pst.c:3316-3324) */
    pkdClearTimer(pkd, 1);
    pkdClearTimer(pkd, 2);
    pkdClearTimer(pkd, 3);
    pkdStartTimer(pkd, 2);

    pass.pkd = pkd; pass.s = s;
    pass.bMom = getenv("GG_SHIM_HOST_MOMENTS") != NULL && atoi(getenv("GG_SHIM_HOST_MOMENTS")) != 0;
    bResident = 0;
    if (s->builtNodes && s->builtNodes == pkd->kdNodes && s->builtN == nNodes && !pass.bMom &&
        !(getenv("GG_SHIM_FORCE_UPLOAD") && atoi(getenv("GG_SHIM_FORCE_UPLOAD")))) { /* (measurement aid: always flatten) */
        /* kdNodes is the tree our pkdBuildBinary built and the device still holds it (same array, same size, root cell
         * equal bit for bit): nothing but the ACTIVE flags -- which msrActiveRung may have changed since -- goes up */
        double r[3], fMass;
        const KDN *c = &pkd->kdNodes[pkd->iRoot];
        bResident = gg_domain_summary(s->ctx, NULL, r, &fMass, NULL, NULL, NULL, NULL) == GG_OK && r[0] == c->r[0] &&
                    r[1] == c->r[1] && r[2] == c->r[2] && fMass == c->fMass;
    }
    if (bPeriodic && bEwald) {
        double root[GG_NROOT];
        const ILCN *R = &pkd->ilcnRoot;
        root[0] = R->m; root[1] = R->x; root[2] = R->y; root[3] = R->z;
        root[4] = R->xx; root[5] = R->yy; root[6] = R->xy; root[7] = R->xz; root[8] = R->yz; root[9] = R->zz;
        root[10] = R->xxx; root[11] = R->xyy; root[12] = R->xxy; root[13] = R->yyy; root[14] = R->xxz;
        root[15] = R->yyz; root[16] = R->xyz; root[17] = R->xzz; root[18] = R->yzz; root[19] = R->zzz;
        root[20] = R->xxxx; root[21] = R->xyyy; root[22] = R->xxxy; root[23] = R->yyyy; root[24] = R->xxxz;
        root[25] = R->yyyz; root[26] = R->xxyy; root[27] = R->xxyz; root[28] = R->xyyz; root[29] = R->xxzz;
        root[30] = R->xyzz; root[31] = R->xzzz; root[32] = R->yyzz; root[33] = R->yzzz; root[34] = R->zzzz;
        if (gg_set_root_moments(s->ctx, root) != GG_OK) die("gg_set_root_moments");
    }
    memset(&prm, 0, sizeof(prm));
    prm.nReps = nReps; prm.bPeriodic = bPeriodic; prm.iOrder = iOrder; prm.bEwald = bEwald; prm.iEwOrder = iEwOrder;
    prm.fEwCut = fEwCut; prm.fEwhCut = fEwhCut; prm.bComove = bComove; prm.dRhoFac = dRhoFac;
    for (j = 0; j < 3; ++j) prm.fPeriod[j] = pkd->fPeriod[j];
    prm.bDoSun = bDoSun; prm.dSunSoft = dSunSoft;
    /* what evaluation follows: with Ewald on, gg_set_local then sends the particles first and starts the correction of
     * every slice beside the remaining copies (pkdBucketEwald needs only pStore and ilcnRoot, ewald.c:15) */
    if (gg_announce(s->ctx, &prm) != GG_OK) die("gg_announce");
    if (bResident) {
        parallel_for((size_t)n, flatten_active, &pass);
        if (gg_set_active(s->ctx, s->active) != GG_OK) die("gg_set_active");
    } else if (!pass.bMom && !(getenv("GG_SHIM_SLICED") && !atoi(getenv("GG_SHIM_SLICED")))) {
        /* flatten and hand over in slices: while these threads flatten slice k + 1 the copy engine moves slice k, and the
         * Ewald correction of the slices that have landed already runs (GG_SHIM_SLICED=0: everything in one piece) */
        const size_t SLP = (size_t)1 << 19, SLN = (size_t)1 << 18;
        const KDN *root = &pkd->kdNodes[pkd->iRoot];
        double rb[6];
        size_t lo;
        for (j = 0; j < 3; ++j) { rb[j] = root->bnd.fMin[j]; rb[3 + j] = root->bnd.fMax[j]; }
        if (gg_local_begin(s->ctx, pkd->idSelf, nNodes, pkd->iRoot, n, rb, 1) != GG_OK) die("gg_local_begin");
        for (lo = 0; lo < (size_t)n; lo += SLP) {
            const size_t hi = lo + SLP < (size_t)n ? lo + SLP : (size_t)n;
            parallel_range(lo, hi, flatten_particles, &pass);
            if (gg_local_particles(s->ctx, (int)lo, (int)(hi - lo), s->x + lo, s->y + lo, s->z + lo, s->m + lo, s->h + lo,
                                   s->active + lo) != GG_OK) die("gg_local_particles");
        }
        for (lo = 0; lo < (size_t)nNodes; lo += SLN) {
            const size_t hi = lo + SLN < (size_t)nNodes ? lo + SLN : (size_t)nNodes;
            parallel_range(lo, hi, flatten_nodes, &pass);
            if (gg_local_nodes(s->ctx, (int)lo, (int)(hi - lo), s->r + 3 * lo, s->fMass + lo, s->fSoft + lo, s->fOpen2 + lo,
                               s->pLower + lo, s->pUpper + lo, s->iLower + lo, s->iUpper + lo) != GG_OK) die("gg_local_nodes");
        }
        if (gg_local_end(s->ctx) != GG_OK) die("gg_local_end");
        s->builtNodes = NULL;
    } else {
        parallel_for((size_t)nNodes, flatten_nodes, &pass);
        parallel_for((size_t)n, flatten_particles, &pass);
        t.nNodes = nNodes; t.iRoot = pkd->iRoot;
        t.bnd = s->bnd; t.r = s->r; t.fMass = s->fMass; t.fSoft = s->fSoft; t.fOpen2 = s->fOpen2;
        t.mom = pass.bMom ? s->mom : NULL;
        t.pLower = s->pLower; t.pUpper = s->pUpper; t.iLower = s->iLower; t.iUpper = s->iUpper;
        pp.n = n; pp.x = s->x; pp.y = s->y; pp.z = s->z; pp.fMass = s->m; pp.fSoft = s->h; pp.active = s->active;
        if (gg_set_local(s->ctx, pkd->idSelf, &t, &pp) != GG_OK) die("gg_set_local");
        s->builtNodes = NULL; /* the device now holds the host's arrays, not a tree it built itself (see pkdCalcRoot below) */
    }
    if (mdlThreads(pkd->mdl) > 1) {
        /* the host's top tree and, in place of pkdRemoteWalk's pulls, one collective push of pruned trees */
        double bndAll[6 * 64];
        shim_top(pkd, s, bndAll);
        if (gg_exchange(s->ctx, &prm, bndAll, NULL) != GG_OK) die("gg_exchange");
    }
    prm.accumulate = 0; /* this call's contribution, delivered zero-copy into the pinned arrays; merged by write_back */
    {
        /* the evaluation in GG_SHIM_CHUNKS pieces (default 8): the += pass over the 184-byte PARTICLE records of a finished
         * range runs on the host's cores while the GPU evaluates the next range */
        const char *e = getenv("GG_SHIM_CHUNKS");
        const int nChunks = e ? atoi(e) : 8;
        if (gg_gravity_chunked(s->ctx, &prm, s->a, s->pot, s->dt, s->w, &st, nChunks < 1 ? 1 : nChunks, write_back_chunk, &pass) != GG_OK)
            die("gg_gravity");
    }
    pkdStopTimer(pkd, 2);
    *nActive = st.nActive;
    *pdPartSum = st.dPartSum; *pdCellSum = st.dCellSum; *pdSoftSum = st.dSoftSum; *pdFlop = st.dFlop;
    if (aSun) { aSun[0] = st.aSun[0]; aSun[1] = st.aSun[1]; aSun[2] = st.aSun[2]; } /* zero without bDoSun */
    memset(pcs, 0, sizeof(*pcs)); /* no software cache on this path */
    pkd->nPart = st.nMaxPart; pkd->nCellSoft = st.nMaxCellSoft; pkd->nCellNewt = st.nMaxCellNewt; /* diag only */
}

/* ------------------------------------------------------------------------------------------------------------------
 * pkdBuildBinary (pkd.c:2627-2724) on the device.
 */
void pkdBuildBinary_cpu(PKD pkd, int nBucket, int iOpenType, double dCrit, int iOrder, int bTreeActiveOnly, int bGravity,
                        KDN *pRoot); /* the host's own, renamed by -DpkdBuildBinary=pkdBuildBinary_cpu */

static void gather_particles(void *arg, size_t lo, size_t hi) {
    PASS *a = (PASS *)arg;
    size_t i;
    for (i = lo; i < hi; ++i) a->tmp[i] = a->pkd->pStore[a->s->order[i]];
}

static void scatter_particles(void *arg, size_t lo, size_t hi) {
    PASS *a = (PASS *)arg;
    memcpy(&a->pkd->pStore[lo], &a->tmp[lo], (hi - lo) * sizeof(PARTICLE));
}

static void fill_nodes(void *arg, size_t lo, size_t hi) {
    PASS *a = (PASS *)arg;
    SHIM *s = a->s;
    size_t i;
    int j;
    for (i = lo; i < hi; ++i) {
        KDN *c = &a->pkd->kdNodes[i];
        struct pkdCalcCellStruct *q = &c->mom;
        const double *mo = &s->mom[(size_t)GG_NMOM * i];
        memset(c, 0, sizeof(KDN));
        c->iDim = s->iDim[i];
        c->fSplit = s->fSplit[i];
        for (j = 0; j < 3; ++j) {
            c->bnd.fMin[j] = s->bnd[6 * i + j];
            c->bnd.fMax[j] = s->bnd[6 * i + 3 + j];
            c->bndBall.fMin[j] = c->bnd.fMin[j]; /* fBallMax = 0 for gravity-only particles (pkd.c:2465-2466) */
            c->bndBall.fMax[j] = c->bnd.fMax[j];
            c->r[j] = s->r[3 * i + j];
        }
        c->pLower = s->pLower[i]; c->pUpper = s->pUpper[i]; c->iLower = s->iLower[i]; c->iUpper = s->iUpper[i];
        c->fMass = s->fMass[i]; c->fSoft = s->fSoft[i]; c->fOpen2 = s->fOpen2[i];
        q->Qxx = mo[0]; q->Qyy = mo[1]; q->Qzz = mo[2]; q->Qxy = mo[3]; q->Qxz = mo[4]; q->Qyz = mo[5];
        if (a->iOrder >= 3) {
            q->Oxxx = mo[6]; q->Oxyy = mo[7]; q->Oxxy = mo[8]; q->Oyyy = mo[9]; q->Oxxz = mo[10]; q->Oyyz = mo[11];
            q->Oxyz = mo[12]; q->Oxzz = mo[13]; q->Oyzz = mo[14]; q->Ozzz = mo[15];
        }
        if (a->iOrder >= 4) {
            q->Hxxxx = mo[16]; q->Hxyyy = mo[17]; q->Hxxxy = mo[18]; q->Hyyyy = mo[19]; q->Hxxxz = mo[20];
            q->Hyyyz = mo[21]; q->Hxxyy = mo[22]; q->Hxxyz = mo[23]; q->Hxyyz = mo[24]; q->Hxxzz = mo[25];
            q->Hxyzz = mo[26]; q->Hxzzz = mo[27]; q->Hyyzz = mo[28]; q->Hyzzz = mo[29]; q->Hzzzz = mo[30];
        }
        q->Bmax = s->fBmax[i]; /* B2..B6 (read by OPEN_ABSPAR only, pkd.c:2137-2179; that criterion is built by the host) stay 0 */
    }
}

void pkdBuildBinary(PKD pkd, int nBucket, int iOpenType, double dCrit, int iOrder, int bTreeActiveOnly, int bGravity,
                    KDN *pRoot) {
    const char *e = getenv("GG_SHIM_DEVICE_TREE");
    SHIM *s;
    PASS pass;
    gg_particles pp;
    int n, nNodes = 0;
    double t0 = 0.0;

    /* OPEN_ABSPAR (dRootBracket on B2..B6, pkd.c:2182-2252) stays with the host's build; OPEN_JOSH and the three criteria
     * whose radius is Bmax (pkd.c:2261-2264) are built on the device */
    if (!e || !atoi(e) || iOpenType == OPEN_ABSPAR || iOpenType < OPEN_JOSH || iOpenType > OPEN_RELTOT || !bGravity ||
        pkd->nLocal < 1 || nBucket > GG_MAX_BUCKET ||
        (bTreeActiveOnly && pkd->nTreeActive != pkd->nLocal)) {
        if (pkd->idSelf >= 0 && pkd->idSelf < 64) g_shim[pkd->idSelf].builtNodes = NULL;
        pkdBuildBinary_cpu(pkd, nBucket, iOpenType, dCrit, iOrder, bTreeActiveOnly, bGravity, pRoot);
        return;
    }
    lap(&t0, NULL);
    pkdActiveTypeOrder(pkd, TYPE_TREEACTIVE); /* pkd.c:2635 */
    lap(&t0, "build: pkdActiveTypeOrder");
    if (pkd->kdNodes) {
        mdlFinishCache(pkd->mdl, CID_CELL);
        mdlFree(pkd->mdl, pkd->kdNodes);
    }
    n = pkd->nLocal;
    mdlassert(pkd->mdl, pkd->idSelf >= 0 && pkd->idSelf < 64);
    s = &g_shim[pkd->idSelf];
    g_shimRanks = mdlThreads(pkd->mdl); /* (every rank writes the same value) */
    if (!s->ctx && gg_create(&s->ctx, shim_device(pkd)) != GG_OK) die("gg_create");
    reserve(s, 0, (size_t)n);
    pass.pkd = pkd; pass.s = s; pass.bMom = 1; pass.iOrder = iOrder; pass.tmp = NULL;
    parallel_for((size_t)n, flatten_particles, &pass);
    lap(&t0, "build: flatten particles");
    pp.n = n; pp.x = s->x; pp.y = s->y; pp.z = s->z; pp.fMass = s->m; pp.fSoft = s->h; pp.active = NULL;
    if (gg_build_local_open(s->ctx, pkd->idSelf, &pp, nBucket, iOpenType, dCrit, s->order, &nNodes, NULL) != GG_OK)
        die("gg_build_local_open");
    lap(&t0, "build: gg_build_local");
    reserve(s, (size_t)nNodes, (size_t)n);
    if (gg_tree_fetch(s->ctx, s->bnd, s->r, s->fMass, s->fSoft, s->fOpen2, s->mom, s->pLower, s->pUpper, s->iLower,
                      s->iUpper, NULL, NULL, NULL, NULL, NULL, NULL) != GG_OK) die("gg_tree_fetch");
    if (gg_tree_fetch_build(s->ctx, s->iDim, s->fSplit, s->fBmax) != GG_OK) die("gg_tree_fetch_build");
    lap(&t0, "build: fetch tree");
    /* pStore into tree order: the permutation BuildBinary's partitions would have applied */
    if ((size_t)n > s->capTmp) {
        free(s->tmp);
        s->capTmp = (size_t)n + (size_t)n / 4 + 16;
        s->tmp = (PARTICLE *)malloc(s->capTmp * sizeof(PARTICLE));
        mdlassert(pkd->mdl, s->tmp != NULL);
    }
    pass.tmp = s->tmp;
    parallel_for((size_t)n, gather_particles, &pass);
    parallel_for((size_t)n, scatter_particles, &pass);
    lap(&t0, "build: permute pStore");
    pkd->nNodes = nNodes;
    pkd->kdNodes = mdlMalloc(pkd->mdl, (size_t)(nNodes + 1) * sizeof(KDN)); /* + the extra cell, pkd.c:2668 */
    mdlassert(pkd->mdl, pkd->kdNodes != NULL);
    parallel_for((size_t)nNodes, fill_nodes, &pass);
    lap(&t0, "build: fill kdNodes");
    pkd->iFreeCell = nNodes;
    pkd->iRoot = 0;
    *pRoot = pkd->kdNodes[pkd->iRoot];
    mdlROcache(pkd->mdl, CID_CELL, pkd->kdNodes, sizeof(KDN), pkdNodes(pkd));
    s->builtNodes = pkd->kdNodes;
    s->builtN = nNodes;
}

/* ------------------------------------------------------------------------------------------------------------------
 * pkdCalcRoot (pkd.c:4395-4470): the complete l = 3, 4 moments of the rank's particles about its root centre, one host
 * pass over pStore (~25 ms per million particles).  When the tree in pkd->kdNodes is the one our pkdBuildBinary just built
 * -- same array, same size, and the device's root cell equal to the host's bit for bit -- they are read from the
 * device's raw moment record of the root instead (gg_domain_summary; same sums in another order, ~1e-15 relative).
 */
void pkdCalcRoot_cpu(PKD pkd, struct ilCellNewt *pcc); /* the host's own, renamed by -DpkdCalcRoot=pkdCalcRoot_cpu */

void pkdCalcRoot(PKD pkd, struct ilCellNewt *pcc) {
    SHIM *s = (pkd->idSelf >= 0 && pkd->idSelf < 64) ? &g_shim[pkd->idSelf] : NULL;
    if (s && s->ctx && s->builtNodes && s->builtNodes == pkd->kdNodes && s->builtN == pkd->nNodes) {
        double root[GG_NROOT], r[3], fMass;
        const KDN *c = &pkd->kdNodes[pkd->iRoot];
        if (gg_domain_summary(s->ctx, NULL, r, &fMass, NULL, NULL, NULL, root) == GG_OK && r[0] == c->r[0] &&
            r[1] == c->r[1] && r[2] == c->r[2] && fMass == c->fMass) {
            pcc->xxx = root[10]; pcc->xyy = root[11]; pcc->xxy = root[12]; pcc->yyy = root[13]; pcc->xxz = root[14];
            pcc->yyz = root[15]; pcc->xyz = root[16]; pcc->xzz = root[17]; pcc->yzz = root[18]; pcc->zzz = root[19];
            pcc->xxxx = root[20]; pcc->xyyy = root[21]; pcc->xxxy = root[22]; pcc->yyyy = root[23]; pcc->xxxz = root[24];
            pcc->yyyz = root[25]; pcc->xxyy = root[26]; pcc->xxyz = root[27]; pcc->xyyz = root[28]; pcc->xxzz = root[29];
            pcc->xyzz = root[30]; pcc->xzzz = root[31]; pcc->yyzz = root[32]; pcc->yzzz = root[33]; pcc->zzzz = root[34];
            return;
        }
    }
    pkdCalcRoot_cpu(pkd, pcc);
}
