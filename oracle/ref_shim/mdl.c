/*
 * mdl.c -- pthread stand-in for the external MDL (see mdl.h).  TEST INFRASTRUCTURE ONLY.
 *
 * "Ranks" are threads of one process (MDL_NTHREADS, default 1).  Services are RPCs through a one-slot
 * mailbox per receiver; the software caches of the real MDL are replaced by direct pointers into the
 * owner's arrays (mdlAquire = base + index*size), with a barrier at cache open/close so that nobody
 * reads an array before it is published or frees it while it is being read.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <sys/resource.h>
#include "mdl.h"

struct mdlShared {
    pthread_mutex_t mux;
    pthread_cond_t cv;
    int nThreads;
    int nBarrier, iBarrierGen;
    pthread_t *threads;
    void (*fcnChild)(MDL);
};

static void barrier(MDL mdl) {
    struct mdlShared *s = mdl->shared;
    int gen;
    if (s->nThreads == 1) return;
    pthread_mutex_lock(&s->mux);
    gen = s->iBarrierGen;
    if (++s->nBarrier == s->nThreads) {
        s->nBarrier = 0;
        ++s->iBarrierGen;
        pthread_cond_broadcast(&s->cv);
    } else {
        while (gen == s->iBarrierGen) pthread_cond_wait(&s->cv, &s->mux);
    }
    pthread_mutex_unlock(&s->mux);
}

static void *thread_main(void *v) {
    MDL mdl = v;
    struct mdlShared *s = mdl->shared;
    (*s->fcnChild)(mdl);
    return NULL;
}

int mdlInitialize(MDL *pmdl, char **argv, void (*fcnChild)(MDL)) {
    struct mdlShared *s = calloc(1, sizeof(*s));
    MDL *all;
    int i, n = 1;
    char *e = getenv("MDL_NTHREADS");
    (void)argv;
    if (e && atoi(e) > 0) n = atoi(e);
    pthread_mutex_init(&s->mux, NULL);
    pthread_cond_init(&s->cv, NULL);
    s->nThreads = n;
    s->fcnChild = fcnChild;
    s->threads = calloc(n, sizeof(pthread_t));
    all = calloc(n, sizeof(MDL));
    for (i = 0; i < n; ++i) {
        MDL m = calloc(1, sizeof(struct mdlContext));
        m->nThreads = n;
        m->idSelf = i;
        m->pmdl = all;
        m->shared = s;
        m->nMaxServices = MDL_MAX_SERVICES;
        m->psrv = calloc(MDL_MAX_SERVICES, sizeof(MDLSERVICE));
        m->pbReply = calloc(n, sizeof(int));
        m->pnReplyOut = calloc(n, sizeof(int));
        m->ppszReply = calloc(n, sizeof(char *));
        m->bDiag = getenv("MDL_DIAG") != NULL;
        all[i] = m;
    }
    for (i = 1; i < n; ++i) pthread_create(&s->threads[i], NULL, thread_main, all[i]);
    *pmdl = all[0];
    return n;
}

void mdlFinish(MDL mdl) {
    struct mdlShared *s = mdl->shared;
    int i;
    for (i = 1; i < s->nThreads; ++i) pthread_join(s->threads[i], NULL);
}

int mdlThreads(MDL mdl) { return mdl->nThreads; }
int mdlSelf(MDL mdl) { return mdl->idSelf; }

void mdlDiag(MDL mdl, char *psz) {
    if (mdl->bDiag) fputs(psz, stderr);
}

void mdlprintf(MDL mdl, const char *fmt, ...) {
    va_list ap;
    if (!mdl->bDiag) return;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

void mdlAddService(MDL mdl, int sid, void *p1, void (*fcn)(void *, void *, int, void *, int *), int nInBytes,
                   int nOutBytes) {
    assert(sid > 0 && sid < mdl->nMaxServices);
    mdl->psrv[sid].p1 = p1;
    mdl->psrv[sid].fcnService = fcn;
    mdl->psrv[sid].nInBytes = nInBytes;
    mdl->psrv[sid].nOutBytes = nOutBytes;
}

static int g_trace = -1;
#define TRACE(...) do { if (g_trace < 0) g_trace = getenv("MDL_TRACE") != NULL; if (g_trace) { fprintf(stderr, __VA_ARGS__); fflush(stderr);} } while (0)
void mdlReqService(MDL mdl, int id, int sid, void *vin, int nInBytes) {
    struct mdlShared *s = mdl->shared;
    TRACE("[%d] req -> %d sid %d\n", mdl->idSelf, id, sid);
    MDL t = mdl->pmdl[id];
    char *copy = NULL;
    if (nInBytes > 0) {
        copy = malloc(nInBytes);
        memcpy(copy, vin, nInBytes);
    }
    pthread_mutex_lock(&s->mux);
    while (t->bReq) pthread_cond_wait(&s->cv, &s->mux);
    t->bReq = 1;
    t->idReqFrom = mdl->idSelf;
    t->sidReq = sid;
    t->nReqIn = nInBytes;
    t->pszReqIn = copy;
    pthread_cond_broadcast(&s->cv);
    pthread_mutex_unlock(&s->mux);
}

void mdlGetReply(MDL mdl, int id, void *vout, int *pnOutBytes) {
    struct mdlShared *s = mdl->shared;
    pthread_mutex_lock(&s->mux);
    while (!mdl->pbReply[id]) pthread_cond_wait(&s->cv, &s->mux);
    if (vout && mdl->pnReplyOut[id] > 0) memcpy(vout, mdl->ppszReply[id], mdl->pnReplyOut[id]);
    if (pnOutBytes) *pnOutBytes = mdl->pnReplyOut[id];
    free(mdl->ppszReply[id]);
    mdl->ppszReply[id] = NULL;
    mdl->pbReply[id] = 0;
    pthread_cond_broadcast(&s->cv);
    pthread_mutex_unlock(&s->mux);
}

void mdlHandler(MDL mdl) {
    struct mdlShared *s = mdl->shared;
    for (;;) {
        int from, sid, nIn, nOut = 0;
        char *in, *out;
        MDL r;
        pthread_mutex_lock(&s->mux);
        while (!mdl->bReq) pthread_cond_wait(&s->cv, &s->mux);
        from = mdl->idReqFrom;
        sid = mdl->sidReq;
        nIn = mdl->nReqIn;
        in = mdl->pszReqIn;
        mdl->bReq = 0;
        pthread_cond_broadcast(&s->cv);
        pthread_mutex_unlock(&s->mux);
        if (sid == SRV_STOP) { /* msrFinish waits for a reply to the stop request too (master.c:3472-3473) */
            free(in);
            r = mdl->pmdl[from];
            pthread_mutex_lock(&s->mux);
            while (r->pbReply[mdl->idSelf]) pthread_cond_wait(&s->cv, &s->mux);
            r->ppszReply[mdl->idSelf] = NULL;
            r->pnReplyOut[mdl->idSelf] = 0;
            r->pbReply[mdl->idSelf] = 1;
            pthread_cond_broadcast(&s->cv);
            pthread_mutex_unlock(&s->mux);
            break;
        }
        assert(sid < mdl->nMaxServices && mdl->psrv[sid].fcnService);
        assert(nIn <= mdl->psrv[sid].nInBytes);
        out = calloc(mdl->psrv[sid].nOutBytes > 0 ? mdl->psrv[sid].nOutBytes : 1, 1); /* (services that fill nothing reply zeros) */
        (*mdl->psrv[sid].fcnService)(mdl->psrv[sid].p1, in, nIn, out, &nOut);
        assert(nOut <= mdl->psrv[sid].nOutBytes);
        free(in);
        r = mdl->pmdl[from];
        pthread_mutex_lock(&s->mux);
        while (r->pbReply[mdl->idSelf]) pthread_cond_wait(&s->cv, &s->mux);
        r->ppszReply[mdl->idSelf] = out;
        r->pnReplyOut[mdl->idSelf] = nOut;
        r->pbReply[mdl->idSelf] = 1;
        pthread_cond_broadcast(&s->cv);
        pthread_mutex_unlock(&s->mux);
    }
}

/*
 * Bilateral transfer (pkd.c:1497,1522): outgoing bytes sit at the TOP of vBuf, incoming bytes land at the bottom.
 * A real MDL moves the two directions piece by piece, so the space an outgoing piece vacates takes incoming data: the
 * exchange is complete whenever each side's WHOLE buffer can hold what the other sends (pkdSwapAll, pkd.c:1504-1527,
 * relies on it: its buffer is full of outgoing particles and it asserts completeness).  Otherwise each side accepts
 * what fits its free space and the caller comes back for the rest (pkdSwapRejects).  Both directions go through a
 * temporary copy, so neither side writes its buffer before the partner has read it.
 */
int mdlSwap(MDL mdl, int id, size_t nBufBytes, void *vBuf, size_t nOutBytes, size_t *pnSndBytes,
            size_t *pnRcvBytes) {
    struct mdlShared *s = mdl->shared;
    MDL o = mdl->pmdl[id];
    size_t nIn, nSnd, nOtherOut, nOtherBuf, nOtherDone0;
    char *tmp;
    int bFull;
    TRACE("[%d] swap with %d buf %zu out %zu\n", mdl->idSelf, id, nBufBytes, nOutBytes);
    pthread_mutex_lock(&s->mux);
    mdl->pSwapBuf = vBuf;
    mdl->nSwapBuf = nBufBytes;
    mdl->nSwapOut = nOutBytes;
    mdl->idSwapWith = id;
    mdl->iSwapState = 1; /* published */
    pthread_cond_broadcast(&s->cv);
    /* the partner may already have copied (state 2) by the time we look: >= 1, not == 1 */
    while (!(o->iSwapState >= 1 && o->idSwapWith == mdl->idSelf)) pthread_cond_wait(&s->cv, &s->mux);
    nOtherDone0 = o->nSwapTaken; /* reused as a monotonic "swaps completed" counter */
    pthread_mutex_unlock(&s->mux);
    nOtherOut = o->nSwapOut;
    nOtherBuf = o->nSwapBuf;
    bFull = nOtherOut <= nBufBytes && nOutBytes <= nOtherBuf; /* both sides evaluate the same condition */
    if (bFull) {
        nIn = nOtherOut;
        nSnd = nOutBytes;
    } else {
        nIn = nOtherOut < nBufBytes - nOutBytes ? nOtherOut : nBufBytes - nOutBytes;
        nSnd = nOutBytes < nOtherBuf - nOtherOut ? nOutBytes : nOtherBuf - nOtherOut;
    }
    /* pull the FIRST nIn bytes of the partner's outgoing region (its top) into a temporary */
    tmp = malloc(nIn ? nIn : 1);
    assert(tmp != NULL);
    memcpy(tmp, o->pSwapBuf + (nOtherBuf - nOtherOut), nIn);
    pthread_mutex_lock(&s->mux);
    mdl->iSwapState = 2; /* copied: the partner may now overwrite its buffer */
    pthread_cond_broadcast(&s->cv);
    while (!(o->iSwapState == 2 || o->nSwapTaken != nOtherDone0)) pthread_cond_wait(&s->cv, &s->mux);
    pthread_mutex_unlock(&s->mux);
    /* the partner has read what it takes from us -- the FIRST nSnd bytes of our outgoing region, so what it did not
     * take is already at the very top, where pkdSwapRejects expects the remaining rejects; incoming bytes go to the bottom */
    memcpy(vBuf, tmp, nIn);
    free(tmp);
    pthread_mutex_lock(&s->mux);
    mdl->iSwapState = 0;
    mdl->idSwapWith = -1;
    mdl->nSwapTaken += 1;
    pthread_cond_broadcast(&s->cv);
    pthread_mutex_unlock(&s->mux);
    *pnSndBytes = nSnd;
    *pnRcvBytes = nIn;
    TRACE("[%d] swap done snd %zu rcv %zu\n", mdl->idSelf, nSnd, nIn);
    return (nSnd == nOutBytes && nIn == nOtherOut);
}

void *mdlMalloc(MDL mdl, size_t iSize) {
    (void)mdl;
    return malloc(iSize);
}
void mdlFree(MDL mdl, void *p) {
    (void)mdl;
    free(p);
}

void mdlROcache(MDL mdl, int cid, void *pData, int iDataSize, int nData) {
    assert(cid >= 0 && cid < MDL_MAX_CACHE);
    mdl->cache[cid].iType = 1;
    mdl->cache[cid].pData = pData;
    mdl->cache[cid].iDataSize = iDataSize;
    mdl->cache[cid].nData = nData;
    mdl->cache[cid].nAccess = 0;
    barrier(mdl);
}

void mdlCOcache(MDL mdl, int cid, void *pData, int iDataSize, int nData, void (*init)(void *),
                void (*combine)(void *, void *)) {
    (void)init;
    (void)combine;
    /* combiner caches are only used by the SPH/smooth path, which is outside the oracle's scope */
    assert(mdl->nThreads == 1);
    mdlROcache(mdl, cid, pData, iDataSize, nData);
}

void mdlFinishCache(MDL mdl, int cid) {
    barrier(mdl);
    mdl->cache[cid].iType = 0;
}

void mdlCacheCheck(MDL mdl) { (void)mdl; }

void *mdlAquire(MDL mdl, int cid, int iIndex, int id) {
    MDLCACHE *c = &mdl->pmdl[id]->cache[cid];
    mdl->cache[cid].nAccess += 1;
    return c->pData + (size_t)iIndex * c->iDataSize;
}
void mdlRelease(MDL mdl, int cid, void *p) {
    (void)mdl;
    (void)cid;
    (void)p;
}
double mdlNumAccess(MDL mdl, int cid) { return mdl->cache[cid].nAccess; }
double mdlMissRatio(MDL mdl, int cid) { (void)mdl; (void)cid; return 0.0; }
double mdlCollRatio(MDL mdl, int cid) { (void)mdl; (void)cid; return 0.0; }
double mdlMinRatio(MDL mdl, int cid) { (void)mdl; (void)cid; return 0.0; }

double mdlCpuTimer(MDL mdl) {
    struct rusage ru;
    (void)mdl;
    getrusage(RUSAGE_THREAD, &ru);
    return ru.ru_utime.tv_sec + 1e-6 * ru.ru_utime.tv_usec;
}
static double wall(void) {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}
void mdlZeroTimer(MDL mdl, mdlTimer *t) {
    struct rusage ru;
    (void)mdl;
    getrusage(RUSAGE_THREAD, &ru);
    t->wallclock = wall();
    t->cpu = ru.ru_utime.tv_sec + 1e-6 * ru.ru_utime.tv_usec;
    t->system = ru.ru_stime.tv_sec + 1e-6 * ru.ru_stime.tv_usec;
}
void mdlGetTimer(MDL mdl, mdlTimer *t0, mdlTimer *t) {
    mdlTimer now;
    mdlZeroTimer(mdl, &now);
    t->wallclock = now.wallclock - t0->wallclock;
    t->cpu = now.cpu - t0->cpu;
    t->system = now.system - t0->system;
}
void mdlPrintTimer(MDL mdl, char *message, mdlTimer *t0) {
    mdlTimer t;
    if (!mdl->bDiag) return;
    mdlGetTimer(mdl, t0, &t);
    fprintf(stderr, "%s %f %f %f\n", message, t.wallclock, t.cpu, t.system);
}
