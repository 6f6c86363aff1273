/*
 * mdl.h -- stand-in for the EXTERNAL machine-dependent layer (N-BodyShop/mdl, not vendored in the
 * reference tree; README.md:22-27 says to clone it next to gasoline).  TEST INFRASTRUCTURE ONLY: it lets
 * the oracle recipe (oracle/Makefile) compile the reference's own sources where they lie under
 * /root/reference.  It moves bytes between "ranks" that are pthreads of one process; it does no force math.
 *
 * API inferred from the reference's call sites: main.c:96, pst.c:13-563,
 * pkd.c:1497 (mdlSwap), pkd.c:2722/2900 (mdlROcache), walk.c:196/231 (mdlAquire), smooth.c:488 (mdlCOcache).
 * Thread count: environment variable MDL_NTHREADS (default 1 == the "null" MDL).
 */
#ifndef MDL_HINCLUDED
#define MDL_HINCLUDED
#include <stdio.h>
#include <stddef.h>
#include <assert.h>

#define SRV_STOP 0
#define MDL_MAX_SERVICES 512
#define MDL_MAX_CACHE 4

typedef struct {
    double wallclock, cpu, system;
} mdlTimer;

typedef struct mdlService {
    int nInBytes, nOutBytes;
    void *p1;
    void (*fcnService)(void *, void *, int, void *, int *);
} MDLSERVICE;

typedef struct mdlCacheSpace {
    int iType; /* 0 none, 1 RO, 2 CO */
    char *pData;
    int iDataSize, nData;
    double nAccess;
} MDLCACHE;

typedef struct mdlContext {
    int nThreads, idSelf;
    struct mdlContext **pmdl; /* all contexts, shared */
    void *shared;             /* struct mdlShared * */
    int nMaxServices;
    MDLSERVICE *psrv;
    MDLCACHE cache[MDL_MAX_CACHE];
    /* request mailbox (one outstanding request per receiver is enough for the PST fan-out) */
    int bReq, idReqFrom, sidReq, nReqIn;
    char *pszReqIn;
    /* reply slots, indexed by the thread the reply comes FROM */
    int *pbReply, *pnReplyOut;
    char **ppszReply;
    /* swap rendezvous */
    int iSwapState, idSwapWith;
    char *pSwapBuf;
    size_t nSwapBuf, nSwapOut, nSwapTaken;
    FILE *fpDiag;
    int bDiag;
} *MDL;

#define mdlassert(mdl, expr) assert(expr)

int mdlInitialize(MDL *pmdl, char **argv, void (*fcnChild)(MDL));
void mdlFinish(MDL mdl);
int mdlThreads(MDL mdl);
int mdlSelf(MDL mdl);
int mdlSwap(MDL mdl, int id, size_t nBufBytes, void *vBuf, size_t nOutBytes, size_t *pnSndBytes,
            size_t *pnRcvBytes);
void mdlDiag(MDL mdl, char *psz);
void mdlprintf(MDL mdl, const char *fmt, ...);
void mdlAddService(MDL mdl, int sid, void *p1, void (*fcn)(void *, void *, int, void *, int *), int nInBytes,
                   int nOutBytes);
void mdlReqService(MDL mdl, int id, int sid, void *vin, int nInBytes);
void mdlGetReply(MDL mdl, int id, void *vout, int *pnOutBytes);
void mdlHandler(MDL mdl);
void *mdlMalloc(MDL mdl, size_t iSize);
void mdlFree(MDL mdl, void *p);
void mdlROcache(MDL mdl, int cid, void *pData, int iDataSize, int nData);
void mdlCOcache(MDL mdl, int cid, void *pData, int iDataSize, int nData, void (*init)(void *),
                void (*combine)(void *, void *));
void mdlFinishCache(MDL mdl, int cid);
void mdlCacheCheck(MDL mdl);
void *mdlAquire(MDL mdl, int cid, int iIndex, int id);
void mdlRelease(MDL mdl, int cid, void *p);
double mdlNumAccess(MDL mdl, int cid);
double mdlMissRatio(MDL mdl, int cid);
double mdlCollRatio(MDL mdl, int cid);
double mdlMinRatio(MDL mdl, int cid);
double mdlCpuTimer(MDL mdl);
void mdlZeroTimer(MDL mdl, mdlTimer *t);
void mdlGetTimer(MDL mdl, mdlTimer *t0, mdlTimer *t);
void mdlPrintTimer(MDL mdl, char *message, mdlTimer *t0);
#endif
