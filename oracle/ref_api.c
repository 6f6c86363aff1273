/*
 * ref_api.c -- ctypes-friendly driver around the UNMODIFIED reference objects (oracle/_ref/libgasref.so).
 *
 * TEST INFRASTRUCTURE ONLY (oracle): compiled by oracle/Makefile against the headers in /root/reference,
 * never linked or loaded by the product path.  It calls the reference's own pstBuildTree / pstColCells /
 * pstDistribCells / pstCalcRoot / pstDistribRoot / pstGravity on a single-rank PST, i.e. exactly the
 * sequence msrBuildTree (master.c:4249-4310) + msrGravity (master.c:5869) run, and exports
 *   - the reference's tree, field by field (never assuming raw KDN layout),
 *   - per-bucket interaction-list counts, captured with a linker --wrap around pkdBucketWalk
 *     (the reference itself only reports sums, pkd.c:2945-2949),
 *   - per-particle a, fPot, dtGrav, fWeight after pkdGravAll (pkd.c:2868).
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <sys/time.h>
#include "pst.h"
#include "pkd.h"
#include "walk.h"
#include "grav.h"
#include "ewald.h"
#include "opentype.h"
#ifdef REF_GPU_HOST
#include "gasoline_b200.h"
gg_context *pkdGravAllContext(int idSelf); /* gasoline_b200/csrc/pkd_gravall_shim.c */
#endif

typedef struct {
    MDL mdl;
    PST pst;
    LCL lcl;
    PKD pkd;
    int n;
} REF;

/* ---- per-bucket counts: --wrap=pkdBucketWalk -------------------------------------------------- */
static int *g_counts = NULL; /* 3 ints per node, indexed by iBucket */
static int g_nCounts = 0;
static int *g_sunCounts = NULL; /* the lists of the bDoSun dummy bucket (pkd.c:3003-3041), iBucket == pkd->iFreeCell */
void __real_pkdBucketWalk(PKD pkd, int iBucket, int nReps, int iOrder);
/*
 * Multi-rank dump (gasoline_ref binary, pthread MDL ranks; env REF_DUMP=<prefix>): on a rank's first pkdBucketWalk
 * write <prefix>.rank<id>: ints nThreads idSelf nLocal nNodes iRoot nTopCells; per particle (tree order) iOrder;
 * kdTop[0..nTop) as records of (int pLower, int pUpper, doubles r[3] fMass fSoft fOpen2 mom[31]); ilcnRoot (35
 * doubles); then one record (iBucket pLower pUpper nPart nCellSoft nCellNewt) per bucket walked.
 */
#define REF_MAX_RANKS 64
static FILE *g_dump[REF_MAX_RANKS];
static int g_dumpStep[REF_MAX_RANKS]; /* REF_DUMP_ALL: one dump per force evaluation, <prefix>.s<k>.rank<id> */
static void dump_header(PKD pkd) {
    const char *prefix = getenv("REF_DUMP");
    char name[512];
    FILE *f;
    int i, k, hdr[6], nTop;
    if (getenv("REF_DUMP_ALL")) snprintf(name, sizeof(name), "%s.s%d.rank%d", prefix, g_dumpStep[pkd->idSelf], pkd->idSelf);
    else snprintf(name, sizeof(name), "%s.rank%d", prefix, pkd->idSelf);
    f = fopen(name, "wb");
    assert(f);
    g_dump[pkd->idSelf] = f;
    nTop = 1;
    while (nTop < 2 * mdlThreads(pkd->mdl)) nTop *= 2; /* master.c:4293: 2^(1+ceil(log2 nThreads)) */
    if (mdlThreads(pkd->mdl) == 1) nTop = 2;
    hdr[0] = mdlThreads(pkd->mdl); hdr[1] = pkd->idSelf; hdr[2] = pkd->nLocal; hdr[3] = pkd->nNodes;
    hdr[4] = pkd->iRoot; hdr[5] = nTop;
    fwrite(hdr, sizeof(int), 6, f);
    for (i = 0; i < pkd->nLocal; ++i) fwrite(&pkd->pStore[i].iOrder, sizeof(int), 1, f);
    for (i = 0; i < nTop; ++i) {
        const KDN *c = &pkd->kdTop[i];
        const double *mo = (const double *)&c->mom;
        double rec[6 + 31];
        int ii[2];
        ii[0] = c->pLower; ii[1] = c->pUpper;
        rec[0] = c->r[0]; rec[1] = c->r[1]; rec[2] = c->r[2]; rec[3] = c->fMass; rec[4] = c->fSoft; rec[5] = c->fOpen2;
        for (k = 0; k < 31; ++k) rec[6 + k] = mo[k];
        fwrite(ii, sizeof(int), 2, f);
        fwrite(rec, sizeof(double), 37, f);
    }
    fwrite(&pkd->ilcnRoot, sizeof(double), 35, f);
    fflush(f);
}

void __wrap_pkdBucketWalk(PKD pkd, int iBucket, int nReps, int iOrder) {
    __real_pkdBucketWalk(pkd, iBucket, nReps, iOrder);
    if (getenv("REF_DUMP") && pkd->idSelf < REF_MAX_RANKS) {
        int rec[6];
        if (!g_dump[pkd->idSelf]) dump_header(pkd);
        rec[0] = iBucket; rec[1] = pkd->kdNodes[iBucket].pLower; rec[2] = pkd->kdNodes[iBucket].pUpper;
        rec[3] = pkd->nPart; rec[4] = pkd->nCellSoft; rec[5] = pkd->nCellNewt;
        fwrite(rec, sizeof(int), 6, g_dump[pkd->idSelf]);
        fflush(g_dump[pkd->idSelf]);
    }
    if (g_sunCounts && iBucket == pkd->iFreeCell) {
        g_sunCounts[0] = pkd->nPart;
        g_sunCounts[1] = pkd->nCellSoft;
        g_sunCounts[2] = pkd->nCellNewt;
    }
    if (g_counts && iBucket < g_nCounts) {
        g_counts[3 * iBucket + 0] = pkd->nPart;
        g_counts[3 * iBucket + 1] = pkd->nCellSoft;
        g_counts[3 * iBucket + 2] = pkd->nCellNewt;
    }
}

/* After the rank's pkdGravAll: marker record (iBucket = -1) + per-particle a[3], fPot, dtGrav, fWeight (tree order). */
void __real_pkdGravAll(PKD pkd, int nReps, int bPeriodic, int iOrder, int bEwald, int iEwOrder, double fEwCut,
                       double fEwhCut, int bComove, double dRhoFac, int bDoSun, double dSunSoft, double *aSun,
                       int *nActive, double *pdPartSum, double *pdCellSum, double *pdSoftSum, CASTAT *pcs, double *pdFlop);
void __wrap_pkdGravAll(PKD pkd, int nReps, int bPeriodic, int iOrder, int bEwald, int iEwOrder, double fEwCut,
                       double fEwhCut, int bComove, double dRhoFac, int bDoSun, double dSunSoft, double *aSun,
                       int *nActive, double *pdPartSum, double *pdCellSum, double *pdSoftSum, CASTAT *pcs, double *pdFlop) {
    __real_pkdGravAll(pkd, nReps, bPeriodic, iOrder, bEwald, iEwOrder, fEwCut, fEwhCut, bComove, dRhoFac, bDoSun,
                      dSunSoft, aSun, nActive, pdPartSum, pdCellSum, pdSoftSum, pcs, pdFlop);
#ifdef REF_GPU_HOST
    /* the drop-in host: pkdGravAll is the product's shim, which never calls pkdBucketWalk -- the per-bucket records of
     * the dump come from the library's counters (gg_bucket_counts), in the same record format */
    if (getenv("REF_DUMP") && pkd->idSelf < REF_MAX_RANKS && !g_dump[pkd->idSelf]) {
        gg_context *ctx = pkdGravAllContext(pkd->idSelf);
        int *cnt = malloc((size_t)pkd->nNodes * 3 * sizeof(int)), i;
        assert(ctx && cnt);
        dump_header(pkd);
        if (gg_bucket_counts(ctx, cnt) != 0) abort();
        for (i = 0; i < pkd->nNodes; ++i) {
            int rec[6];
            if (pkd->kdNodes[i].iLower != -1 || cnt[3 * i] < 0) continue;
            rec[0] = i; rec[1] = pkd->kdNodes[i].pLower; rec[2] = pkd->kdNodes[i].pUpper;
            rec[3] = cnt[3 * i]; rec[4] = cnt[3 * i + 1]; rec[5] = cnt[3 * i + 2];
            fwrite(rec, sizeof(int), 6, g_dump[pkd->idSelf]);
        }
        free(cnt);
    }
#endif
    if (getenv("REF_DUMP") && pkd->idSelf < REF_MAX_RANKS && g_dump[pkd->idSelf]) {
        FILE *f = g_dump[pkd->idSelf];
        int rec[6] = {-1, 0, 0, 0, 0, 0}, i;
        double sums[4];
        fwrite(rec, sizeof(int), 6, f);
        sums[0] = *pdPartSum; sums[1] = *pdCellSum; sums[2] = *pdSoftSum; sums[3] = *pdFlop;
        fwrite(sums, sizeof(double), 4, f);
        for (i = 0; i < pkd->nLocal; ++i) {
            const PARTICLE *p = &pkd->pStore[i];
            double v[6];
            v[0] = p->a[0]; v[1] = p->a[1]; v[2] = p->a[2]; v[3] = p->fPot; v[4] = p->dtGrav; v[5] = p->fWeight;
            fwrite(v, sizeof(double), 6, f);
        }
        if (getenv("REF_DUMP_ALL")) { /* the positions this evaluation (and the decomposition before it) saw */
            for (i = 0; i < pkd->nLocal; ++i) {
                const PARTICLE *p = &pkd->pStore[i];
                double v[3];
                v[0] = p->r[0]; v[1] = p->r[1]; v[2] = p->r[2];
                fwrite(v, sizeof(double), 3, f);
            }
            ++g_dumpStep[pkd->idSelf];
        }
        fclose(f);
        g_dump[pkd->idSelf] = NULL;
    }
}

static double wallclock(void) {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

REF *ref_create(int n, const double *x, const double *y, const double *z, const double *m, const double *h,
                const int *active, const double *fPeriod) {
    REF *r = calloc(1, sizeof(REF));
    FLOAT per[3];
    int i;
    char *argv[2] = {"ref", NULL};
    mdlInitialize(&r->mdl, argv, NULL);
    assert(mdlThreads(r->mdl) == 1);
    r->lcl.pszDataPath = NULL;
    r->lcl.pkd = NULL;
    pstInitialize(&r->pst, r->mdl, &r->lcl);
    for (i = 0; i < 3; ++i) per[i] = fPeriod[i];
    pkdInitialize(&r->pkd, r->mdl, 4, n, 1, per, -FLOAT_MAXVAL, FLOAT_MAXVAL, n, 0, 0);
    r->lcl.pkd = r->pkd;
    r->n = n;
    memset(r->pkd->pStore, 0, (size_t)(n + 1) * sizeof(PARTICLE));
    for (i = 0; i < n; ++i) {
        PARTICLE *p = &r->pkd->pStore[i];
        p->iOrder = i;
        p->iActive = TYPE_DARK | TYPE_TREEACTIVE | ((!active || active[i]) ? TYPE_ACTIVE : 0);
        p->fMass = m[i];
        p->fSoft = h[i];
        p->fSoft0 = h[i];
        p->r[0] = x[i];
        p->r[1] = y[i];
        p->r[2] = z[i];
        p->fWeight = 1.0;
    }
    r->pkd->nLocal = n;
    r->pkd->nActive = n;
    r->pkd->nTreeActive = n;
    return r;
}

void ref_destroy(REF *r) {
    pstFinish(r->pst);
    free(r);
}

/* msrBuildTree (master.c:4249) for one rank with any opening criterion of opentype.h:5-9 (dCrit = theta for OPEN_JOSH,
 * the error bound for OPEN_ABSPAR; master.c:1897-1934). */
double ref_build_tree_open(REF *r, int nBucket, int iOpenType, double dCrit, int iOrder);

/* ... iOpenType = OPEN_JOSH, dCrit = theta: the default of a run that sets dTheta. */
double ref_build_tree(REF *r, int nBucket, double dTheta, int iOrder) {
    return ref_build_tree_open(r, nBucket, OPEN_JOSH, dTheta, iOrder);
}

double ref_build_tree_open(REF *r, int nBucket, int iOpenType, double dCrit, int iOrder) {
    struct inBuildTree in;
    struct outBuildTree out;
    struct inColCells inc;
    struct ioCalcRoot root;
    KDN *pkdn;
    int iDum, nCell;
    double t0 = wallclock();

    pkdActiveTypeOrder(r->pkd, TYPE_ACTIVE | TYPE_TREEACTIVE); /* msrActiveTypeOrder, master.c:4263 */
    in.nBucket = nBucket;
    in.iOpenType = iOpenType;
    in.iOrder = iOrder;
    in.dCrit = dCrit;
    in.bActiveOnly = 0;
    in.bTreeActiveOnly = 0;
    in.bBinary = 1;
    in.bGravity = 1;
    pstBuildTree(r->pst, &in, sizeof(in), &out, &iDum);
    nCell = 1 << (1 + (int)ceil(log((double)1) / log(2.0)));
    pkdn = malloc(nCell * sizeof(KDN));
    inc.iCell = ROOT;
    inc.nCell = nCell;
    pstColCells(r->pst, &inc, sizeof(inc), pkdn, NULL);
    pstDistribCells(r->pst, pkdn, nCell * sizeof(KDN), NULL, NULL);
    free(pkdn);
    pstCalcRoot(r->pst, NULL, 0, &root, &iDum);
    pstDistribRoot(r->pst, &root, sizeof(struct ioCalcRoot), NULL, NULL);
    return wallclock() - t0;
}

int ref_num_nodes(REF *r) { return r->pkd->nNodes; }
int ref_root(REF *r) { return r->pkd->iRoot; }

/* Field-wise export of kdNodes[0..nNodes) into SoA arrays (mom: 31 doubles Q6,O10,H15; bmom: Bmax,B2..B6). */
void ref_export_nodes(REF *r, double *bnd, double *rc, double *fMass, double *fSoft, double *fOpen2,
                      double *mom, double *bmom, int *pLower, int *pUpper, int *iLower, int *iUpper,
                      int *iDim) {
    int i, j, nn = r->pkd->nNodes;
    for (i = 0; i < nn; ++i) {
        KDN *c = &r->pkd->kdNodes[i];
        double *q = &mom[31 * (size_t)i];
        for (j = 0; j < 3; ++j) {
            bnd[6 * (size_t)i + j] = c->bnd.fMin[j];
            bnd[6 * (size_t)i + 3 + j] = c->bnd.fMax[j];
            rc[3 * (size_t)i + j] = c->r[j];
        }
        fMass[i] = c->fMass;
        fSoft[i] = c->fSoft;
        fOpen2[i] = c->fOpen2;
        q[0] = c->mom.Qxx; q[1] = c->mom.Qyy; q[2] = c->mom.Qzz;
        q[3] = c->mom.Qxy; q[4] = c->mom.Qxz; q[5] = c->mom.Qyz;
        q[6] = c->mom.Oxxx; q[7] = c->mom.Oxyy; q[8] = c->mom.Oxxy; q[9] = c->mom.Oyyy;
        q[10] = c->mom.Oxxz; q[11] = c->mom.Oyyz; q[12] = c->mom.Oxyz; q[13] = c->mom.Oxzz;
        q[14] = c->mom.Oyzz; q[15] = c->mom.Ozzz;
        q[16] = c->mom.Hxxxx; q[17] = c->mom.Hxyyy; q[18] = c->mom.Hxxxy; q[19] = c->mom.Hyyyy;
        q[20] = c->mom.Hxxxz; q[21] = c->mom.Hyyyz; q[22] = c->mom.Hxxyy; q[23] = c->mom.Hxxyz;
        q[24] = c->mom.Hxyyz; q[25] = c->mom.Hxxzz; q[26] = c->mom.Hxyzz; q[27] = c->mom.Hxzzz;
        q[28] = c->mom.Hyyzz; q[29] = c->mom.Hyzzz; q[30] = c->mom.Hzzzz;
        if (bmom) {
            double *b = &bmom[6 * (size_t)i];
            b[0] = c->mom.Bmax; b[1] = c->mom.B2; b[2] = c->mom.B3;
            b[3] = c->mom.B4; b[4] = c->mom.B5; b[5] = c->mom.B6;
        }
        pLower[i] = c->pLower;
        pUpper[i] = c->pUpper;
        iLower[i] = c->iLower;
        iUpper[i] = c->iUpper;
        if (iDim) iDim[i] = c->iDim;
    }
}

/* Particles in the reference's tree order. */
void ref_export_particles(REF *r, int *iOrder, double *x, double *y, double *z, double *m, double *h,
                          int *active) {
    int i;
    for (i = 0; i < r->n; ++i) {
        PARTICLE *p = &r->pkd->pStore[i];
        iOrder[i] = p->iOrder;
        x[i] = p->r[0];
        y[i] = p->r[1];
        z[i] = p->r[2];
        m[i] = p->fMass;
        h[i] = p->fSoft;
        if (active) active[i] = TYPEQueryACTIVE(p) ? 1 : 0;
    }
}

/*
 * After ref_build_tree: set the ACTIVE bit from active[] given in TREE order (what msrActiveRung does between
 * the build and msrGravity when only some rungs are kicked, master.c:8403-8420).  Used by bench.py's CPU legs to
 * let P processes each evaluate a disjoint slice of the sink buckets with the reference's own pkdGravAll.
 */
void ref_set_active_tree(REF *r, const int *active) {
    int i, n = 0;
    for (i = 0; i < r->n; ++i) {
        PARTICLE *p = &r->pkd->pStore[i];
        if (active[i]) {
            TYPESet(p, TYPE_ACTIVE);
            ++n;
        } else
            TYPEReset(p, TYPE_ACTIVE);
    }
    r->pkd->nActive = n;
}

/* pkd->ilcnRoot in ILCN field order (pkd.h:481-494): m,x,y,z,xx,yy,xy,xz,yz,zz, 10 octopole, 15 hexadecapole */
void ref_export_root(REF *r, double *out35) { memcpy(out35, &r->pkd->ilcnRoot, 35 * sizeof(double)); }

int ref_ewald_table(REF *r, double fhCut, int iOrder, double *ewt5, int nMax) {
    int i;
    pkdEwaldInit(r->pkd, fhCut, iOrder);
    for (i = 0; i < r->pkd->nEwhLoop && i < nMax; ++i) {
        ewt5[5 * i + 0] = r->pkd->ewt[i].hx;
        ewt5[5 * i + 1] = r->pkd->ewt[i].hy;
        ewt5[5 * i + 2] = r->pkd->ewt[i].hz;
        ewt5[5 * i + 3] = r->pkd->ewt[i].hCfac;
        ewt5[5 * i + 4] = r->pkd->ewt[i].hSfac;
    }
    return r->pkd->nEwhLoop;
}

/*
 * pkdInitAccel (pkd.c:5443) + pstGravity -> pkdGravAll (pst.c:3248, pkd.c:2868).
 * counts: 3 ints per node (nPart,nCellSoft,nCellNewt), filled for buckets with an active sink, else -1.
 * stats[0..7] = nActive, dPartSum, dCellSum, dSoftSum, dFlop, wallclock seconds of pstGravity, 0, 0.
 * Outputs are in the reference's tree order (see ref_export_particles).
 */
void ref_gravity(REF *r, int nReps, int bPeriodic, int iOrder, int bEwald, int iEwOrder, double dEwCut,
                 double dEwhCut, double *acc3, double *pot, double *dtGrav, double *fWeight, int *counts,
                 double *stats) {
    struct inGravity in;
    struct outGravity out;
    int i, iDum;
    double t0;
    memset(&in, 0, sizeof(in));
    in.nReps = nReps;
    in.bPeriodic = bPeriodic;
    in.iOrder = iOrder;
    in.bEwald = bEwald;
    in.iEwOrder = iEwOrder;
    in.bComove = 0;
    in.bDoSun = 0;
    in.dSunSoft = 0.0;
    in.dEwCut = dEwCut;
    in.dEwhCut = dEwhCut;
    in.dRhoFac = 0.0;
    pkdInitAccel(r->pkd);
    g_counts = counts;
    g_nCounts = r->pkd->nNodes;
    if (counts)
        for (i = 0; i < 3 * g_nCounts; ++i) counts[i] = -1;
    t0 = wallclock();
    pstGravity(r->pst, &in, sizeof(in), &out, &iDum);
    stats[5] = wallclock() - t0;
    g_counts = NULL;
    stats[0] = out.nActive;
    stats[1] = out.dPartSum;
    stats[2] = out.dCellSum;
    stats[3] = out.dSoftSum;
    stats[4] = out.dFlop;
    stats[6] = stats[7] = 0.0;
    for (i = 0; i < r->n; ++i) {
        PARTICLE *p = &r->pkd->pStore[i];
        acc3[3 * (size_t)i + 0] = p->a[0];
        acc3[3 * (size_t)i + 1] = p->a[1];
        acc3[3 * (size_t)i + 2] = p->a[2];
        pot[i] = p->fPot;
        dtGrav[i] = p->dtGrav;
        fWeight[i] = p->fWeight;
    }
}

/*
 * One bucket's lists, straight from pkdBucketWalk (walk.c:306) with the active-bbox swap pkdGravAll
 * performs around it (pkd.c:2916-2944).  ilp: 5 doubles (m,h,x,y,z); ilcs: 11; ilcn: 35.
 */
void ref_bucket_lists(REF *r, int iBucket, int nReps, int iOrder, int *n3, double *ilp, int nMaxP,
                      double *ilcs, int nMaxS, double *ilcn, int nMaxN) {
    PKD pkd = r->pkd;
    KDN *c = pkd->kdNodes;
    BND bndActive, bndTmp;
    int i, j;
    for (j = 0; j < 3; ++j) {
        bndActive.fMin[j] = FLOAT_MAXVAL;
        bndActive.fMax[j] = -FLOAT_MAXVAL;
    }
    for (i = c[iBucket].pLower; i <= c[iBucket].pUpper; ++i) {
        if (!TYPEQueryACTIVE(&pkd->pStore[i])) continue;
        for (j = 0; j < 3; ++j) {
            if (pkd->pStore[i].r[j] < bndActive.fMin[j]) bndActive.fMin[j] = pkd->pStore[i].r[j];
            if (pkd->pStore[i].r[j] > bndActive.fMax[j]) bndActive.fMax[j] = pkd->pStore[i].r[j];
        }
    }
    bndTmp = c[iBucket].bnd;
    c[iBucket].bnd = bndActive;
    __real_pkdBucketWalk(pkd, iBucket, nReps, iOrder);
    c[iBucket].bnd = bndTmp;
    n3[0] = pkd->nPart;
    n3[1] = pkd->nCellSoft;
    n3[2] = pkd->nCellNewt;
    if (ilp) memcpy(ilp, pkd->ilp, sizeof(ILP) * (size_t)(pkd->nPart < nMaxP ? pkd->nPart : nMaxP));
    if (ilcs) memcpy(ilcs, pkd->ilcs, sizeof(ILCS) * (size_t)(pkd->nCellSoft < nMaxS ? pkd->nCellSoft : nMaxS));
    if (ilcn) memcpy(ilcn, pkd->ilcn, sizeof(ILCN) * (size_t)(pkd->nCellNewt < nMaxN ? pkd->nCellNewt : nMaxN));
}

int ref_sizeof(int which) {
    switch (which) {
    case 0: return (int)sizeof(PARTICLE);
    case 1: return (int)sizeof(KDN);
    case 2: return (int)sizeof(ILP);
    case 3: return (int)sizeof(ILCS);
    case 4: return (int)sizeof(ILCN);
    case 5: return (int)sizeof(EWT);
    }
    return -1;
}

/*
 * The steps either side of the force evaluation, run by the reference's OWN pkdKick
 * (pkd.c:3780, -DNBODY branch :3956-3962), pkdDrift (pkd.c:3686-3777) and pkdGravStep (pkd.c:4609-4623) on a
 * throw-away one-rank PKD filled from the given arrays.  what: bit 0 kick, bit 1 drift, bit 2 grav-step; applied in
 * that order.  r3, v3, dt are updated in place.  Pins oracle_step_ops (oracle/gravity_oracle.c) and the golden
 * fixture tests/golden/stepops.npz.
 */
void ref_step_ops(int n, double *r3, double *v3, const double *a3, const int *active, const double *dtGrav, double *dt,
                  double dvFacOne, double dvFacTwo, double dDelta, const double *fCenter, int bPeriodic,
                  const double *fPeriod, double dEta, int what) {
    static const double open3[3] = {FLOAT_MAXVAL, FLOAT_MAXVAL, FLOAT_MAXVAL};
    double *zero = calloc((size_t)n + 1, sizeof(double));
    REF *r = ref_create(n, zero, zero, zero, zero, zero, NULL, bPeriodic ? fPeriod : open3);
    FLOAT c[3];
    UHC uhc;
    int i, j;
    free(zero);
    memset(&uhc, 0, sizeof(uhc));
    for (i = 0; i < n; ++i) {
        PARTICLE *p = &r->pkd->pStore[i];
        p->iActive = TYPE_DARK | TYPE_TREEACTIVE | ((!active || active[i]) ? TYPE_ACTIVE : 0);
        for (j = 0; j < 3; ++j) {
            p->r[j] = r3[3 * i + j];
            p->v[j] = v3[3 * i + j];
            p->a[j] = a3[3 * i + j];
        }
        p->dtGrav = dtGrav[i];
        p->dt = dt[i];
    }
    for (j = 0; j < 3; ++j) c[j] = fCenter[j];
    if (what & 1) pkdKick(r->pkd, dvFacOne, dvFacTwo, 0.0, 0.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0, uhc);
    if (what & 2) pkdDrift(r->pkd, dDelta, c, bPeriodic, 0, 0, 0.0, 0.0);
    if (what & 4) pkdGravStep(r->pkd, dEta);
    for (i = 0; i < n; ++i) {
        PARTICLE *p = &r->pkd->pStore[i];
        for (j = 0; j < 3; ++j) {
            r3[3 * i + j] = p->r[j];
            v3[3 * i + j] = p->v[j];
        }
        dt[i] = p->dt;
    }
    ref_destroy(r);
}

/*
 * Time-step selection on the reference's own code: pkdInitDt (pkd.c:4818), pkdAccelStep
 * (pkd.c:4625, -DNBODY: the gas branch is not compiled), pkdGravStep (pkd.c:4609), pkdDtToRung (pkd.c:4715) and
 * pkdActiveRung (pkd.c:4569), in that order as selected by `what` (bits 0..4), on a throw-away one-rank PKD.
 * dt, rung, active are updated in place; out[0] = pkdDtToRung's return (iMaxRungOut), out[1] = nMaxRung,
 * out[2] = iMaxRungIdeal, out[3] = pkdActiveRung's return.
 */
void ref_rung_ops(int n, const double *v3, const double *a3, const double *fPot, const double *fSoft,
                  const double *dtGrav, int *active, double *dt, int *rung, double dDelta, double dEta, double dVelFac,
                  double dAccFac, int bEpsAcc, int bSqrtPhi, int iRung, int iMaxRung, int bAll, int iRungActive,
                  int bGreater, int what, int *out) {
    static const double open3[3] = {FLOAT_MAXVAL, FLOAT_MAXVAL, FLOAT_MAXVAL};
    double *zero = calloc((size_t)n + 1, sizeof(double));
    REF *r = ref_create(n, zero, zero, zero, zero, zero, NULL, open3);
    int i, j, nMaxRung = 0, iMaxRungIdeal = 0;
    free(zero);
    out[0] = out[1] = out[2] = out[3] = 0;
    for (i = 0; i < n; ++i) {
        PARTICLE *p = &r->pkd->pStore[i];
        p->iActive = TYPE_DARK | TYPE_TREEACTIVE | (active[i] ? TYPE_ACTIVE : 0);
        for (j = 0; j < 3; ++j) {
            p->v[j] = v3[3 * i + j];
            p->a[j] = a3[3 * i + j];
        }
        p->fPot = fPot[i];
        p->fSoft = fSoft[i];
        p->dtGrav = dtGrav[i];
        p->dt = dt[i];
        p->iRung = rung[i];
    }
    if (what & 1) pkdInitDt(r->pkd, dDelta);
    if (what & 2) pkdAccelStep(r->pkd, dEta, dVelFac, dAccFac, 1, bEpsAcc, bSqrtPhi, 0.0);
    if (what & 4) pkdGravStep(r->pkd, dEta);
    if (what & 8) {
        out[0] = pkdDtToRung(r->pkd, iRung, dDelta, iMaxRung, bAll, 0, &nMaxRung, &iMaxRungIdeal);
        out[1] = nMaxRung;
        out[2] = iMaxRungIdeal;
    }
    if (what & 16) out[3] = pkdActiveRung(r->pkd, iRungActive, bGreater);
    for (i = 0; i < n; ++i) {
        PARTICLE *p = &r->pkd->pStore[i];
        dt[i] = p->dt;
        rung[i] = p->iRung;
        active[i] = TYPEQueryACTIVE(p) ? 1 : 0;
    }
    ref_destroy(r);
}

/*
 * pstGravity with bDoSun = 1 (pkd.c:3003-3041: the indirect acceleration of solar-system runs -- a dummy sink at the
 * origin with softening dSunSoft walks the tree and is evaluated like any bucket; open boundaries only): returns aSun
 * and the dummy bucket's list counts.  The particles' own results are untouched by the dummy pass.
 */
void ref_gravity_sun(REF *r, int iOrder, double dSunSoft, double *aSun3, int *sunCounts3) {
    struct inGravity in;
    struct outGravity out;
    int iDum, j;
    memset(&in, 0, sizeof(in));
    in.nReps = 0;
    in.bPeriodic = 0;
    in.iOrder = iOrder;
    in.bEwald = 0;
    in.iEwOrder = iOrder;
    in.bDoSun = 1;
    in.dSunSoft = dSunSoft;
    in.dEwCut = 2.6;
    in.dEwhCut = 2.8;
    pkdInitAccel(r->pkd);
    g_sunCounts = sunCounts3;
    pstGravity(r->pst, &in, sizeof(in), &out, &iDum);
    g_sunCounts = NULL;
    for (j = 0; j < 3; ++j) aSun3[j] = out.aSun[j];
}

/*
 * pstGravity with bComove = 1 on open boundaries (pkd.c:2967-2991: acceleration and potential of the uniform negative
 * background, a += dRhoFac r, fPot -= dRhoFac r^2 / 2 on ACTIVE particles, added per bucket right after its interactions).
 */
void ref_gravity_comove(REF *r, int iOrder, double dRhoFac, double *acc3, double *pot) {
    struct inGravity in;
    struct outGravity out;
    int i, iDum;
    memset(&in, 0, sizeof(in));
    in.iOrder = iOrder;
    in.iEwOrder = iOrder;
    in.bComove = 1;
    in.dRhoFac = dRhoFac;
    in.dEwCut = 2.6;
    in.dEwhCut = 2.8;
    pkdInitAccel(r->pkd);
    pstGravity(r->pst, &in, sizeof(in), &out, &iDum);
    for (i = 0; i < r->n; ++i) {
        PARTICLE *p = &r->pkd->pStore[i];
        acc3[3 * (size_t)i + 0] = p->a[0];
        acc3[3 * (size_t)i + 1] = p->a[1];
        acc3[3 * (size_t)i + 2] = p->a[2];
        pot[i] = p->fPot;
    }
}
