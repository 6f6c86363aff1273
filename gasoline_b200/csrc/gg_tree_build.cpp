// gg_tree_build.cpp -- host-side k-d tree construction with the semantics of the reference's gravity tree.
//
// Hosts that already own a Gasoline tree (the pkdGravAll shim) never call this; bench.py and the tests do, because
// the force path needs a tree with exactly the reference's geometry to produce the reference's interaction lists:
//   * spatial bisection: split the longest axis of the SQUEEZED bounding box at its midpoint, first axis wins
//     ties; a cell becomes a bucket when it holds <= nBucket particles or has zero extent (BuildBinary,
//     pkd.c:2437-2587); the particle exchange order of the partition is the reference's (pkdUpperPart,
//     pkd.c:1106-1133), which fixes the summation order of every centre of mass;
//   * cells are numbered in depth-first pre-order (cell, lower subtree, upper subtree), as pkd->iFreeCell++ does;
//   * mass, centre of mass and mass-weighted softening come from the children (buckets: from the particles);
//     reduced multipoles to hexadecapole and Bmax are summed particle by particle about the cell's centre
//     (pkdCalcCell, pkd.c:2018-2135); fOpen2 = max(Bmax, 2/sqrt(3) Bmax/theta)^2 (OPEN_JOSH, pkd.c:2253-2260), or one of
//     the other opening criteria of pkdCalcOpen (pkd.c:2228-2264) through gg_tree_build_open;
//   * links are threaded: iLower = first child, iUpper = next cell (pkdThreadTree, pkd.c:2590-2620);
//   * the Ewald root expansion holds COMPLETE l=3,4 moments about the root centre (pkdCalcRoot, pkd.c:4395).
// Unlike the reference's single recursion this builder partitions the top levels serially, hands the subtrees to
// worker threads, and stitches the pre-order numbering afterwards; results do not depend on the thread count.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include "../../include/gasoline_b200.h"

namespace {

struct P {
    double r[3], m, h;
    int active, iOrder;
};

struct Cell { // one tree cell during construction; child links are indices into the owning vector
    int lo, hi, left, right, dim;
};

struct gg_built_tree_impl {
    int nNodes = 0, iRoot = -1;
    std::vector<double> bnd, r, fMass, fSoft, fOpen2, mom, bmom; // bmom: Bmax, B2..B6 per cell (pkd.h:441-451)
    std::vector<int> pLower, pUpper, iLower, iUpper;
    double root[GG_NROOT];
};

inline void bounds(const P *p, int lo, int hi, double *b) {
    for (int j = 0; j < 3; ++j) b[j] = b[3 + j] = p[lo].r[j];
    for (int i = lo + 1; i <= hi; ++i)
        for (int j = 0; j < 3; ++j) {
            const double v = p[i].r[j];
            if (v < b[j]) b[j] = v;
            else if (v > b[3 + j]) b[3 + j] = v;
        }
}

// Exchange the leftmost element >= split with the rightmost < split until the pointers cross.
inline int partition_upper(P *p, int d, double split, int lo, int hi) {
    for (;;) {
        while (lo <= hi && p[lo].r[d] < split) ++lo;
        while (lo <= hi && p[hi].r[d] >= split) --hi;
        if (lo >= hi) return lo;
        std::swap(p[lo], p[hi]);
        ++lo;
        --hi;
    }
}

// Decide whether [lo,hi] splits; if so partition it and return the first index of the upper part, else -1.
inline int split_cell(P *p, int lo, int hi, int nBucket, int *dim) {
    double b[6];
    bounds(p, lo, hi, b);
    bool good = false;
    for (int j = 0; j < 3; ++j) good = good || (b[3 + j] > b[j]);
    *dim = -1;
    if (!(hi - lo + 1 > nBucket && good)) return -1;
    int d = 0;
    for (int j = 1; j < 3; ++j)
        if (b[3 + j] - b[j] > b[3 + d] - b[d]) d = j;
    *dim = d;
    return partition_upper(p, d, 0.5 * (b[d] + b[3 + d]), lo, hi);
}

// Build the shape (ranges + child links, pre-order) of the subtree over [lo,hi] into `cells`.
void build_shape(P *p, int lo, int hi, int nBucket, std::vector<Cell> &cells) {
    struct Frame { int cell, stage; };
    std::vector<Frame> st;
    cells.push_back({lo, hi, -1, -1, -1});
    st.push_back({(int)cells.size() - 1, 0});
    while (!st.empty()) {
        Frame &f = st.back();
        Cell c = cells[f.cell];
        if (f.stage == 0) {
            int dim, m = split_cell(p, c.lo, c.hi, nBucket, &dim);
            cells[f.cell].dim = dim;
            if (m < 0) { st.pop_back(); continue; }
            cells[f.cell].right = m; // remember the split point until the upper child exists
            f.stage = 1;
            cells.push_back({c.lo, m - 1, -1, -1, -1});
            cells[f.cell].left = (int)cells.size() - 1;
            st.push_back({(int)cells.size() - 1, 0});
        } else if (f.stage == 1) {
            int m = c.right;
            f.stage = 2;
            cells.push_back({m, c.hi, -1, -1, -1});
            cells[f.cell].right = (int)cells.size() - 1;
            st.push_back({(int)cells.size() - 1, 0});
        } else st.pop_back();
    }
}

// bnum (may be null): Bmax and the radial moments B2..B6 = sum m d^k the opening criteria are made of (pkd.c:2080-2087)
void cell_moments(const P *p, int lo, int hi, const double *rc, int iOrder, double *q, double *bmax, double *bnum = nullptr) {
    double B = 0.0, b2 = 0.0, b3 = 0.0, b4 = 0.0, b5 = 0.0, b6 = 0.0;
    for (int k = 0; k < GG_NMOM; ++k) q[k] = 0.0;
    for (int j = lo; j <= hi; ++j) {
        const double m = p[j].m, dx = p[j].r[0] - rc[0], dy = p[j].r[1] - rc[1], dz = p[j].r[2] - rc[2];
        const double d2 = dx * dx + dy * dy + dz * dz, d1 = std::sqrt(d2);
        if (d1 > B) B = d1;
        if (bnum) {
            b2 += m * d2; b3 += m * d2 * d1; b4 += m * d2 * d2; b5 += m * d2 * d2 * d1; b6 += m * d2 * d2 * d2;
        }
        if (iOrder >= 4) {
            q[16] += m * (dx * dx * dx * dx - 6.0 / 7.0 * d2 * (dx * dx - 0.1 * d2));
            q[17] += m * (dx * dy * dy * dy - 3.0 / 7.0 * d2 * dx * dy);
            q[18] += m * (dx * dx * dx * dy - 3.0 / 7.0 * d2 * dx * dy);
            q[19] += m * (dy * dy * dy * dy - 6.0 / 7.0 * d2 * (dy * dy - 0.1 * d2));
            q[20] += m * (dx * dx * dx * dz - 3.0 / 7.0 * d2 * dx * dz);
            q[21] += m * (dy * dy * dy * dz - 3.0 / 7.0 * d2 * dy * dz);
            q[22] += m * (dx * dx * dy * dy - 1.0 / 7.0 * d2 * (dx * dx + dy * dy - 0.2 * d2));
            q[23] += m * (dx * dx * dy * dz - 1.0 / 7.0 * d2 * dy * dz);
            q[24] += m * (dx * dy * dy * dz - 1.0 / 7.0 * d2 * dx * dz);
            q[25] += m * (dx * dx * dz * dz - 1.0 / 7.0 * d2 * (dx * dx + dz * dz - 0.2 * d2));
            q[26] += m * (dx * dy * dz * dz - 1.0 / 7.0 * d2 * dx * dy);
            q[27] += m * (dx * dz * dz * dz - 3.0 / 7.0 * d2 * dx * dz);
            q[28] += m * (dy * dy * dz * dz - 1.0 / 7.0 * d2 * (dy * dy + dz * dz - 0.2 * d2));
            q[29] += m * (dy * dz * dz * dz - 3.0 / 7.0 * d2 * dy * dz);
            q[30] += m * (dz * dz * dz * dz - 6.0 / 7.0 * d2 * (dz * dz - 0.1 * d2));
        }
        if (iOrder >= 3) {
            q[6] += m * (dx * dx * dx - 0.6 * d2 * dx);
            q[7] += m * (dx * dy * dy - 0.2 * d2 * dx);
            q[8] += m * (dx * dx * dy - 0.2 * d2 * dy);
            q[9] += m * (dy * dy * dy - 0.6 * d2 * dy);
            q[10] += m * (dx * dx * dz - 0.2 * d2 * dz);
            q[11] += m * (dy * dy * dz - 0.2 * d2 * dz);
            q[12] += m * dx * dy * dz;
            q[13] += m * (dx * dz * dz - 0.2 * d2 * dx);
            q[14] += m * (dy * dz * dz - 0.2 * d2 * dy);
            q[15] += m * (dz * dz * dz - 0.6 * d2 * dz);
        }
        q[0] += m * dx * dx; q[1] += m * dy * dy; q[2] += m * dz * dz;
        q[3] += m * dx * dy; q[4] += m * dx * dz; q[5] += m * dy * dz;
    }
    *bmax = B;
    if (bnum) { bnum[0] = B; bnum[1] = b2; bnum[2] = b3; bnum[3] = b4; bnum[4] = b5; bnum[5] = b6; }
}

// The error estimate of a cell's expansion truncated after `order`, seen from distance r (fcnAbsMono .. fcnAbsHex,
// pkd.c:2137-2179): ((l + 2) B_{l+1} - (l + 1) B_{l+2} / r) / (r^l (r - Bmax))^2 ... with the reference's grouping.
inline double abs_error(const double *b, int order, double r) {
    double t;
    switch (order) {
    case 1: t = r * (r - b[0]); t *= t; return (3.0 * b[1] - 2.0 * b[2] / r) / t;
    case 2: t = r * (r - b[0]); t *= r * t; return (4.0 * b[2] - 3.0 * b[3] / r) / t;
    case 3: t = r * r * (r - b[0]); t *= t; return (5.0 * b[3] - 4.0 * b[4] / r) / t;
    default: t = r * r * (r - b[0]); t *= r * t; return (6.0 * b[4] - 5.0 * b[5] / r) / t;
    }
}

// OPEN_ABSPAR: the distance at which that estimate falls to dErrBnd, by the reference's hunt + bisection
// (dRootBracket, pkd.c:2182-2224): start just outside Bmax, double until the estimate is below the bound, halve the
// bracket until the estimate is within 1e-6 dErrBnd below it (or 33 halvings).  A cell without extent (one particle,
// coincident particles) has Bmax = B_k = 0: the estimate is 0/0 there and the reference's hunt never ends -- reported
// through *bad instead of repeated.
double open_abs_partial(const double *b, double dErrBnd, int order, bool *bad) {
    const double crit = 1e-6 * dErrBnd;
    double lower = (1.0 + 1e-6) * b[0];
    const double e0 = abs_error(b, order, lower);
    if (e0 < dErrBnd) return lower;
    if (!(e0 == e0) || !(lower > 0.0)) { *bad = true; return lower; }
    double upper = 2 * lower, mid;
    for (;;) {
        const double dif = dErrBnd - abs_error(b, order, upper);
        if (dif > crit) break;
        if (!(dif == dif) || upper > 1.0e300) { *bad = true; return upper; }
        upper = 2 * upper;
    }
    for (int iter = 0;;) {
        mid = 0.5 * (lower + upper);
        const double dif = dErrBnd - abs_error(b, order, mid);
        if (dif < 0) lower = mid;
        else {
            upper = mid;
            if (dif < crit) break;
        }
        if (++iter > 32) break;
    }
    return mid;
}

// pkdCalcOpen (pkd.c:2228-2264)
inline double open_radius(const double *b, int iOpenType, double dCrit, int order, bool *bad) {
    if (iOpenType == GG_OPEN_ABSPAR) return open_abs_partial(b, dCrit, order, bad);
    if (iOpenType == GG_OPEN_JOSH) {
        const double dOpen = 2 / std::sqrt(3.0) * b[0] / dCrit;
        return dOpen < b[0] ? b[0] : dOpen;
    }
    return b[0]; // OPEN_RELPAR, OPEN_ABSTOT, OPEN_RELTOT: "the minimal, i.e., Bmax"
}

} // namespace

struct gg_built_tree : gg_built_tree_impl {};

extern "C" int gg_tree_build(int n, double *x, double *y, double *z, double *fMass, double *fSoft, int *active,
                             int *iOrderOut, int nBucket, double dTheta, int iOrderMom, int nThreads,
                             gg_built_tree **out) {
    return gg_tree_build_open(n, x, y, z, fMass, fSoft, active, iOrderOut, nBucket, GG_OPEN_JOSH, dTheta, iOrderMom,
                              nThreads, out);
}

extern "C" int gg_tree_build_open(int n, double *x, double *y, double *z, double *fMass, double *fSoft, int *active,
                                  int *iOrderOut, int nBucket, int iOpenType, double dCrit, int iOrderMom, int nThreads,
                                  gg_built_tree **out) {
    if (n <= 0 || !x || !y || !z || !fMass || !fSoft || !out || nBucket < 1 || !(dCrit > 0)) return GG_ERR_ARG;
    if (iOpenType < GG_OPEN_JOSH || iOpenType > GG_OPEN_RELTOT || iOrderMom < 1 || iOrderMom > 4) return GG_ERR_ARG;
    if (nThreads <= 0) nThreads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<P> ps((size_t)n);
    for (int i = 0; i < n; ++i) {
        ps[i].r[0] = x[i]; ps[i].r[1] = y[i]; ps[i].r[2] = z[i];
        ps[i].m = fMass[i]; ps[i].h = fSoft[i];
        ps[i].active = active ? active[i] : 1;
        ps[i].iOrder = i;
    }
    P *p = ps.data();

    // ---- phase 1: shape of the top of the tree, serially, until there is enough independent work
    std::vector<Cell> top;
    top.push_back({0, n - 1, -1, -1, -1});
    std::vector<int> frontier{0}, pending; // pending: top cells whose subtree goes to a worker (dim == -2)
    const size_t want = nThreads > 1 ? (size_t)nThreads * 8 : 0;
    while (!frontier.empty() && frontier.size() + pending.size() < want) {
        std::vector<int> next;
        for (int c : frontier) {
            int dim, m = split_cell(p, top[c].lo, top[c].hi, nBucket, &dim);
            top[c].dim = dim;
            if (m < 0) continue; // bucket
            const int lo = top[c].lo, hi = top[c].hi;
            top.push_back({lo, m - 1, -1, -1, -1});
            top.push_back({m, hi, -1, -1, -1});
            top[c].left = (int)top.size() - 2;
            top[c].right = (int)top.size() - 1;
            next.push_back(top[c].left);
            next.push_back(top[c].right);
        }
        frontier.swap(next);
    }
    for (int c : frontier) { top[c].dim = -2; pending.push_back(c); }

    // ---- phase 2: subtrees in parallel (each worker partitions a disjoint particle range)
    std::vector<std::vector<Cell>> sub(pending.size());
    {
        std::atomic<size_t> nextJob{0};
        auto work = [&]() {
            for (;;) {
                size_t j = nextJob.fetch_add(1);
                if (j >= pending.size()) break;
                build_shape(p, top[pending[j]].lo, top[pending[j]].hi, nBucket, sub[j]);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nThreads && (size_t)t < pending.size(); ++t) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }

    // ---- phase 3: pre-order numbering across top + subtrees
    gg_built_tree *bt = new gg_built_tree();
    std::vector<int> pendingOf(top.size(), -1);
    for (size_t j = 0; j < pending.size(); ++j) pendingOf[pending[j]] = (int)j;
    size_t total = 0;
    for (size_t c = 0; c < top.size(); ++c) total += (top[c].dim == -2) ? sub[pendingOf[c]].size() : 1;
    const int nn = (int)total;
    bt->nNodes = nn;
    bt->iRoot = 0;
    bt->bnd.resize((size_t)nn * 6); bt->r.resize((size_t)nn * 3); bt->fMass.resize(nn); bt->fSoft.resize(nn);
    bt->fOpen2.resize(nn); bt->mom.resize((size_t)nn * GG_NMOM); bt->bmom.resize((size_t)nn * 6);
    bt->pLower.resize(nn); bt->pUpper.resize(nn); bt->iLower.resize(nn); bt->iUpper.resize(nn);
    std::vector<int> left(nn, -1), right(nn, -1);
    std::vector<int> topIndex(top.size(), -1), subBase(pending.size(), -1);
    {
        int next = 0;
        std::vector<int> st{0};
        while (!st.empty()) {
            int c = st.back();
            st.pop_back();
            if (top[c].dim == -2) {
                subBase[pendingOf[c]] = next;
                topIndex[c] = next;
                next += (int)sub[pendingOf[c]].size();
            } else {
                topIndex[c] = next++;
                if (top[c].left >= 0) { st.push_back(top[c].right); st.push_back(top[c].left); }
            }
        }
    }
    for (size_t c = 0; c < top.size(); ++c) {
        if (top[c].dim == -2) continue;
        int g = topIndex[c];
        bt->pLower[g] = top[c].lo; bt->pUpper[g] = top[c].hi;
        if (top[c].left >= 0) { left[g] = topIndex[top[c].left]; right[g] = topIndex[top[c].right]; }
    }
    for (size_t j = 0; j < pending.size(); ++j)
        for (size_t k = 0; k < sub[j].size(); ++k) {
            int g = subBase[j] + (int)k;
            bt->pLower[g] = sub[j][k].lo; bt->pUpper[g] = sub[j][k].hi;
            if (sub[j][k].left >= 0) { left[g] = subBase[j] + sub[j][k].left; right[g] = subBase[j] + sub[j][k].right; }
        }

    // ---- phase 4: bounds, mass/COM/softening (children before parents: descending pre-order index), then
    //      moments and opening radius of every cell in parallel.
    for (int g = nn - 1; g >= 0; --g) {
        const int lo = bt->pLower[g], hi = bt->pUpper[g];
        double *rc = &bt->r[3 * (size_t)g];
        double M = 0.0, S = 0.0;
        rc[0] = rc[1] = rc[2] = 0.0;
        if (left[g] >= 0) {
            const int kids[2] = {left[g], right[g]};
            for (int k = 0; k < 2; ++k) {
                const double fm = bt->fMass[kids[k]];
                M += fm;
                S += fm * bt->fSoft[kids[k]];
                for (int j = 0; j < 3; ++j) rc[j] += fm * bt->r[3 * (size_t)kids[k] + j];
            }
            double *b = &bt->bnd[6 * (size_t)g];
            const double *bl = &bt->bnd[6 * (size_t)left[g]], *br = &bt->bnd[6 * (size_t)right[g]];
            for (int j = 0; j < 3; ++j) { // the squeezed box of a cell is the union of its children's
                b[j] = std::min(bl[j], br[j]);
                b[3 + j] = std::max(bl[3 + j], br[3 + j]);
            }
        } else {
            bounds(p, lo, hi, &bt->bnd[6 * (size_t)g]);
            for (int i = lo; i <= hi; ++i) {
                const double fm = p[i].m;
                M += fm;
                S += fm * p[i].h;
                for (int j = 0; j < 3; ++j) rc[j] += fm * p[i].r[j];
            }
        }
        if (M > 0) {
            S /= M;
            for (int j = 0; j < 3; ++j) rc[j] /= M;
        }
        bt->fMass[g] = M;
        bt->fSoft[g] = S;
    }
    std::atomic<bool> badOpen{false};
    {
        std::atomic<int> nextCell{0};
        auto work = [&]() {
            for (;;) {
                int g0 = nextCell.fetch_add(256);
                if (g0 >= nn) break;
                for (int g = g0; g < std::min(nn, g0 + 256); ++g) {
                    double bmax, *b = &bt->bmom[6 * (size_t)g];
                    bool bad = false;
                    cell_moments(p, bt->pLower[g], bt->pUpper[g], &bt->r[3 * (size_t)g], iOrderMom,
                                 &bt->mom[(size_t)GG_NMOM * g], &bmax, b);
                    const double dOpen = open_radius(b, iOpenType, dCrit, iOrderMom, &bad);
                    if (bad) badOpen = true;
                    bt->fOpen2[g] = dOpen * dOpen;
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nThreads; ++t) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
    }
    if (badOpen) { // OPEN_ABSPAR on a tree with a zero-extent cell: the reference does not return from it either
        delete bt;
        return GG_ERR_UNSUPPORTED;
    }
    // ---- threading: next[g] = sibling if g is a lower child, else the parent's next
    {
        std::vector<int> st{0};
        bt->iUpper[0] = -1;
        while (!st.empty()) {
            int g = st.back();
            st.pop_back();
            bt->iLower[g] = left[g];
            if (left[g] >= 0) {
                bt->iUpper[left[g]] = right[g];
                bt->iUpper[right[g]] = bt->iUpper[g];
                st.push_back(right[g]);
                st.push_back(left[g]);
            }
        }
    }
    // ---- Ewald root expansion
    {
        double *R = bt->root;
        const double *rc = &bt->r[0], *q = &bt->mom[0];
        for (int k = 0; k < GG_NROOT; ++k) R[k] = 0.0;
        for (int i = 0; i < n; ++i) {
            const double m = p[i].m, dx = p[i].r[0] - rc[0], dy = p[i].r[1] - rc[1], dz = p[i].r[2] - rc[2];
            R[20] += m * (dx * dx * dx * dx); R[21] += m * (dx * dy * dy * dy); R[22] += m * (dx * dx * dx * dy);
            R[23] += m * (dy * dy * dy * dy); R[24] += m * (dx * dx * dx * dz); R[25] += m * (dy * dy * dy * dz);
            R[26] += m * (dx * dx * dy * dy); R[27] += m * (dx * dx * dy * dz); R[28] += m * (dx * dy * dy * dz);
            R[29] += m * (dx * dx * dz * dz); R[30] += m * (dx * dy * dz * dz); R[31] += m * (dx * dz * dz * dz);
            R[32] += m * (dy * dy * dz * dz); R[33] += m * (dy * dz * dz * dz); R[34] += m * (dz * dz * dz * dz);
            R[10] += m * (dx * dx * dx); R[11] += m * (dx * dy * dy); R[12] += m * (dx * dx * dy);
            R[13] += m * (dy * dy * dy); R[14] += m * (dx * dx * dz); R[15] += m * (dy * dy * dz);
            R[16] += m * dx * dy * dz; R[17] += m * (dx * dz * dz); R[18] += m * (dy * dz * dz);
            R[19] += m * (dz * dz * dz);
        }
        R[0] = bt->fMass[0]; R[1] = rc[0]; R[2] = rc[1]; R[3] = rc[2];
        R[4] = q[0]; R[5] = q[1]; R[6] = q[3]; R[7] = q[4]; R[8] = q[5]; R[9] = q[2];
    }
    for (int i = 0; i < n; ++i) {
        x[i] = p[i].r[0]; y[i] = p[i].r[1]; z[i] = p[i].r[2];
        fMass[i] = p[i].m; fSoft[i] = p[i].h;
        if (active) active[i] = p[i].active;
        if (iOrderOut) iOrderOut[i] = p[i].iOrder;
    }
    *out = bt;
    return GG_OK;
}

extern "C" int gg_tree_view(const gg_built_tree *bt, gg_tree *v, double root[GG_NROOT]) {
    if (!bt || !v) return GG_ERR_ARG;
    v->nNodes = bt->nNodes; v->iRoot = bt->iRoot;
    v->bnd = bt->bnd.data(); v->r = bt->r.data(); v->fMass = bt->fMass.data(); v->fSoft = bt->fSoft.data();
    v->fOpen2 = bt->fOpen2.data(); v->mom = bt->mom.data();
    v->pLower = bt->pLower.data(); v->pUpper = bt->pUpper.data(); v->iLower = bt->iLower.data();
    v->iUpper = bt->iUpper.data();
    if (root) std::memcpy(root, bt->root, sizeof(bt->root));
    return GG_OK;
}

extern "C" int gg_tree_bnumbers(const gg_built_tree *bt, double *bmom) {
    if (!bt || !bmom) return GG_ERR_ARG;
    std::memcpy(bmom, bt->bmom.data(), bt->bmom.size() * sizeof(double));
    return GG_OK;
}

extern "C" void gg_tree_free(gg_built_tree *bt) { delete bt; }

// pkdCalcCell (pkd.c:2018) over a whole domain about a given centre: one rank's contribution to a top-tree cell.
extern "C" int gg_cell_moments(int n, const double *x, const double *y, const double *z, const double *fMass,
                               const double rcm[3], int iOrder, double mom[GG_NMOM], double *pBmax) {
    if (n < 0 || !x || !y || !z || !fMass || !rcm || !mom || !pBmax || iOrder < 1 || iOrder > 4) return GG_ERR_ARG;
    std::vector<P> p((size_t)n);
    for (int i = 0; i < n; ++i) {
        p[i].r[0] = x[i]; p[i].r[1] = y[i]; p[i].r[2] = z[i];
        p[i].m = fMass[i]; p[i].h = 0.0; p[i].active = 1; p[i].iOrder = i;
    }
    cell_moments(p.data(), 0, n - 1, rcm, iOrder, mom, pBmax);
    return GG_OK;
}

// The moments of every cell by the DEVICE's algorithm (gg_moments.cu: raw moments of the buckets, children translated
// and summed in (lower, upper) order, gg_raw_reduce), run on the host: lets the CPU test suite check the translation
// and reduction formulas of gg_m2m.h against the particle-by-particle sums of pkdCalcCell without a GPU.
#include "gg_m2m.h"
extern "C" int gg_tree_moments_m2m(const gg_tree *t, const gg_particles *pp, double *mom) {
    if (!t || !pp || !mom || t->nNodes < 1) return GG_ERR_ARG;
    const int nn = t->nNodes;
    std::vector<GGRawMom> raw((size_t)nn);
    std::vector<char> done((size_t)nn, 0);
    std::vector<int> st{t->iRoot};
    while (!st.empty()) {
        const int g = st.back();
        const int c0 = t->iLower[g];
        if (c0 < 0) {
            GGRawMom a;
            gg_raw_zero(a);
            for (int i = t->pLower[g]; i <= t->pUpper[g]; ++i)
                gg_raw_add_particle(a, pp->fMass[i], pp->x[i] - t->r[3 * (size_t)g], pp->y[i] - t->r[3 * (size_t)g + 1],
                                    pp->z[i] - t->r[3 * (size_t)g + 2]);
            raw[g] = a;
        } else {
            const int c1 = t->iUpper[c0];
            if (!done[c0]) { st.push_back(c1); st.push_back(c0); continue; }
            GGRawMom a;
            gg_raw_zero(a);
            const int kids[2] = {c0, c1};
            for (int k = 0; k < 2; ++k)
                gg_raw_shift_add(a, raw[kids[k]], t->r[3 * (size_t)kids[k]] - t->r[3 * (size_t)g],
                                 t->r[3 * (size_t)kids[k] + 1] - t->r[3 * (size_t)g + 1],
                                 t->r[3 * (size_t)kids[k] + 2] - t->r[3 * (size_t)g + 2]);
            raw[g] = a;
        }
        gg_raw_reduce(raw[g], &mom[(size_t)GG_NMOM * g]);
        done[g] = 1;
        st.pop_back();
    }
    return GG_OK;
}
