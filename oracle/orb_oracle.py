"""CPU restatement (numpy) of the reference's ORB domain decomposition.  TEST INFRASTRUCTURE ONLY: tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product path (gasoline_b200/) must not.

PINNED by execution (tests/test_oracle_orb.py):
  * the first call of a run (pst->iSplitDim == -1, bDoRootFind = bDoSplitDimFind = 1, master.c:4176-4177) against the
    domains the compiled reference produced on 2, 3 and 4 pthread-MDL ranks (tests/golden/multirank_*.npz: r<k>_iOrder)
    and the reference binary run live on 5-8 ranks, particle for particle;
  * later calls (`prev` = the cells' split axis and split of the previous decomposition: the NEWSPLITDIMCUT hysteresis
    of pst.c:1900-1910, work weights) against a time-stepping multi-rank run (tests/golden/orbsteps_*.npz);
  * the store-overflow ("reverse" / wrap) split of pst.c:1049-1270 (`stores` given) against stepping runs whose work-weighted
    splits send a rank more particles than its pStore holds (tests/golden/orbsteps_*_overflow*.npz).
Restated from the source only: the bDoRootFind = 0 / bDoSplitDimFind = 0 shortcuts of master.c:4210-4222 (the host takes them
for small active sets of a multi-rung run).

Follows, per node of the rank tree:
  * pstDomainDecomp (pst.c:1854-1935): the node's bounds are the bounds of its particles (pstCalcBound); the split
    dimension is the first axis of strictly largest extent (iSplitDim == -1 on the first call, pst.c:1900-1910);
  * _pstRootSplit (pst.c:959-1034): bisection on fSplit between the bounds, at most MAX_ITTR = 64 steps, while
    fl < fmm < fu; a particle is "low" when r[d] < fSplit (pkdLowerPart/pkdUpperPart, pkd.c:1064-1133); the branch
    taken compares fLow/nLower with fHigh/nUpper (bSplitWork, the default, master.c:964: weights of the particles,
    fWeight = 1 after reading a file) or nLow/nLower with nHigh/nUpper; equal shares or nLow == nHigh == 1 stop it;
  * the "reverse" split (pst.c:1049-1270): a second boundary fSplitInactive; the lower set of ranks receives the particles
    of the WRAPPED interval between the two (pkdColRejects -> pkdLowerPartWrap, pkd.c:1165-1211, 1463-1485).  While both
    sides' stores have room fSplitInactive lies just outside the bounds and the interval is r[d] < fSplit; when one side's
    particles would not fit its ranks' stores, fSplitInactive is bisected into the cell until they do.
"""
from __future__ import annotations

import numpy as np

MAX_ITTR = 64  # pst.c:874
NEWSPLITDIMCUT = 0.707  # pst.c:1851


def split_dim(lo, hi, prev_dim: int = -1) -> int:
    """pst.c:1900-1910: the axis loop with the previous axis's extent x NEWSPLITDIMCUT as the bar to beat."""
    d = prev_dim
    dimsize = -1.0 if d == -1 else (hi[d] - lo[d]) * NEWSPLITDIMCUT
    for j in range(3):
        if hi[j] - lo[j] > dimsize:
            d, dimsize = j, hi[j] - lo[j]
    return d


def root_split(c: np.ndarray, w: np.ndarray | None, nLower: int, nUpper: int, fl: float, fu: float):
    """The root finder of _pstRootSplit on the coordinates c of one node.  Returns (fSplit, iterations)."""
    fm = np.nan
    fmm = (fl + fu) / 2
    ittr = 0
    while fl < fmm and fmm < fu and ittr < MAX_ITTR:
        fm = fmm
        low = c < fm
        nLow = int(np.count_nonzero(low))
        nHigh = c.size - nLow
        if nLow == 1 and nHigh == 1:
            break
        if w is not None:
            a, b = float(w[low].sum()) / nLower, float(w[~low].sum()) / nUpper
        else:
            a, b = nLow / float(nLower), nHigh / float(nUpper)
        if a > b:
            fu = fm
        elif a < b:
            fl = fm
        else:
            break
        fmm = (fl + fu) / 2
        ittr += 1
    return fm, ittr


NUM_SAFETY = 4  # pst.c:882 (no STARFORM): minimum margin per rank when a store fills up


def wrap_low(c: np.ndarray, fInactive: float, fSplit: float) -> np.ndarray:
    """The particles pkdLowerPartWrap(d, fSplit1 = fSplitInactive, fSplit2 = fSplit) keeps on the lower side
    (pkd.c:1165-1211; pkdUpperPartWrap keeps the complement on the upper side)."""
    if fInactive > fSplit:
        return (c < fSplit) | (c >= fInactive)
    return (c < fSplit) & (c >= fInactive)


def reverse_split(c: np.ndarray, fSplit: float, bmin: float, bmax: float, nLowerStore: int, nUpperStore: int, nLower: int,
                  nUpper: int, prev_inactive=None):
    """pst.c:1049-1270: fSplitInactive of one node.  c = the node's coordinates along the split axis; n*Store = the free
    stores of the lower / upper ranks added up (pstFreeStore); prev_inactive = pst->fSplitInactive when bDoSplitDimFind = 0.
    Returns (fSplitInactive, fixed) -- fixed: one side's store was too small and the boundary moved into the cell."""
    ext = bmax - bmin
    nLeaves = nLower + nUpper
    fl = fSplit + 1e-6 * ext
    fu = fSplit - 1e-6 * ext

    def mid(fl, fu):
        if fu > fl:
            return 0.5 * (fl + fu)
        fmm = 0.5 * (fl + fu + ext)
        if fmm > bmax:
            fmm = 0.5 * (fl + fu - ext)
        return fmm

    if prev_inactive is not None:
        fm = prev_inactive
    else:
        fm = mid(fl, fu)
        fm = bmin - 1e-6 * ext if abs(fm - bmin) < abs(fm - bmax) else bmax + 1e-6 * ext

    def count(fm):
        nLow = int(np.count_nonzero(wrap_low(c, fm, fSplit)))
        return nLow, c.size - nLow

    nLowTot, nHighTot = count(fm)
    safety, sloppy = NUM_SAFETY, 2
    nSafeTot = nLowerStore + nUpperStore - (nLowTot + nHighTot)
    if nSafeTot <= sloppy * safety * nLeaves:
        sloppy = 1
    if int(nSafeTot / nLeaves) < safety:  # C integer division (truncation)
        safety = int(nSafeTot / nLeaves)
    margin = int(0.05 * nSafeTot / nLeaves)
    if margin < safety:
        margin = safety
    fixed = False
    if nLowTot > nLowerStore - safety * nLower:
        fixed = True
        fm = min(max(fm, bmin), bmax)
        fl = fm
        fmm = mid(fl, fu)
        ittr = 1
        while ittr < MAX_ITTR:
            fm = fmm
            nLowTot, nHighTot = count(fm)
            if nLowTot > nLowerStore - margin * nLower:
                fl = fm
            elif nLowTot < nLowerStore - sloppy * margin * nLower:
                fu = fm
            else:
                fl = fm
                break
            if fu == fl:
                break
            fmm = mid(fl, fu)
            ittr += 1
        assert nLowTot <= nLowerStore
    elif nHighTot > nUpperStore - safety * nUpper:
        fixed = True
        fm = min(max(fm, bmin), bmax)
        fu = fm
        fmm = mid(fl, fu)
        ittr = 1
        while ittr < MAX_ITTR:
            fm = fmm
            nLowTot, nHighTot = count(fm)
            if nHighTot > nUpperStore - margin * nUpper:
                fu = fm
            elif nHighTot < nUpperStore - sloppy * margin * nUpper:
                fl = fm
            else:
                fu = fm
                break
            if fu == fl:
                break
            fmm = mid(fl, fu)
            ittr += 1
        assert nHighTot <= nUpperStore
    return fm, fixed


def rank_stores(nThreads: int, nTotal: int, fExtraStore: float) -> np.ndarray:
    """pkd->nStore of every rank: the file is split down the rank tree (pstReadTipsy, pst.c:676-725: the lower ranks get
    nLower * (n / nLeaves) particles) and a rank's store is its share plus ceil(share * dExtraStore)."""
    from gasoline_b200.domain import pst_tree
    out = np.zeros(nThreads, np.int64)

    def walk(node, n):
        if node.leaf:
            out[node.ranks[0]] = n + int(np.ceil(n * fExtraStore))
            return
        nl = len(node.lower.ranks) * (n // len(node.ranks))
        walk(node.lower, nl)
        walk(node.upper, n - nl)

    walk(pst_tree(nThreads), nTotal)
    return out


def domain_decomp(x, y, z, nThreads: int, weights=None, split_work: bool = True, prev=None, do_root_find: bool = True,
                  do_split_dim_find: bool = True, stores=None):
    """-> (list of index arrays, one per rank, ascending; list of (iCell, d, fSplit, bnd[6][, fSplitInactive, fixed]) per
    interior node in pre-order, lower subtree first).  `weights` None: fWeight = 1 for every particle.  prev: {iCell: (d,
    fSplit[, fSplitInactive])} of the previous decomposition (pst->iSplitDim, pst->fSplit, pst->fSplitInactive), None on the
    first call.  stores: pkd->nStore of every rank (rank_stores) -- the reverse split of pst.c:1049-1270 then runs for every
    node; None: stores with room (the lower ranks receive r[d] < fSplit)."""
    from gasoline_b200.domain import pst_tree  # the rank tree of pstSetAdd (host logic, no CUDA)

    pos = np.stack([np.asarray(x, np.float64), np.asarray(y, np.float64), np.asarray(z, np.float64)], axis=1)
    w = None
    if split_work:
        w = np.ones(len(pos)) if weights is None else np.asarray(weights, np.float64)
    out = [None] * nThreads
    nodes = []

    def split(node, idx):
        if node.leaf:
            out[node.ranks[0]] = np.sort(idx)
            return
        p = pos[idx]
        lo, hi = p.min(axis=0), p.max(axis=0)
        pv = prev[node.iCell] if prev and node.iCell in prev else (-1, np.nan)
        pd, pf = pv[0], pv[1]
        d = split_dim(lo, hi, pd) if do_split_dim_find or pd == -1 else pd  # pst.c:1900, :945-947
        fm = pf
        if do_root_find or not (lo[d] <= fm <= hi[d]):  # pst.c:963: bDoRootFind || fm < fl || fm > fu
            fm, _ = root_split(p[:, d], None if w is None else w[idx], len(node.lower.ranks), len(node.upper.ranks),
                               float(lo[d]), float(hi[d]))
        if stores is None:
            nodes.append((node.iCell, d, fm, np.concatenate([lo, hi])))
            low = p[:, d] < fm
        else:
            keep = pv[2] if (not do_split_dim_find and len(pv) > 2) else None  # pst.c:1060
            fI, fixed = reverse_split(p[:, d], fm, float(lo[d]), float(hi[d]), int(sum(stores[r] for r in node.lower.ranks)),
                                      int(sum(stores[r] for r in node.upper.ranks)), len(node.lower.ranks),
                                      len(node.upper.ranks), keep)
            nodes.append((node.iCell, d, fm, np.concatenate([lo, hi]), fI, fixed))
            low = wrap_low(p[:, d], fI, fm)
        split(node.lower, idx[low])
        split(node.upper, idx[~low])

    split(pst_tree(nThreads), np.arange(len(pos)))
    return out, nodes
