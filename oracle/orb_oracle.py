"""CPU restatement (numpy) of the reference's ORB domain decomposition (every thread's store has room, so the inactive
"wrap" split of pst.c:1049-1270 never moves the boundary).  The first call of a run (pst->iSplitDim == -1,
bDoRootFind = bDoSplitDimFind = 1, master.c:4176-4177) is pinned by execution; later calls (`prev` = the cells' split
axis and split of the previous decomposition: the NEWSPLITDIMCUT hysteresis of pst.c:1900-1910, and the
bDoRootFind = 0 / bDoSplitDimFind = 0 shortcuts of master.c:4210-4222) are restated from the source only -- PARITY
UNPINNED for those, no time-stepping run of the compiled reference exists here.  TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import this; the product path (gasoline_b200/) must not.

PINNED: tests/test_oracle_orb.py checks it against the domains the compiled reference produced on 2, 3 and 4
pthread-MDL ranks (tests/golden/multirank_*.npz: r<k>_iOrder), particle for particle.

Follows, per node of the rank tree:
  * pstDomainDecomp (pst.c:1854-1935): the node's bounds are the bounds of its particles (pstCalcBound); the split
    dimension is the first axis of strictly largest extent (iSplitDim == -1 on the first call, pst.c:1900-1910);
  * _pstRootSplit (pst.c:959-1034): bisection on fSplit between the bounds, at most MAX_ITTR = 64 steps, while
    fl < fmm < fu; a particle is "low" when r[d] < fSplit (pkdLowerPart/pkdUpperPart, pkd.c:1064-1133); the branch
    taken compares fLow/nLower with fHigh/nUpper (bSplitWork, the default, master.c:964: weights of the particles,
    fWeight = 1 after reading a file) or nLow/nLower with nHigh/nUpper; equal shares or nLow == nHigh == 1 stop it;
  * the lower set of ranks receives the particles with r[d] < fSplit (pkdColRejects after the split).
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


def domain_decomp(x, y, z, nThreads: int, weights=None, split_work: bool = True, prev=None, do_root_find: bool = True,
                  do_split_dim_find: bool = True):
    """-> (list of index arrays, one per rank, ascending; list of (iCell, d, fSplit, bnd[6]) per interior node in
    pre-order, lower subtree first).  `weights` None: fWeight = 1 for every particle.  prev: {iCell: (d, fSplit)} of
    the previous decomposition (pst->iSplitDim, pst->fSplit), None on the first call."""
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
        pd, pf = prev[node.iCell] if prev and node.iCell in prev else (-1, np.nan)
        d = split_dim(lo, hi, pd) if do_split_dim_find or pd == -1 else pd  # pst.c:1900, :945-947
        fm = pf
        if do_root_find or not (lo[d] <= fm <= hi[d]):  # pst.c:963: bDoRootFind || fm < fl || fm > fu
            fm, _ = root_split(p[:, d], None if w is None else w[idx], len(node.lower.ranks), len(node.upper.ranks),
                               float(lo[d]), float(hi[d]))
        nodes.append((node.iCell, d, fm, np.concatenate([lo, hi])))
        low = p[:, d] < fm
        split(node.lower, idx[low])
        split(node.upper, idx[~low])

    split(pst_tree(nThreads), np.arange(len(pos)))
    return out, nodes
