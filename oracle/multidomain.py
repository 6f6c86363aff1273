"""The multi-rank walk for the CPU oracle.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's parity self-check).

In the reference a bucket of rank r walks the replicated top tree kdTop from ROOT (heap: LOWER(i) = 2i, UPPER(i) = 2i+1)
and, wherever it reaches a leaf of kdTop, continues in that rank's tree -- its own (pkdLocalWalk) or a remote one read
through the MDL cache (pkdRemoteWalk), walk.c:342-435.  Interior top cells are tested with INTERSECTNP against their
fOpen2 only (no "fewer than 4 particles" rule, walk.c:363-371); the cells of every rank's tree, its root included, obey
the rules of pkdLocalWalk (walk.c:58-178).

That is exactly ONE walk of ONE tree when the ranks' trees are hung under the interior top cells: this module builds
that combined tree -- interior top cells first, then every rank's nodes, threaded links (iLower = first child, iUpper =
next cell of the depth-first sweep) re-pointed across the seams, particles concatenated in the rank tree's leaf order so
that every cell covers a contiguous range (top cells then hold >= 4 particles and the "< 4" rule never fires for them)
-- and hands it to the single-tree oracle (gravity_oracle.c through oracle.OracleGravity(tree=...)).  Pinned by
tests/test_oracle_multidomain.py: per-bucket interaction-list counts of every rank equal the dumps of the reference
BINARY run on 2, 3 and 4 ranks (tests/golden/multirank_*.npz) bit for bit, forces to the v_sqrt1 tolerance (2e-7).
"""
from __future__ import annotations

import numpy as np

NMOM = 31


def combined_tree(pst, kdTop, ilcnRoot, ranks: dict, period):
    """pst: gasoline_b200.domain.PstNode tree; kdTop: dict of heap arrays (assemble_top: r, fMass, fSoft, fOpen2, mom, bnd);
    ilcnRoot: the distributed Ewald root (35); ranks[r] = dict(tree=<dict or Tree of rank r's local tree>, x, y, z, m, h
    [, active]) with the particles in that tree's order.  Returns (tree dict for OracleGravity(tree=...), nodeBase,
    partBase): rank r's node i is combined node nodeBase[r] + i, its particle j is combined particle partBase[r] + j."""
    def get(t, k):
        return t[k] if isinstance(t, dict) else getattr(t, k)

    leaves = list(pst.ranks)  # lower subtree first: the order in which a depth-first sweep meets the ranks
    interior = []

    def collect(n):
        if not n.leaf:
            interior.append(n)
            collect(n.lower)
            collect(n.upper)

    collect(pst)
    nI = len(interior)
    gidx = {n.iCell: k for k, n in enumerate(interior)}
    nodeBase, partBase = {}, {}
    nb, pb = nI, 0
    for r in leaves:
        nodeBase[r], partBase[r] = nb, pb
        nb += int(get(ranks[r]["tree"], "nNodes"))
        pb += len(ranks[r]["x"])
    nn, n = nb, pb
    T = dict(bnd=np.zeros((nn, 6)), r=np.zeros((nn, 3)), fMass=np.zeros(nn), fSoft=np.zeros(nn), fOpen2=np.zeros(nn),
             mom=np.zeros((nn, NMOM)), pLower=np.zeros(nn, np.int32), pUpper=np.zeros(nn, np.int32),
             iLower=np.full(nn, -1, np.int32), iUpper=np.full(nn, -1, np.int32))

    def entry(node):  # combined index of the cell that stands for a PST node
        if node.leaf:
            r = node.ranks[0]
            return nodeBase[r] + int(get(ranks[r]["tree"], "iRoot"))
        return gidx[node.iCell]

    def thread(node, nxt):
        """nxt: combined index of the cell after this subtree in the sweep (-1 at the end)"""
        if node.leaf:
            r = node.ranks[0]
            t, b, p0 = ranks[r]["tree"], nodeBase[r], partBase[r]
            k = int(get(t, "nNodes"))
            sl = slice(b, b + k)
            for f in ("bnd", "r", "fMass", "fSoft", "fOpen2", "mom"):
                T[f][sl] = get(t, f)
            T["pLower"][sl] = np.asarray(get(t, "pLower")) + p0
            T["pUpper"][sl] = np.asarray(get(t, "pUpper")) + p0
            lo, up = np.asarray(get(t, "iLower")), np.asarray(get(t, "iUpper"))
            T["iLower"][sl] = np.where(lo >= 0, lo + b, -1)
            T["iUpper"][sl] = np.where(up >= 0, up + b, nxt)  # the rank's own sweep ends -> continue in the top tree
            return
        g, i = gidx[node.iCell], node.iCell
        for f in ("r", "fMass", "fSoft", "fOpen2", "mom", "bnd"):
            T[f][g] = kdTop[f][i]
        first, last = node.ranks[0], node.ranks[-1]
        T["pLower"][g] = partBase[first]
        T["pUpper"][g] = partBase[last] + len(ranks[last]["x"]) - 1
        T["iLower"][g] = entry(node.lower)
        T["iUpper"][g] = nxt
        thread(node.lower, entry(node.upper))
        thread(node.upper, nxt)

    thread(pst, -1)
    cols = {}
    for k in ("x", "y", "z", "m", "h"):
        cols[k] = np.concatenate([np.asarray(ranks[r][k], dtype=np.float64) for r in leaves])
    act = np.concatenate([np.asarray(ranks[r].get("active", np.ones(len(ranks[r]["x"]), np.int32)), dtype=np.int32)
                          for r in leaves])
    T.update(cols, nNodes=nn, iRoot=entry(pst), active=act, root=np.asarray(ilcnRoot, dtype=np.float64),
             period=np.asarray(period, dtype=np.float64), iOrder=np.arange(n, dtype=np.int32))
    return T, nodeBase, partBase


def from_domains(doms, active_by_rank=None):
    """The combined tree of a list of gasoline_b200.domain.Domain objects with HOST-built trees, after their top tree has
    been assembled (Domain.assemble): returns (tree dict, nodeBase, partBase).  active_by_rank[r]: ACTIVE flags (tree
    order) of rank r's particles; default: none active (the caller marks the sinks it wants through the returned bases)."""
    d0 = doms[0]
    ranks = {}
    for d in doms:
        h = d.host
        a = np.zeros(h.nLocal, np.int32) if active_by_rank is None else np.asarray(active_by_rank[d.idSelf], np.int32)
        ranks[d.idSelf] = dict(tree=h.tree, x=h.x, y=h.y, z=h.z, m=h.fMass, h=h.fSoft, active=a)
    return combined_tree(d0.pst, d0.kdTop, d0.ilcnRoot, ranks, d0.fPeriod)
