"""Multi-GPU runs: one spatial domain per rank/GPU, the reference's top tree on every rank, remote trees over NCCL.

The reference splits the particles over its ranks by ORB (pstDomainDecomp, pst.c:1854), builds one local tree per
rank (pkdBuildBinary), hangs the local roots under a small "top tree" that mirrors the rank tree (PST) and that
every rank holds (pstBuildTree interior branch pst.c:2918-2951, pstColCells/pkdDistribCells), and then lets every
bucket walk top tree -> local tree or remote trees (pkdBucketWalk walk.c:342-435; remote cells and particles come
through the MDL software cache, element by element).  Here the same structures are assembled with collectives:

  1. all-gather of each rank's root cell (78 doubles)                      -> every rank knows all local roots
  2. every rank combines them bottom-up like pkdCombine (pkd.c:1973)       -> mass, centre, softening of each top cell
  3. every rank sums its particles' moments about the centre of each top cell above it (pkdCalcCell pkd.c:2018 via
     gg_cell_moments), all-gather, add in the PST's lower-then-upper order (pstCalcCell pst.c:3789), pkdCalcOpen
                                                                            -> kdTop, bit-identical to the reference's
  4. Ewald root expansion like pkdCalcRoot/pkdDistribRoot (pkd.c:4395-4493): m, centre, quadrupole of kdTop[ROOT];
     l = 3, 4 moments summed over the ranks' LOCAL expansions (the reference's convention)
  5. all-gather of the trees + particles themselves (NCCL over NVLink when the ranks are GPUs) and gg_set_remote:
     a push of whole domains instead of the reference's pull-on-demand cache.

`Domain` is the per-rank state and holds only host arrays, so steps 1-4 run (and are tested) without a GPU;
`attach()` hands the result to a PKD.  `run_in_process` drives several Domains inside one process (tests; several
contexts on one GPU); `DistributedExchange` drives one Domain per torch.distributed rank (bench.py --gpus N).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import pkd as _pkd
from .pkd import GG_NMOM, GG_NROOT, PKD, Tree

SUMMARY_LEN = 6 + 3 + 3 + GG_NMOM + GG_NROOT  # bnd, r, (fMass, fSoft, fOpen2), mom, local root expansion


# ---------------------------------------------------------------------------------------------- rank tree (PST)
@dataclass
class PstNode:
    iCell: int                 # heap index in kdTop: ROOT = 1, LOWER(i) = 2i, UPPER(i) = 2i+1 (pkd.h:77-86)
    ranks: list                # ranks below this node, lower subtree first
    lower: "PstNode | None" = None
    upper: "PstNode | None" = None

    @property
    def leaf(self) -> bool:
        return self.lower is None


def pst_tree(nThreads: int) -> PstNode:
    """The rank tree msrInitialize grows by adding ranks 1..n-1 one at a time with pstSetAdd (pst.c:598-625): a new
    rank goes to the upper side when the lower side is heavier, else to the lower side; a leaf that receives a rank
    becomes a node with itself as the lower and the newcomer as the upper leaf."""
    class N:
        def __init__(self, rank):
            self.rank, self.nLeaves, self.nLower, self.nUpper, self.lo, self.up = rank, 1, 0, 0, None, None

    def add(n, rank):
        if n.nLeaves > 1:
            n.nLeaves += 1
            if n.nLower > n.nUpper:
                n.nUpper += 1
                add(n.up, rank)
            else:
                n.nLower += 1
                add(n.lo, rank)
        else:
            n.nLeaves, n.nLower, n.nUpper = 2, 1, 1
            n.lo, n.up = N(n.rank), N(rank)

    root = N(0)
    for r in range(1, nThreads):
        add(root, r)

    def conv(n, iCell):
        if n.lo is None:
            return PstNode(iCell, [n.rank])
        lo, up = conv(n.lo, 2 * iCell), conv(n.up, 2 * iCell + 1)
        return PstNode(iCell, lo.ranks + up.ranks, lo, up)

    return conv(root, 1)


def top_cells(nThreads: int) -> int:
    """nCell of kdTop: 2^(1+ceil(log2 nThreads)) (master.c:4293)."""
    return 1 << (1 + int(np.ceil(np.log(float(nThreads)) / np.log(2.0)))) if nThreads > 1 else 2


def interior_nodes(root: PstNode) -> list:
    out = []

    def walk(n):
        if not n.leaf:
            out.append(n)
            walk(n.lower)
            walk(n.upper)

    walk(root)
    return out


# ---------------------------------------------------------------------------------------------- decomposition
def orb_decompose(x, y, z, nThreads: int, weights=None) -> list:
    """ORB over the rank tree (pstDomainDecomp pst.c:1854): at every node split the LONGEST axis of the node's
    bounding box so that the lower side receives the share nLower/nLeaves of the weight (particle count when
    weights is None -- the reference's bSplitWork=0 path; pass last step's fWeight for work balancing).
    Returns one index array per rank.  (A Gasoline host does this itself; bench.py and the tests use it.)"""
    pos = np.stack([np.asarray(x), np.asarray(y), np.asarray(z)], axis=1)
    w = np.ones(len(pos)) if weights is None else np.asarray(weights, dtype=np.float64)
    out = [None] * nThreads

    def split(node, idx):
        if node.leaf:
            out[node.ranks[0]] = np.sort(idx)
            return
        p = pos[idx]
        d = int(np.argmax(p.max(axis=0) - p.min(axis=0)))
        order = np.argsort(p[:, d], kind="stable")
        cw = np.cumsum(w[idx][order])
        share = len(node.lower.ranks) / len(node.ranks)
        k = int(np.searchsorted(cw, share * cw[-1], side="left")) + 1
        k = min(max(k, 1), len(idx) - 1)
        split(node.lower, idx[order[:k]])
        split(node.upper, idx[order[k:]])

    split(pst_tree(nThreads), np.arange(len(pos)))
    return out


MAX_ITTR = 64  # pst.c:874


NEWSPLITDIMCUT = 0.707  # pst.c:1851


def pst_domain_decomp(ranks: list, nThreads: int, split_work: bool = True, reduce=None, prev=None,
                      bDoRootFind: bool = True, bDoSplitDimFind: bool = True, device_bisect: bool | None = None,
                      collective_bisect: bool = False, stores=None):
    """pstDomainDecomp (pst.c:1854-1935) with _pstRootSplit's root finder (pst.c:959-1034) for hosts that are not
    Gasoline: the first call of a run (bDoRootFind = bDoSplitDimFind = 1, master.c:4176) and, with `prev`, the later ones;
    with `stores` also the second boundary of ranks with fixed particle stores (pst.c:1049-1270).  The per-rank work --
    bounds, trial weights, the final split --
    is done on the device by every rank's PKD (pkdCalcBound, pkdWeight, pkdOrbSplit after pkdOrbLoad); this function
    only runs the bisection and adds the ranks' answers, level by level of the rank tree with all cells of a level in
    one request (the reference recurses cell by cell; the outcome per cell is the same).

    ranks:  the PKD-like objects of the ranks THIS process drives (all of them in a single process; one under torchrun)
    reduce: None, or f(kind, array) -> array combining the answers of the processes ("sum", "min", "max") --
            DistributedExchange.orb_reduce wraps torch.distributed.all_reduce.
    split_work: bSplitWork (default 1, master.c:964): compare fLow/nLower with fHigh/nUpper, else nLow/nLower, nHigh/nUpper.
    prev: the list this function returned for the previous decomposition (pst->iSplitDim / pst->fSplit of every cell):
            the split axis then only changes when another axis beats the old one's extent x NEWSPLITDIMCUT
            (pst.c:1900-1910); bDoSplitDimFind = 0 keeps the axis, bDoRootFind = 0 keeps the old split while it lies
            inside the cell's bounds (pst.c:963; the host sets both to 0 for small active sets, master.c:4210-4222).
    device_bisect: run each level's root finder on the device (gg_orb_bisect: no host round trip per trial).  Default:
            whenever ONE context holds all particles (len(ranks) == 1, no reduce) and offers it.  Same splits, bit for bit.
    collective_bisect: this process drives ONE rank whose context has a library communicator (commInitNccl /
            commInitLocal) over all ranks of the decomposition: each level's root finder is gg_orb_bisect_all -- the ranks'
            answers to a trial travel device to device (in-stream over NCCL), added in rank order like `reduce` does.
            `reduce` still combines the bounds (orb_reduce_lib keeps that below the C ABI too).
    stores: pkd->nStore of every rank (rank_stores) for hosts whose ranks keep the reference's fixed particle stores: the
            "reverse" split of pst.c:1049-1270 then runs for every cell -- a second boundary fSplitInactive, bisected into the
            cell when one side's particles would not fit its ranks' stores, and the lower ranks receive the WRAPPED interval
            between the two boundaries (pkdColRejects -> pkdLowerPartWrap, pkd.c:1165-1211, 1463-1485).  The counts come from
            pkdWeight; a boundary that really moved is applied by the ranks' pkdOrbSplitWrap (gg_orb_split_wrap).  None
            (default): stores with room -- what a run with device-resident stores sized by need wants.
    Returns the interior PST cells as a list of dicts (iCell, iDim, fSplit, bnd, ittr[, fSplitInactive, fixed]) in level
    order; the particles' destination ranks are leaf_rank(nThreads)[pkd.pkdOrbCells()]."""
    can = len(ranks) == 1 and reduce is None and hasattr(ranks[0], "pkdOrbBisect")
    device_bisect = can if device_bisect is None else (device_bisect and can)
    if collective_bisect:
        if len(ranks) != 1 or reduce is None or getattr(ranks[0], "commSize", 0) != nThreads:
            raise _pkd.GasolineB200Error("pst_domain_decomp: collective_bisect needs one rank per caller, a reduce for the "
                                         "bounds and a communicator over all nThreads ranks")
        device_bisect = True
    old = {c["iCell"]: (c["iDim"], c["fSplit"]) for c in prev} if prev else {}
    oldInactive = {c["iCell"]: c["fSplitInactive"] for c in prev if "fSplitInactive" in c} if prev else {}
    def combine(kind, parts):
        a = parts[0].copy()
        for b in parts[1:]:
            a = a + b if kind == "sum" else (np.minimum(a, b) if kind == "min" else np.maximum(a, b))
        return reduce(kind, a) if reduce is not None else a

    level = [pst_tree(nThreads)]
    out = []
    while True:
        level = [n for n in level if not n.leaf]
        if not level:
            break
        if len(level) > 64:
            raise _pkd.GasolineB200Error("pst_domain_decomp: more than 64 PST cells on one level")
        ic = np.array([n.iCell for n in level], np.int32)
        got = [r.pkdCalcBound(ic) for r in ranks]
        lo = combine("min", [g[0][:, :3] for g in got])
        hi = combine("max", [g[0][:, 3:] for g in got])
        k = len(level)
        d = np.zeros(k, np.int32)
        fm = np.full(k, np.nan)
        for j in range(k):  # pst.c:1900-1910: first call (iSplitDim == -1) = the first axis of strictly largest extent
            pd, pf = old.get(int(ic[j]), (-1, np.nan))
            dj = pd
            if bDoSplitDimFind or pd == -1:
                dimsize = -1.0 if pd == -1 else (hi[j, pd] - lo[j, pd]) * NEWSPLITDIMCUT
                for a in range(3):
                    if hi[j, a] - lo[j, a] > dimsize:
                        dj, dimsize = a, hi[j, a] - lo[j, a]
            d[j], fm[j] = dj, pf
        fl, fu = lo[np.arange(k), d].copy(), hi[np.arange(k), d].copy()
        fmm = (fl + fu) / 2
        ittr = np.zeros(k, np.int32)
        nLower = np.array([len(n.lower.ranks) for n in level], np.float64)
        nUpper = np.array([len(n.upper.ranks) for n in level], np.float64)
        # pst.c:963: the root finder runs when asked for, or when the previous split has left the cell's bounds
        live = np.array([bool(bDoRootFind or not (fl[j] <= fm[j] <= fu[j])) for j in range(k)])
        fm[live] = np.nan
        if device_bisect:
            fs, has, it = ranks[0].pkdOrbBisect(ic, d, fl, fu, live, nLower, nUpper, split_work,
                                                **(dict(collective=True) if collective_bisect else {}))
            fm[has] = fs[has]
            ittr[:] = it
            live[:] = False
        while True:
            live &= (fl < fmm) & (fmm < fu) & (ittr < MAX_ITTR)
            if not live.any():
                break
            idx = np.nonzero(live)[0]
            fm[idx] = fmm[idx]
            got = [r.pkdWeight(ic[idx], d[idx], fm[idx]) for r in ranks]
            nLow = combine("sum", [g[0].astype(np.int64) for g in got])
            nHigh = combine("sum", [g[1].astype(np.int64) for g in got])
            if split_work:
                a = combine("sum", [g[2] for g in got]) / nLower[idx]
                b = combine("sum", [g[3] for g in got]) / nUpper[idx]
            else:
                a, b = nLow / nLower[idx], nHigh / nUpper[idx]
            for t, j in enumerate(idx):
                if (nLow[t] == 1 and nHigh[t] == 1) or a[t] == b[t]:
                    live[j] = False
                    continue
                if a[t] > b[t]:
                    fu[j] = fm[j]
                else:
                    fl[j] = fm[j]
                fmm[j] = (fl[j] + fu[j]) / 2
                ittr[j] += 1
        if np.isnan(fm).any():
            raise _pkd.GasolineB200Error("pst_domain_decomp: a PST cell has no extent along its longest axis")
        extra = [{} for _ in level]
        if stores is None:
            for r in ranks:
                r.pkdOrbSplit(ic, d, fm)
        else:
            bmin, bmax = lo[np.arange(k), d], hi[np.arange(k), d]

            def n_low(cells, f):  # particles of the cells with r[d] < f, over all ranks
                sel = np.asarray(cells, np.int64)
                got = [r.pkdWeight(ic[sel], d[sel], np.asarray(f, np.float64)) for r in ranks]
                return combine("sum", [g[0].astype(np.int64) for g in got])

            nIn = combine("sum", [np.asarray(r.pkdCalcBound(ic)[1], np.int64) for r in ranks])
            nLowSplit = n_low(np.arange(k), fm)
            fI = np.zeros(k)
            for j, n in enumerate(level):
                def count(f, j=j):  # pkdWeightWrap over the ranks: (nLowTot, nHighTot) of the wrapped interval
                    if f > fm[j]:
                        nl = int(nLowSplit[j]) + int(nIn[j]) - int(n_low([j], [f])[0])
                    else:
                        nl = int(nLowSplit[j]) - int(n_low([j], [f])[0])
                    return nl, int(nIn[j]) - nl
                keep = oldInactive.get(int(ic[j])) if not bDoSplitDimFind else None
                fI[j], fixed = _reverse_split(count, float(fm[j]), float(bmin[j]), float(bmax[j]),
                                              int(sum(stores[q] for q in n.lower.ranks)), int(sum(stores[q] for q in n.upper.ranks)),
                                              len(n.lower.ranks), len(n.upper.ranks), keep)
                extra[j] = dict(fSplitInactive=float(fI[j]), fixed=bool(fixed))
            plain = (fI > bmax) | (fI <= bmin)  # the wrapped interval is r[d] < fSplit for every particle of the cell
            if plain.all():
                for r in ranks:
                    r.pkdOrbSplit(ic, d, fm)
            else:
                for r in ranks:
                    if not hasattr(r, "pkdOrbSplitWrap"):
                        raise _pkd.GasolineB200Error(
                            "pst_domain_decomp: the split of PST cell %d sends a side more particles than its ranks' stores hold "
                            "(pst.c:1049-1270) and the ranks' services have no pkdOrbSplitWrap" % int(ic[np.nonzero(~plain)[0][0]]))
                    r.pkdOrbSplitWrap(ic, d, fm, fI)
        for j, n in enumerate(level):
            out.append(dict(iCell=n.iCell, iDim=int(d[j]), fSplit=float(fm[j]), bnd=np.concatenate([lo[j], hi[j]]),
                            ittr=int(ittr[j]), **extra[j]))
        level = [c for n in level for c in (n.lower, n.upper)]
    return out


NUM_SAFETY = 4  # pst.c:882 (no STARFORM): minimum margin per rank when a store fills up


def _reverse_split(count, fSplit: float, bmin: float, bmax: float, nLowerStore: int, nUpperStore: int, nLower: int, nUpper: int,
                   prev_inactive=None):
    """pst.c:1049-1270 for one PST cell: the second boundary fSplitInactive.  count(f) -> (nLowTot, nHighTot) of the wrapped
    interval between f and fSplit (pstWeightWrap).  Starts just outside the bounds (or from the previous decomposition's
    value when bDoSplitDimFind = 0, pst.c:1060); when a side holds more than its ranks' stores less a safety margin, the
    boundary is bisected (midpoints wrap around the cell's extent) until the side fits to within 5 % of the free space.
    Returns (fSplitInactive, moved into the cell?)."""
    ext = bmax - bmin
    nLeaves = nLower + nUpper
    fl, fu = fSplit + 1e-6 * ext, fSplit - 1e-6 * ext

    def mid(fl, fu):
        if fu > fl:
            return 0.5 * (fl + fu)
        fmm = 0.5 * (fl + fu + ext)
        return 0.5 * (fl + fu - ext) if fmm > bmax else fmm

    if prev_inactive is not None:
        fm = prev_inactive
    else:
        fm = mid(fl, fu)
        fm = bmin - 1e-6 * ext if abs(fm - bmin) < abs(fm - bmax) else bmax + 1e-6 * ext
    nLowTot, nHighTot = count(fm)
    safety, sloppy = NUM_SAFETY, 2
    nSafeTot = nLowerStore + nUpperStore - (nLowTot + nHighTot)
    if nSafeTot <= sloppy * safety * nLeaves:
        sloppy = 1  # no slop to play with
    if int(nSafeTot / nLeaves) < safety:
        safety = int(nSafeTot / nLeaves)
    margin = max(int(0.05 * nSafeTot / nLeaves), safety)  # only accurate to 5 % of the available space
    lowSide = nLowTot > nLowerStore - safety * nLower
    if not lowSide and not nHighTot > nUpperStore - safety * nUpper:
        return fm, False
    fm = min(max(fm, bmin), bmax)
    if lowSide:
        fl = fm
    else:
        fu = fm
    fmm = mid(fl, fu)
    for _ in range(1, MAX_ITTR):
        fm = fmm
        nLowTot, nHighTot = count(fm)
        nSide, store, nRanks = (nLowTot, nLowerStore, nLower) if lowSide else (nHighTot, nUpperStore, nUpper)
        over, under = nSide > store - margin * nRanks, nSide < store - sloppy * margin * nRanks
        if lowSide:
            if over or not under:
                fl = fm
            else:
                fu = fm
        else:
            if over or not under:
                fu = fm
            else:
                fl = fm
        if not over and not under:
            break
        if fu == fl:
            break
        fmm = mid(fl, fu)
    if (nLowTot if lowSide else nHighTot) > (nLowerStore if lowSide else nUpperStore):
        raise _pkd.GasolineB200Error("pst_domain_decomp: the particles do not fit the ranks' stores (pst.c:1194 / 1253)")
    return fm, True


def rank_stores(nThreads: int, nTotal: int, fExtraStore: float = 0.1) -> np.ndarray:
    """pkd->nStore of every rank of a run that read nTotal particles: the file is split down the rank tree (pstReadTipsy,
    pst.c:676-725: the lower ranks get nLower * (n / nLeaves) particles) and a rank's store holds its share plus
    ceil(share * dExtraStore) (dExtraStore defaults to 0.1, master.c:883)."""
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


def leaf_rank(nThreads: int) -> np.ndarray:
    """PST heap index -> rank for the leaves of the rank tree (-1 elsewhere)."""
    m = np.full(top_cells(nThreads) * 2, -1, np.int32)

    def walk(n):
        if n.leaf:
            m[n.iCell] = n.ranks[0]
        else:
            walk(n.lower)
            walk(n.upper)

    walk(pst_tree(nThreads))
    return m


def orb_reduce_dist(backend_device: str):
    """The `reduce` of pst_domain_decomp over torch.distributed: every process contributes its ranks' combined answer;
    the answers are all-gathered and combined in rank order on every process, so all processes see the same bits and
    take the same branch of the bisection (what the reference gets from adding outWtLow and outWtHigh up the PST)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()

    def reduce(kind, a):
        a = np.ascontiguousarray(a)
        t = torch.from_numpy(a.reshape(-1).copy()).to(backend_device)
        out = torch.empty(world * t.numel(), dtype=t.dtype, device=backend_device)
        dist.all_gather_into_tensor(out, t)
        g = out.cpu().numpy().reshape((world,) + a.shape)
        r = g[0].copy()
        for k in range(1, world):
            r = r + g[k] if kind == "sum" else (np.minimum(r, g[k]) if kind == "min" else np.maximum(r, g[k]))
        return r

    return reduce


def orb_reduce_lib(pkd):
    """The `reduce` of pst_domain_decomp over the library's own communicator (gg_comm_allgather of the small host arrays,
    NCCL or in-process group), combined in rank order on every rank -- no torch.distributed on the decomposition's path."""
    def reduce(kind, a):
        a = np.ascontiguousarray(a)
        g = pkd.commAllgather(a)
        r = g[0].copy()
        for k in range(1, g.shape[0]):
            r = r + g[k] if kind == "sum" else (np.minimum(r, g[k]) if kind == "min" else np.maximum(r, g[k]))
        return r

    return reduce


def orb_exchange(cols: np.ndarray, dest, backend_device: str) -> np.ndarray:
    """After the decomposition: rows of cols [n][k] (float64 particle records) travel to process dest[i] by ONE
    all_to_all_single (the outcome of the reference's pkdColRejects / pkdSwapRejects rounds, pst.c:1275-1334).
    Returns the rows this process now owns, grouped by sender in rank order, each sender's rows in their old order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dest = np.asarray(dest)
    order = np.argsort(dest, kind="stable")
    send_counts = np.bincount(dest, minlength=world).astype(np.int64)
    k = cols.shape[1]
    cnt_t = torch.from_numpy(send_counts).to(backend_device)
    recv_t = torch.empty(world, dtype=torch.int64, device=backend_device)
    dist.all_to_all_single(recv_t, cnt_t)
    recv_counts = recv_t.cpu().numpy()
    send = torch.from_numpy(np.ascontiguousarray(cols[order], dtype=np.float64).reshape(-1)).to(backend_device)
    recv = torch.empty(int(recv_counts.sum()) * k, dtype=torch.float64, device=backend_device)
    dist.all_to_all_single(recv, send, output_split_sizes=[int(c) * k for c in recv_counts],
                           input_split_sizes=[int(c) * k for c in send_counts])
    return recv.cpu().numpy().reshape(-1, k)


# ---------------------------------------------------------------------------------------------- per-rank state
class Domain:
    """One rank: its particles (tree order) and local tree on the host; the top tree once assemble() has run."""

    def __init__(self, idSelf: int, nThreads: int, x, y, z, m, h, fPeriod, theta: float, nBucket: int = 8,
                 iOrder: int = 4, active=None, pinned: bool = False, device: int | None = None,
                 device_build: bool = False):
        """device_build: the local tree is built ON THE GPU from the particles (gg_build_local) and never exists on the
        host; the rank's root summary and its ancestor sums come from the device (gg_domain_summary,
        gg_domain_moments_about).  Needs a device."""
        self.idSelf, self.nThreads, self.theta, self.iOrder = idSelf, nThreads, float(theta), iOrder
        self.nBucket = nBucket
        self.device_build = bool(device_build)
        if self.device_build and device is None:
            raise _pkd.GasolineB200Error("Domain: device_build needs a device")
        self.fPeriod = tuple(float(v) for v in fPeriod)
        self.pst = pst_tree(nThreads)
        self.L = _pkd.load_library()
        # the tree builder is host code; a PKD (GPU context) is only created when a device is given
        self.pkd = PKD(device=device, idSelf=idSelf, fPeriod=fPeriod, pinned=pinned) if device is not None else None
        host = self.pkd if self.pkd is not None else _HostStore()
        host.pkdLoadParticles(x, y, z, m, h, active)
        self.host = host
        if self.device_build:
            self.rebuild()
        else:
            _build(host, nBucket, theta, iOrder)
            # this rank's OWN pkdCalcRoot expansion: attach() later replaces host.ilcnRoot by the distributed one
            self.localRoot = np.array(host.ilcnRoot, dtype=np.float64, copy=True)
        self.kdTop = None
        self.ilcnRoot = None

    def rebuild(self):
        """(device_build) pkdBuildBinary on the GPU from the rank's current particles; the domain is left loaded."""
        self.pkd.pkdBuildBinaryDevice(self.nBucket, self.theta)
        self.kdTop = None

    # -- step 1
    def summary(self) -> np.ndarray:
        if self.device_build:
            s = self.pkd.pkdDomainSummary()
            return np.concatenate([s["bnd"], s["r"], [s["fMass"], s["fSoft"], s["fOpen2"]], s["mom"],
                                   s["root"]]).astype(np.float64)
        t, r = self.host.tree, self.host.tree.iRoot
        return np.concatenate([t.bnd[r], t.r[r], [t.fMass[r], t.fSoft[r], t.fOpen2[r]], t.mom[r],
                               self.localRoot]).astype(np.float64)

    # -- steps 2 + 3 (local part)
    def ancestor_moments(self, summaries: np.ndarray) -> np.ndarray:
        """[nInterior][32]: this rank's pkdCalcCell sums (31 moments + Bmax) about the centre of every top cell above
        it, zeros for top cells that do not contain it."""
        cells = combine_top(self.pst, summaries)
        nodes = interior_nodes(self.pst)
        out = np.zeros((len(nodes), GG_NMOM + 1))
        h = self.host
        for k, n in enumerate(nodes):
            if self.idSelf not in n.ranks:
                continue
            rcm = np.ascontiguousarray(cells[n.iCell]["r"], dtype=np.float64)
            if self.device_build:
                out[k, :GG_NMOM], out[k, GG_NMOM] = self.pkd.pkdDomainMomentsAbout(rcm)
                continue
            mom, bmax = np.zeros(GG_NMOM), C.c_double()
            rc = self.L.gg_cell_moments(h.nLocal, _pkd._d(h.x), _pkd._d(h.y), _pkd._d(h.z), _pkd._d(h.fMass),
                                        _pkd._d(rcm), self.iOrder, _pkd._d(mom), C.byref(bmax))
            if rc != 0:
                raise _pkd.GasolineB200Error(f"gg_cell_moments failed ({rc})")
            out[k, :GG_NMOM], out[k, GG_NMOM] = mom, bmax.value
        return out

    # -- steps 3 (sum) + 4
    def assemble(self, summaries: np.ndarray, anc: np.ndarray):
        """summaries [nThreads][SUMMARY_LEN], anc [nThreads][nInterior][32] -> self.kdTop (dict of heap arrays) and
        self.ilcnRoot, the same on every rank."""
        self.kdTop, self.ilcnRoot = assemble_top(self.pst, self.nThreads, summaries, anc, self.theta)
        return self.kdTop, self.ilcnRoot

    # -- step 5
    def export_raw(self):
        """(doubles, ints) holding this domain's tree + particles in the field order gg_set_remote reads."""
        t, h = self.host.tree, self.host
        dbl = np.concatenate([t.r.ravel(), t.fMass, t.fSoft, t.fOpen2, t.mom.ravel(), h.x, h.y, h.z, h.fMass, h.fSoft])
        ints = np.concatenate([[t.nNodes, t.iRoot, h.nLocal, 0], t.pLower, t.pUpper, t.iLower, t.iUpper]).astype(np.int32)
        return np.ascontiguousarray(dbl, dtype=np.float64), ints

    def attach(self, upload: bool = True, announce=None):
        """Hand kdTop + ilcnRoot to the GPU context (gg_set_local via upload, gg_set_top, gg_set_root_moments).
        upload=False: the local domain is already resident (only the top tree / Ewald root go up).  announce: the
        parameters of the force evaluation that follows (PKD.upload)."""
        if self.pkd is None:
            raise _pkd.GasolineB200Error("Domain.attach: created without a device")
        self.pkd.pkdDistribRoot(self.ilcnRoot)  # (first: an announced Ewald correction starts during the upload)
        if upload and not self.device_build:  # (a device-built domain is already loaded)
            self.pkd.upload(announce=announce)
        k = self.kdTop
        self.pkd.pkdDistribCells(k["pLower"], k["bUsed"], k["r"], k["fMass"], k["fSoft"], k["fOpen2"], k["mom"])

    def set_remote_raw(self, id_: int, dbl, ints, dbl_ptr: int | None = None, int_ptr: int | None = None):
        """A remote domain from export_raw() data.  With dbl_ptr/int_ptr (device addresses of the same layout, e.g.
        inside an NCCL all-gather buffer) nothing is staged through the host: gg_set_remote(bDevice=1)."""
        nn, iRoot, n = int(ints[0]), int(ints[1]), int(ints[2])
        esz_d, esz_i = 8, 4
        base_d = dbl_ptr if dbl_ptr is not None else dbl.ctypes.data
        base_i = (int_ptr if int_ptr is not None else ints.ctypes.data) + 4 * esz_i
        off = 0

        def dptr(count):
            nonlocal off
            p = C.cast(base_d + off * esz_d, _pkd._dp)
            off += count
            return p

        tv = _pkd.gg_tree()
        tv.nNodes, tv.iRoot = nn, iRoot
        tv.bnd = None
        tv.r, tv.fMass, tv.fSoft, tv.fOpen2, tv.mom = dptr(3 * nn), dptr(nn), dptr(nn), dptr(nn), dptr(GG_NMOM * nn)
        pv = _pkd.gg_particles()
        pv.n = n
        pv.x, pv.y, pv.z, pv.fMass, pv.fSoft = dptr(n), dptr(n), dptr(n), dptr(n), dptr(n)
        pv.active = None
        ip = lambda k: C.cast(base_i + k * nn * esz_i, _pkd._ip)
        tv.pLower, tv.pUpper, tv.iLower, tv.iUpper = ip(0), ip(1), ip(2), ip(3)
        _pkd._check(self.L.gg_set_remote(self.pkd._ctx, id_, C.byref(tv), C.byref(pv), 1 if dbl_ptr is not None else 0),
                    "gg_set_remote")


class _HostStore:
    """The particle/tree half of PKD without a GPU context (CPU-only hosts and tests)."""
    pinned = False

    def __init__(self):
        self._L = _pkd.load_library()

    _own = PKD._own
    pkdLoadParticles = PKD.pkdLoadParticles
    pkdBuildBinary = PKD.pkdBuildBinary


def _build(host, nBucket, theta, iOrder):
    host.pkdBuildBinary(nBucket, theta, iOrder)


# ---------------------------------------------------------------------------------------------- top tree arithmetic
def _unpack(s):
    return dict(bnd=s[0:6], r=s[6:9], fMass=s[9], fSoft=s[10], fOpen2=s[11], mom=s[12:12 + GG_NMOM],
                root=s[12 + GG_NMOM:12 + GG_NMOM + GG_NROOT])


def combine_top(pst: PstNode, summaries: np.ndarray) -> dict:
    """pkdCombine (pkd.c:1973-2015) bottom-up over the rank tree: {iCell: {bnd, r, fMass, fSoft}}; leaves carry the
    ranks' root cells.  Plain double arithmetic in the reference's operation order (products, then one sum, then
    the division), so every rank -- and the reference -- gets the same bits."""
    cells = {}

    def walk(n):
        if n.leaf:
            cells[n.iCell] = _unpack(summaries[n.ranks[0]])
            return cells[n.iCell]
        a, b = walk(n.lower), walk(n.upper)
        bnd = np.concatenate([np.where(b["bnd"][:3] < a["bnd"][:3], b["bnd"][:3], a["bnd"][:3]),
                              np.where(b["bnd"][3:] > a["bnd"][3:], b["bnd"][3:], a["bnd"][3:])])
        m1, m2 = float(a["fMass"]), float(b["fMass"])
        mass = m1 + m2
        soft = m1 * float(a["fSoft"]) + m2 * float(b["fSoft"])
        r = np.array([m1 * float(a["r"][j]) + m2 * float(b["r"][j]) for j in range(3)])
        if mass > 0:
            soft /= mass
            r = np.array([float(v) / mass for v in r])
        cells[n.iCell] = dict(bnd=bnd, r=r, fMass=mass, fSoft=soft)
        return cells[n.iCell]

    walk(pst)
    return cells


def calc_open(bmax: float, theta: float) -> float:
    """pkdCalcOpen, OPEN_JOSH (pkd.c:2253-2260), squared (pst.c:2950)."""
    d = 2 / np.sqrt(3.0) * bmax / theta
    if d < bmax:
        d = bmax
    return float(d * d)


def assemble_top(pst: PstNode, nThreads: int, summaries: np.ndarray, anc: np.ndarray, theta: float):
    nCell = top_cells(nThreads)
    cells = combine_top(pst, summaries)
    nodes = interior_nodes(pst)
    index = {n.iCell: k for k, n in enumerate(nodes)}
    top = dict(pLower=np.full(nCell, -1, np.int32), bUsed=np.zeros(nCell, np.int32), r=np.zeros((nCell, 3)),
               fMass=np.zeros(nCell), fSoft=np.zeros(nCell), fOpen2=np.zeros(nCell), mom=np.zeros((nCell, GG_NMOM)),
               bnd=np.zeros((nCell, 6)))

    def moment_sum(sub: PstNode, k: int):
        """pstCalcCell (pst.c:3789-3845): lower subtree + upper subtree, element by element; Bmax is a maximum."""
        if sub.leaf:
            return anc[sub.ranks[0]][k].copy()
        a, b = moment_sum(sub.lower, k), moment_sum(sub.upper, k)
        out = a.copy()
        out[:GG_NMOM] = a[:GG_NMOM] + b[:GG_NMOM]
        out[GG_NMOM] = b[GG_NMOM] if b[GG_NMOM] > a[GG_NMOM] else a[GG_NMOM]
        return out

    def fill(n: PstNode):
        c, i = cells[n.iCell], n.iCell
        top["bUsed"][i] = 1
        top["r"][i], top["fMass"][i], top["fSoft"][i], top["bnd"][i] = c["r"], c["fMass"], c["fSoft"], c["bnd"]
        if n.leaf:
            top["pLower"][i] = n.ranks[0]
            top["fOpen2"][i], top["mom"][i] = c["fOpen2"], c["mom"]
            return
        s = moment_sum(n, index[i])
        top["mom"][i] = s[:GG_NMOM]
        top["fOpen2"][i] = calc_open(float(s[GG_NMOM]), theta)
        fill(n.lower)
        fill(n.upper)

    fill(pst)

    # pkdCalcRoot summed like pstCalcRoot (lower + upper), then pkdDistribRoot's overwrite of m, centre, quadrupole
    def root_sum(sub: PstNode):
        if sub.leaf:
            return _unpack(summaries[sub.ranks[0]])["root"].copy()
        return root_sum(sub.lower) + root_sum(sub.upper)

    root = root_sum(pst)
    q = top["mom"][1]
    root[0] = top["fMass"][1]
    root[1:4] = top["r"][1]
    root[4], root[5], root[6], root[7], root[8], root[9] = q[0], q[1], q[3], q[4], q[5], q[2]  # xx yy xy xz yz zz
    return top, root


# ---------------------------------------------------------------------------------------------- drivers
def run_in_process(domains: list, exchange_trees: bool = True, packed: bool = False, let=None):
    """All ranks inside this process (tests; several contexts on one GPU): the all-gathers are list comprehensions.
    packed=True moves the trees device to device in record layout (gg_export_local -> gg_set_remote_packed), the form
    DistributedExchange.exchange_packed sends through NCCL."""
    summaries = np.stack([d.summary() for d in domains])
    anc = np.stack([d.ancestor_moments(summaries) for d in domains])
    for d in domains:
        d.assemble(summaries, anc)
    if exchange_trees and let is not None:
        # pruned locally-essential trees (gg_let_export), `let` = the GravityParams of the coming force evaluation
        import torch
        for d in domains:
            d.attach()
        bnds = {d.idSelf: summaries[d.idSelf][0:6] for d in domains}
        stats = {}
        for o in domains:
            others = [d for d in domains if d.idSelf != o.idSelf]
            ptr, offs, hdr = o.pkd.let_export(np.stack([bnds[d.idSelf] for d in others]), let)
            full, _ = o.pkd.export_size()
            for k, d in enumerate(others):
                d.pkd.pkdSetRemotePacked(o.idSelf, hdr[k], ptr + int(offs[k]))  # (copied before o's next export)
                stats[(o.idSelf, d.idSelf)] = (int(offs[k + 1] - offs[k]), full)
        torch.cuda.synchronize()
        run_in_process.let_stats = stats
        return summaries, anc
    if exchange_trees and packed:
        import torch
        for d in domains:
            d.attach()
        bufs = []
        for d in domains:
            nbytes, hdr = d.pkd.export_size()
            b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            d.pkd.export_local(b.data_ptr())
            bufs.append((b, hdr))
        torch.cuda.synchronize()
        for d in domains:
            for o, (b, hdr) in zip(domains, bufs):
                if o.idSelf != d.idSelf:
                    d.pkd.pkdSetRemotePacked(o.idSelf, hdr, b.data_ptr())
        return summaries, anc
    if exchange_trees and all(d.pkd is not None for d in domains):
        raws = [d.export_raw() for d in domains]
        for d in domains:
            d.attach()
            for o, (dbl, ints) in zip(domains, raws):
                if o.idSelf != d.idSelf:
                    d.set_remote_raw(o.idSelf, dbl, ints)
    return summaries, anc


class DistributedExchange:
    """One Domain per torch.distributed rank.  Small collectives (steps 1-4) go through host tensors on gloo or
    device tensors on NCCL; the bulk tree exchange (step 5) is one padded all_gather_into_tensor per element type
    on the device, and the remote domains are ingested straight from the gather buffer (no host staging)."""

    def __init__(self, domain: Domain, backend_device: str):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.d, self.dev = torch, dist, domain, backend_device
        self.world = dist.get_world_size()
        self._bufs = None
        self._pbufs = None
        self._lbufs = None
        self._summaries = None

    def _all_gather_np(self, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a)
        t = self.torch.from_numpy(a.reshape(-1)).to(self.dev)
        out = self.torch.empty(self.world * t.numel(), dtype=t.dtype, device=self.dev)
        self.dist.all_gather_into_tensor(out, t)  # flat in, flat out: the form both gloo and NCCL accept
        return out.cpu().numpy().reshape((self.world,) + a.shape)

    def top_tree(self):
        summaries = self._all_gather_np(self.d.summary())
        anc = self._all_gather_np(self.d.ancestor_moments(summaries))
        self.d.assemble(summaries, anc)
        self._summaries = summaries
        return summaries, anc

    def exchange(self, attach: bool = True):
        """Steps 1-5.  Returns the bytes this rank received over the interconnect.  self.timing holds the seconds of
        each phase of the last call (host wall clock, device work synchronised at the phase boundaries)."""
        import time
        torch, dist, d = self.torch, self.dist, self.d
        tm = {}
        t0 = time.perf_counter()

        def lap(name):
            nonlocal t0
            if self.dev != "cpu":
                torch.cuda.synchronize()
            t1 = time.perf_counter()
            tm[name] = tm.get(name, 0.0) + (t1 - t0)
            t0 = t1

        self.top_tree()
        lap("top_tree")
        dbl, ints = d.export_raw()
        lap("export_raw")
        sizes = self._all_gather_np(np.array([dbl.size, ints.size], dtype=np.int64))
        nd, ni = int(sizes[:, 0].max()), int(sizes[:, 1].max())
        if self._bufs is None or self._bufs[2].numel() != nd or self._bufs[3].numel() != ni:
            self._bufs = (torch.empty(self.world * nd, dtype=torch.float64, device=self.dev),
                          torch.empty(self.world * ni, dtype=torch.int32, device=self.dev),
                          torch.zeros(nd, dtype=torch.float64, device=self.dev),
                          torch.zeros(ni, dtype=torch.int32, device=self.dev))
        fd, fi, sd, si = self._bufs
        sd[:dbl.size].copy_(torch.from_numpy(dbl), non_blocking=True)
        si[:ints.size].copy_(torch.from_numpy(ints), non_blocking=True)
        lap("h2d")
        dist.all_gather_into_tensor(fd, sd)  # every domain padded to the largest: one collective per element type
        dist.all_gather_into_tensor(fi, si)
        gd, gi = fd.view(self.world, nd), fi.view(self.world, ni)
        lap("all_gather")
        if attach and d.pkd is not None:
            d.attach()
            lap("attach_local")
            hosts = gi[:, :4].cpu().numpy()
            for r in range(self.world):
                if r == d.idSelf:
                    continue
                if self.dev == "cpu":
                    d.set_remote_raw(r, gd[r].numpy(), gi[r].numpy())
                else:
                    d.set_remote_raw(r, None, hosts[r], gd[r].data_ptr(), gi[r].data_ptr())
            lap("set_remote")
        self.gathered = (gd, gi, sizes)
        self.timing = tm
        return int((sizes[:, 0].sum() - dbl.size) * 8 + (sizes[:, 1].sum() - ints.size) * 4)


    def exchange_packed(self, top: bool = True):
        """The per-step exchange on GPUs, device to device: this rank's upload (gg_set_local), then ONE padded NCCL
        all-gather of every domain's device records (64 B walk + 128 B moment + 48 B quadrupole record per node, 32 B
        per particle) and ingestion straight from the receive buffer (gg_set_remote_packed).  top=False reuses the
        top tree / Ewald root of the last top_tree() call (they belong to the tree build, like pkdBuildBinary).
        Returns the bytes this rank received over NVLink."""
        import time
        torch, dist, d = self.torch, self.dist, self.d
        tm = {}
        t0 = time.perf_counter()

        def lap(name):
            nonlocal t0
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            tm[name] = t1 - t0
            t0 = t1

        if top or d.kdTop is None:
            self.top_tree()
            lap("top_tree")
        d.attach()
        lap("attach_local")
        nbytes, hdr = d.pkd.export_size()
        meta = self._all_gather_np(np.array([nbytes, hdr[0], hdr[1], hdr[2]], dtype=np.int64))
        cap = int(meta[:, 0].max())
        cap = (cap + 255) // 256 * 256
        if self._pbufs is None or self._pbufs[0].numel() < cap:
            self._pbufs = (torch.empty(cap, dtype=torch.uint8, device=self.dev),
                           torch.empty(self.world * cap, dtype=torch.uint8, device=self.dev))
        send, recv = self._pbufs
        cap = send.numel()
        d.pkd.export_local(send.data_ptr())
        lap("export")
        dist.all_gather_into_tensor(recv, send)
        lap("all_gather")
        for r in range(self.world):
            if r != d.idSelf:
                d.pkd.pkdSetRemotePacked(r, meta[r, 1:4], recv.data_ptr() + r * cap)
        lap("set_remote")
        self.timing = tm
        return int(meta[:, 0].sum() - nbytes)


    def exchange_let(self, g, top: bool = True):
        """The per-step exchange with PRUNED trees: this rank's upload, gg_let_export against every other domain's root
        bounds (known from the top-tree summaries), sizes by a small all-gather, ONE NCCL all-to-all of the pruned
        domains in device record layout, gg_set_remote_packed from the receive buffer.  Returns bytes received."""
        import time
        torch, dist, d = self.torch, self.dist, self.d
        tm = {}
        t0 = time.perf_counter()

        def lap(name):
            nonlocal t0
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            tm[name] = t1 - t0
            t0 = t1

        if top or d.kdTop is None:
            self._summaries, _ = self.top_tree()
            lap("top_tree")
        d.attach()
        lap("attach_local")
        others = [r for r in range(self.world) if r != d.idSelf]
        ptr, offs, hdr = d.pkd.let_export(np.stack([self._summaries[r][0:6] for r in others]), g)
        lap("let_export")
        # meta[r] = what THIS rank sends to rank r: bytes, nNodes, nPart, iRoot
        meta = np.zeros((self.world, 4), dtype=np.int64)
        for k, r in enumerate(others):
            meta[r] = (offs[k + 1] - offs[k], hdr[k][0], hdr[k][1], hdr[k][2])
        allmeta = self._all_gather_np(meta)  # [sender][receiver][4]
        send_sizes = [int(meta[r, 0]) for r in range(self.world)]
        recv_sizes = [int(allmeta[r, d.idSelf, 0]) for r in range(self.world)]
        nsend, nrecv = sum(send_sizes), sum(recv_sizes)
        if self._lbufs is None or self._lbufs[0].numel() < nsend or self._lbufs[1].numel() < nrecv:
            self._lbufs = (torch.empty(int(nsend * 1.25) + 256, dtype=torch.uint8, device=self.dev),
                           torch.empty(int(nrecv * 1.25) + 256, dtype=torch.uint8, device=self.dev))
        send, recv = self._lbufs[0][:nsend], self._lbufs[1][:nrecv]
        # the export buffer already holds the domains back to back in rank order (with alignment gaps): compact it
        pos = 0
        src = _DevView(ptr, int(offs[-1]), self.dev, torch)
        for k, r in enumerate(others):
            n = send_sizes[r]
            send[pos:pos + n].copy_(src.t[int(offs[k]):int(offs[k]) + n])
            pos += n
        dist.all_to_all_single(recv, send, output_split_sizes=recv_sizes, input_split_sizes=send_sizes)
        lap("all_to_all")
        pos = 0
        for r in range(self.world):
            if r != d.idSelf:
                d.pkd.pkdSetRemotePacked(r, allmeta[r, d.idSelf, 1:4], recv.data_ptr() + pos)
            pos += recv_sizes[r]
        lap("set_remote")
        self.timing = tm
        self.let_bytes = (nsend, nrecv)
        return nrecv


class LibExchange:
    """One Domain per rank with every collective BELOW the C ABI (csrc/gg_comm.cu): the small all-gathers of the top-tree
    assembly go through gg_comm_allgather, the tree exchange is gg_exchange (LET export -> NCCL send/receive or
    in-process peer copies -> ingest).  The rank's PKD must have a communicator (commInitNccl / commInitLocal)."""

    def __init__(self, domain: Domain):
        self.d = domain
        self.timing = {}
        self.let_bytes = (0, 0)
        self.stats = None
        self._summaries = None

    def top_tree(self):
        pk = self.d.pkd
        summaries = pk.commAllgather(self.d.summary())
        anc = pk.commAllgather(self.d.ancestor_moments(summaries))
        self.d.assemble(summaries, anc)
        self._summaries = summaries
        return summaries, anc

    def exchange(self, g, top: bool = True, upload: bool = True):
        """upload: gg_set_local of the host's tree + particles (a new step of a host-driven run); top: assemble the top
        tree again (after a tree build).  Then gg_set_top / gg_set_root_moments and the collective gg_exchange."""
        import time
        tm = {}
        t0 = time.perf_counter()

        def lap(name):
            nonlocal t0
            t1 = time.perf_counter()
            tm[name] = t1 - t0
            t0 = t1

        d = self.d
        if top or d.kdTop is None:
            self.top_tree()
            lap("top_tree")
        d.attach(upload=upload, announce=g)
        lap("attach_local" if upload else "set_top")
        st = d.pkd.pkdExchange(g, self._summaries[:, 0:6])
        lap("gg_exchange")
        tm.update(let_export=st["msExport"] * 1e-3, transfer=st["msTransfer"] * 1e-3, ingest=st["msIngest"] * 1e-3)
        self.timing, self.stats = tm, st
        self.let_bytes = (int(st["bytesSent"]), int(st["bytesReceived"]))
        return int(st["bytesReceived"])


def run_threads(domains: list, g, rounds: int = 1):
    """All ranks inside this process as THREADS with an in-process group (gg_group): every rank runs the same collective
    sequence a multi-process job runs -- top-tree all-gathers and gg_exchange below the ABI.  Returns the LibExchange
    drivers.  (ctypes releases the GIL inside the library, so the group's barriers make progress.)"""
    import threading
    grp = _pkd.Group(len(domains))
    drivers = [LibExchange(d) for d in domains]
    errs = []

    def work(k):
        try:
            d = domains[k]
            d.pkd.commInitLocal(grp, d.idSelf)
            for _ in range(rounds):
                drivers[k].exchange(g)
        except Exception as e:  # surfaced below; a failing rank would otherwise leave the others in a barrier
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(domains))]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    if errs:
        raise errs[0]
    if any(t.is_alive() for t in th):
        raise _pkd.GasolineB200Error("run_threads: a rank did not return (collective mismatch)")
    return drivers


class _DevView:
    """A torch uint8 view of a raw device allocation owned by the library (no copy, no ownership)."""

    def __init__(self, ptr: int, nbytes: int, dev: str, torch):
        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self.t = torch.as_tensor(a, device=dev)


def device_orb_share(p, rank: int, world: int, device: int, backend_device: str = "cuda", weights=None, timing=None,
                     collective: bool = False):
    """pstDomainDecomp across the processes of a torch.distributed job, per-rank work on the GPUs: this rank starts with
    a contiguous chunk of `p`, as the reference's ranks start with a contiguous range of the file (pstReadTipsy splits
    the file range down the rank tree, pst.c:676-725; the outcome does not depend on the initial distribution), answers the bisection's questions from its
    device (gg_orb_*), and receives the particles of its domain by one all-to-all.  Returns their indices in `p`,
    ascending.
    collective: the service context gets a library communicator (NCCL; the id travels by torch.distributed's object
    broadcast) and every level's bisection is ONE gg_orb_bisect_all -- the trial answers never visit the host."""
    import time as _time
    n = len(p.x)
    lo = rank * (n // world) + min(rank, n % world)
    hi = lo + n // world + (1 if rank < n % world else 0)
    mine = np.arange(lo, hi)
    t0 = _time.perf_counter()
    svc = PKD(device=device, fPeriod=p.period)
    svc.pkdOrbLoad(p.x[mine], p.y[mine], p.z[mine], fWeight=None if weights is None else np.asarray(weights)[mine])
    collective = bool(collective and world > 1)
    if collective:
        import torch
        import torch.distributed as dist
        ok = 1
        try:
            ident = [_pkd.comm_unique_id() if rank == 0 else None]
        except _pkd.GasolineB200Error:  # (no NCCL to load: every rank then takes the host path together)
            ident, ok = [None], 0
        dist.broadcast_object_list(ident, src=0)
        try:
            if ident[0] is None:
                ok = 0
            else:
                svc.commInitNccl(ident[0], rank, world)
        except _pkd.GasolineB200Error:
            ok = 0
        agreed = torch.tensor([ok], device=backend_device)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)  # all ranks take the same path
        collective = bool(agreed.item())
        if collective:
            svc.commAllgather(np.zeros(1))  # NCCL sets its channels up at the first collective: not part of a decomposition
    t1 = _time.perf_counter()
    if collective:
        nodes = pst_domain_decomp([svc], world, reduce=orb_reduce_lib(svc), collective_bisect=True)
    else:
        nodes = pst_domain_decomp([svc], world, reduce=orb_reduce_dist(backend_device))
    dest = leaf_rank(world)[svc.pkdOrbCells()]
    t2 = _time.perf_counter()
    svc.close()  # (with a library communicator this tears NCCL down: not part of a decomposition)
    got = orb_exchange(mine.astype(np.float64).reshape(-1, 1), dest, backend_device)
    t3 = _time.perf_counter()
    if timing is not None:
        timing.update(load_ms=(t1 - t0) * 1e3, decomp_ms=(t2 - t1) * 1e3, exchange_ms=(t3 - t2) * 1e3,
                      trials=int(sum(c["ittr"] for c in nodes)), collective=collective)
    return np.sort(got[:, 0].astype(np.int64))


def setup_rank(p, theta: float, rank: int, world: int, device: int | None, nBucket: int = 8, iOrder: int = 4,
               weights=None, backend_device: str | None = None, device_build: bool = False, device_orb: bool = False,
               comm: str = "lib", idx=None):
    """What one torch.distributed rank does before its first force evaluation (bench.py --gpus N, tests): every rank
    holds the same particle set `p`, takes ITS share of the ORB decomposition (the host's job in a Gasoline run,
    pstDomainDecomp pst.c:1854), builds its local tree and creates its GPU context.  Returns (pkd, exchange) where
    exchange() runs the top-tree assembly + the tree exchange (steps 1-5 of this module) and returns the bytes this
    rank received; with device=None no GPU context exists (host-only checks) and pkd is the host store.
    device_orb: the share comes from the reference's decomposition run on the devices (device_orb_share) instead of
    orb_decompose.  idx: this rank's particle indices when the caller has already decomposed.
    comm: "lib" (default on GPUs) -- the rank's context gets an NCCL communicator of its own (gg_comm_init; the id
    travels by torch.distributed's object broadcast, the host's control plane) and every data-plane collective runs
    below the C ABI (LibExchange); "torch" -- the bring-up path through torch.distributed tensors."""
    if idx is not None:
        idx = np.asarray(idx)  # the caller already knows this rank's share
    elif device_orb:
        idx = device_orb_share(p, rank, world, device, backend_device or "cuda", weights=weights)
    else:
        parts = orb_decompose(p.x, p.y, p.z, world, weights=weights)
        idx = parts[rank]
    d = Domain(rank, world, p.x[idx], p.y[idx], p.z[idx], p.m[idx], p.h[idx], p.period, theta, nBucket=nBucket,
               iOrder=iOrder, pinned=device is not None, device=device, device_build=device_build)
    d.global_index = idx[d.pkd.treeOrder if device_build else d.host.iOrderMap]  # tree position -> index in p
    attach = device is not None
    lib = attach and (backend_device or "cuda") != "cpu" and comm != "torch"
    if lib:
        # the library's own NCCL communicator: the id is created on rank 0 and handed round by the host's control plane
        import torch
        import torch.distributed as dist
        ident = [_pkd.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        d.pkd.commInitNccl(ident[0], rank, world)
        ex = LibExchange(d)
    else:
        ex = DistributedExchange(d, backend_device or ("cpu" if device is None else "cuda"))

    def exchange(top: bool = True, let=None, rebuild: bool = False, upload: bool = True):
        """let = GravityParams: send pruned locally-essential trees (gg_let_export) instead of whole domains.
        rebuild (device_build only): build the local tree again from the rank's particles first (a new step).
        upload=False (library exchange only): the local domain is already resident on the device."""
        if rebuild:
            d.rebuild()
            top = True
        if lib:
            return ex.exchange(let, top=top, upload=upload)
        if attach and ex.dev != "cpu":
            return ex.exchange_let(let, top=top) if let is not None else ex.exchange_packed(top=top)
        return ex.exchange(attach=attach)

    exchange.domain = d
    exchange.driver = ex
    return (d.pkd if d.pkd is not None else d.host), exchange
