"""Development aid (uses the oracle's tree: lives under tests/): how much of k_eval's work comes from list entries that
ALL buckets of a walk group share?  Simulates k_walk's group traversal (GB consecutive sink buckets in tree order, the
reference's opening test per bucket, walk.h:12-30, walk.c:81) on the CPU and counts sink-cell and sink-particle
interactions by entry kind.  Open boundaries only.

    python tests/diag/group_share.py --n 100000 --gb 10
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gasoline_b200 import ics  # noqa: E402
from oracle import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--theta", type=float, default=0.7)
ap.add_argument("--gb", type=int, default=10)
ap.add_argument("--groups", type=int, default=400, help="walk groups sampled (evenly over the tree)")
a = ap.parse_args()

p = ics.plummer(a.n)
o = oracle.OracleGravity(p)
o.build_tree(8, a.theta, 4)
t = o.tree()
o.close()
iLower, iUpper, pLo, pUp = t["iLower"], t["iUpper"], t["pLower"], t["pUpper"]
r, fOpen2, bnd = t["r"], t["fOpen2"], t["bnd"]
nP = pUp - pLo + 1
buckets = []
c = t["iRoot"]
while c != -1 and c < t["nNodes"]:  # threaded pre-order (pkd.c:2909-3001)
    if iLower[c] == -1:
        buckets.append(c)
        c = iUpper[c]
    else:
        c = iLower[c]
    if c == 0 and len(buckets) > 1:
        break
buckets = np.array(buckets)
nG = (len(buckets) + a.gb - 1) // a.gb
sample = np.unique(np.linspace(0, nG - 1, min(a.groups, nG)).astype(int))
tot = dict(cell_shared=0, cell_masked=0, part_shared=0, part_masked=0, ent_cell_shared=0, ent_cell_masked=0)
for g in sample:
    bk = buckets[g * a.gb:(g + 1) * a.gb]
    nb = len(bk)
    full = (1 << nb) - 1
    sinks = nP[bk].astype(np.int64)
    lo, hi = bnd[bk, :3], bnd[bk, 3:]
    stack = [(t["iRoot"], full)]
    while stack:
        c, mask = stack.pop()
        d = np.maximum(np.maximum(lo - r[c], r[c] - hi), 0.0)
        opened = (np.sum(d * d, axis=1) <= fOpen2[c]) | (nP[c] < 4)
        om = 0
        for b in range(nb):
            if opened[b] and (mask >> b) & 1:
                om |= 1 << b
        acc = mask & ~om
        if acc:
            w = int(sum(sinks[b] for b in range(nb) if (acc >> b) & 1))
            if acc == full:
                tot["cell_shared"] += w; tot["ent_cell_shared"] += 1
            else:
                tot["cell_masked"] += w; tot["ent_cell_masked"] += 1
        if om:
            if iLower[c] == -1:
                w = int(sum(sinks[b] for b in range(nb) if (om >> b) & 1)) * int(nP[c])
                tot["part_shared" if om == full else "part_masked"] += w
            else:
                c1 = iLower[c]
                stack.append((iUpper[c1], om))
                stack.append((c1, om))
cells = tot["cell_shared"] + tot["cell_masked"]
parts = tot["part_shared"] + tot["part_masked"]
print(f"{p.name} theta {a.theta}, {len(buckets)} buckets, groups of {a.gb}, {len(sample)} groups sampled")
print(f"sink-cell interactions: {cells}  shared by the whole group {tot['cell_shared'] / cells:.3f}  "
      f"(list entries: shared {tot['ent_cell_shared']}, masked {tot['ent_cell_masked']})")
print(f"sink-particle interactions: {parts}  shared {tot['part_shared'] / max(parts, 1):.3f}")
print(f"cell share of all interactions: {cells / (cells + parts):.3f}")
