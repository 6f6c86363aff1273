"""Loader for tests/golden/multirank_*.npz (multi-rank runs of the reference binary, make_golden_multirank.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_multirank import CASES  # noqa: E402
from gasoline_b200 import ics  # noqa: E402

NAMES = sorted(CASES)


def load(name):
    gen, args, theta, nThreads = CASES[name]
    p = getattr(ics, gen)(**args)
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return p, theta, nThreads, z


def make_domains(p, theta, nThreads, z, device=None):
    """One Domain per reference rank, holding exactly the particles the reference's decomposition gave that rank,
    in the reference's tree order (building a tree from tree-ordered particles reproduces the same tree)."""
    from gasoline_b200.domain import Domain
    doms = []
    for r in range(nThreads):
        io = z[f"r{r}_iOrder"]
        doms.append(Domain(r, nThreads, p.x[io], p.y[io], p.z[io], p.m[io], p.h[io], p.period, theta, device=device))
    return doms
