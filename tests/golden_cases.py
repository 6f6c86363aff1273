"""Shared loader for tests/golden/*.npz (written by tests/golden/make_golden.py from the compiled reference)."""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES, make_active, make_particles  # noqa: E402

NAMES = sorted(CASES)


def load(name):
    gen, args, theta, kw, frac = CASES[name]
    p = make_particles(gen, args)
    active = make_active(p.n, frac)
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return p, active, theta, kw, z


def sorted_rows(a):
    """Interaction lists are sets: the reference's order depends on its walk order, ours on ours."""
    a = np.asarray(a)
    if a.size == 0:
        return a
    return a[np.lexsort(a.T[::-1])]
