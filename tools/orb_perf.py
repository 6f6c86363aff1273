"""ORB decomposition timing (development aid): pst_domain_decomp on one device context, repeated."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gasoline_b200 import ics, build, domain
from gasoline_b200.pkd import PKD

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--ranks", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--weights", type=int, default=0)
a = ap.parse_args()
build.build()
p = ics.plummer(a.n)
w = np.random.default_rng(1).uniform(0.5, 2.0, p.n) if a.weights else None
pkd = PKD(fPeriod=p.period)
for i in range(a.reps):
    t = time.perf_counter()
    pkd.pkdOrbLoad(p.x, p.y, p.z, fWeight=w)
    t1 = time.perf_counter()
    nodes = domain.pst_domain_decomp([pkd], a.ranks)
    cells = pkd.pkdOrbCells()
    t2 = time.perf_counter()
    print(f"{p.name} -> {a.ranks} domains, rep {i}: load {1e3*(t1-t):.2f} ms, decomposition {1e3*(t2-t1):.2f} ms "
          f"({sum(n['ittr'] for n in nodes)} trials), sizes {np.bincount(domain.leaf_rank(a.ranks)[cells]).tolist()}")
pkd.close()
