"""Parity + timing of a kernel variant library (development aid; GASOLINE_B200_LIB selects the .so):
small cases against the oracle (list counts bit-exact, forces within tolerance), then the 1 M Plummer timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gasoline_b200 import ics
from gasoline_b200.pkd import PKD, GravityParams
from oracle import oracle

tag = os.environ.get("GASOLINE_B200_LIB", "product")
for p, theta, active_frac in ((ics.plummer(6000, seed=3), 0.7, 1.0), (ics.plummer(9000, seed=4), 0.5, 0.6)):
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    act = None if active_frac == 1.0 else (np.random.default_rng(1).uniform(0, 1, p.n) < active_frac).astype(np.int32)
    o = oracle.OracleGravity(p, active=act); o.build_tree(8, theta, 4); t = o.tree()
    ref = o.gravity(0, 0, 4, 0, 4, 2.6, 2.8); o.close()
    k = PKD(fPeriod=p.period); k.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, act); k.pkdBuildBinary(8, theta, 4)
    out = k.pkdGravAll(g)
    a = t["active"].astype(bool)
    ok_counts = np.array_equal(k.pkdBucketCounts(), ref["counts"])
    d = np.linalg.norm(out["acc"][a] - ref["acc"][a], axis=1) / np.linalg.norm(ref["acc"][a], axis=1)
    dp = np.abs(out["pot"][a] - ref["pot"][a]) / np.abs(ref["pot"][a])
    dt = np.abs(out["dtGrav"][a] - ref["dtGrav"][a]) / ref["dtGrav"][a]
    print(f"[{tag}] {p.name} active {active_frac}: counts {'OK' if ok_counts else 'DIFFER'}, acc rms {np.sqrt(np.mean(d*d)):.2e} "
          f"max {d.max():.2e}, pot max {dp.max():.2e}, dtGrav max {dt.max():.2e}", flush=True)
    k.close()
p = ics.plummer(1000000)
k = PKD(fPeriod=p.period); k.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h); k.pkdBuildBinary(8, 0.7, 4); k.upload()
g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
for i in range(4):
    out = k.pkdGravAll(g, download=False)
    print(f"[{tag}] 1M rep {i}: walk {out['msWalk']:.3f} eval {out['msEval']:.3f} total {out['msTotal']:.3f} ms", flush=True)
