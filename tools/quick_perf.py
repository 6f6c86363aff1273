"""Quick device-side timing of one workload (development aid; bench.py is the contract)."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gasoline_b200 import ics, build
from gasoline_b200.pkd import PKD, GravityParams

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="plummer")
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--theta", type=float, default=0.7)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--order", type=int, default=4)
a = ap.parse_args()
build.build()
t0 = time.time()
if a.workload == "plummer":
    p = ics.plummer(a.n); g = GravityParams(nReps=0, bPeriodic=0, bEwald=0, iOrder=a.order)
else:
    p = ics.periodic_box(a.n); g = GravityParams(nReps=1, bPeriodic=1, bEwald=1, iOrder=a.order, iEwOrder=a.order)
t1 = time.time()
pkd = PKD(fPeriod=p.period)
pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
pkd.pkdBuildBinary(8, a.theta, 4)
t2 = time.time()
pkd.upload()
t3 = time.time()
print(f"{p.name}: ic {t1-t0:.2f}s tree {t2-t1:.2f}s upload {t3-t2:.3f}s nodes {pkd.tree.nNodes}")
for i in range(a.reps):
    t = time.time()
    out = pkd.pkdGravAll(g, download=False)
    dt = time.time() - t
    inter = out["dPartSum"] + out["dCellSum"] + out["dSoftSum"]
    print(f"  rep {i}: walk {out["msWalk"]:.3f} eval {out["msEval"]:.3f} tree {out["msTree"]:.3f} ms ewald {out['msEwald']:.3f} ms total {out['msTotal']:.3f} ms wall {dt*1e3:.2f} ms | "
          f"inter {inter:.4g} -> {inter/out['msTotal']*1e3:.4g} int/s, flop {out['dFlop']:.4g} -> {out['dFlop']/out['msTotal']*1e-9:.1f} TFLOP/s(ref-scored), "
          f"maxlists {out['nMaxPart']}/{out['nMaxCellSoft']}/{out['nMaxCellNewt']}")
