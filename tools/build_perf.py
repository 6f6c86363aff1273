"""Device tree build timing (development aid): pkdBuildBinaryDevice repeated on one workload."""
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
ap.add_argument("--gravity", type=int, default=1)
a = ap.parse_args()
build.build()
if a.workload == "plummer":
    p = ics.plummer(a.n); g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
else:
    p = ics.periodic_box(a.n); g = GravityParams(nReps=1, bPeriodic=1, bEwald=1)
pkd = PKD(fPeriod=p.period, pinned=True)
pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
for i in range(a.reps):
    t = time.time()
    nn = pkd.pkdBuildBinaryDevice(8, a.theta)
    dt = time.time() - t
    _, nl, ms = pkd.pkdBuildInfo()
    line = f"{p.name} build {i}: {nn} cells, {nl} levels, device {ms:.3f} ms, wall (H2D + build + pack + moments enqueue) {dt*1e3:.2f} ms"
    if a.gravity:
        t = time.time()
        out = pkd.pkdGravAll(g, download=False)
        inter = out["dPartSum"] + out["dCellSum"] + out["dSoftSum"]
        line += (f" | gravity total {out['msTotal']:.3f} ms (walk {out['msWalk']:.2f} eval {out['msEval']:.2f} ewald {out['msEwald']:.2f})"
                 f" wall {(time.time()-t)*1e3:.2f} ms, {inter:.4g} interactions -> {inter/out['msTotal']*1e3:.4g}/s")
    print(line)
