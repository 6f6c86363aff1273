"""Where the end-to-end time of one force evaluation goes (development aid; bench.py is the contract).

Times, with host buffers pinned: gg_set_local alone, gg_gravity device-resident, gg_gravity with the results delivered
to host arrays (zero-copy when pinned, staged copy when pageable), and the full upload + gravity step."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gasoline_b200 import ics, build
from gasoline_b200.pkd import PKD, GravityParams, pinned_empty

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="plummer")
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--theta", type=float, default=0.7)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--device-moments", type=int, default=0)
a = ap.parse_args()
build.build()
if a.workload == "plummer":
    p = ics.plummer(a.n); g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
else:
    p = ics.periodic_box(a.n); g = GravityParams(nReps=1, bPeriodic=1, bEwald=1)
pkd = PKD(fPeriod=p.period, pinned=True, device_moments=bool(a.device_moments))
pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
pkd.pkdBuildBinary(8, a.theta, 4)
n = pkd.nLocal
pin = [pinned_empty((n, 3)), pinned_empty(n), pinned_empty(n), pinned_empty(n)]
pag = [np.zeros((n, 3)), np.zeros(n), np.zeros(n), np.zeros(n)]


def timed(f, reps=a.reps):
    f(); f()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); f(); ts.append((time.perf_counter() - t) * 1e3)
    return min(ts), sum(ts) / len(ts)


print(f"{p.name}: {n} particles, {pkd.tree.nNodes} nodes, upload {pkd.upload_bytes()/1e6:.1f} MB")
print("upload (gg_set_local)            min %.3f avg %.3f ms" % timed(pkd.upload))
print("gravity, results stay on device  min %.3f avg %.3f ms" % timed(lambda: pkd.pkdGravAll(g, download=False)),
      " device msTotal %.3f" % pkd.stats["msTotal"])
print("gravity -> pinned host arrays    min %.3f avg %.3f ms" % timed(lambda: pkd.pkdGravAll(g, *pin, accumulate=False)),
      " device msTotal %.3f" % pkd.stats["msTotal"])
print("gravity -> pageable host arrays  min %.3f avg %.3f ms" % timed(lambda: pkd.pkdGravAll(g, *pag, accumulate=False)))
print("gravity += pageable (reference)  min %.3f avg %.3f ms" % timed(lambda: pkd.pkdGravAll(g, *pag, accumulate=True)))


def step():
    pkd.upload()
    pkd.pkdGravAll(g, *pin, accumulate=False)


print("upload + gravity -> pinned       min %.3f avg %.3f ms" % timed(step), " device msTotal %.3f walk %.3f eval %.3f" %
      (pkd.stats["msTotal"], pkd.stats["msWalk"], pkd.stats["msEval"]))
ref = pkd.pkdGravAll(g)  # fresh pageable arrays
for nm, u, v in (("acc", pin[0], ref["acc"]), ("pot", pin[1], ref["pot"]), ("dtGrav", pin[2], ref["dtGrav"]),
                 ("fWeight", pin[3], ref["fWeight"])):
    print(f"zero-copy vs staged {nm}: identical = {np.array_equal(u, v)}")
