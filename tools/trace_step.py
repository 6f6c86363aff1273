"""GG_TRACE=1 python tools/trace_step.py [--device-moments 1]: host-side phase timings of upload + gravity (stderr)."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gasoline_b200 import ics, build
from gasoline_b200.pkd import PKD, GravityParams, pinned_empty
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--device-moments", type=int, default=0)
a = ap.parse_args()
build.build()
p = ics.plummer(a.n); g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
pkd = PKD(fPeriod=p.period, pinned=True, device_moments=bool(a.device_moments))
pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
pkd.pkdBuildBinary(8, 0.7, 4)
n = pkd.nLocal
pin = [pinned_empty((n, 3)), pinned_empty(n), pinned_empty(n), pinned_empty(n)]
os.environ.pop("GG_TRACE", None)
for i in range(4):
    if i == 3:
        os.environ["GG_TRACE"] = "1"
        sys.stderr.write(f"---- traced step (device_moments={a.device_moments})\n")
    t0 = time.perf_counter()
    pkd.upload()
    t1 = time.perf_counter()
    pkd.pkdGravAll(g, *pin, accumulate=False)
    t2 = time.perf_counter()
sys.stderr.write(f"python wall: upload {(t1-t0)*1e3:.3f} ms, gravity {(t2-t1)*1e3:.3f} ms; device msTotal {pkd.stats['msTotal']:.3f}\n")
