"""Print parity metrics of the CUDA path against the CPU oracle for a set of cases (development aid; lives under tests/ because it uses the oracle)."""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tests"))
import numpy as np
from gasoline_b200 import ics, build
from gasoline_b200.pkd import PKD, GravityParams
from oracle import oracle
from parity import acc_errors, pot_errors
build.build()
cases = [("zeld32", ics.periodic_box(32), 0.7, GravityParams(nReps=1, bPeriodic=1)),
         ("jitter16", ics.periodic_box(16, mode="jitter"), 0.7, GravityParams(nReps=1, bPeriodic=1)),
         ("jitter32", ics.periodic_box(32, mode="jitter"), 0.7, GravityParams(nReps=1, bPeriodic=1)),
         ("zeld32_noewald", ics.periodic_box(32), 0.7, GravityParams(nReps=1, bPeriodic=1, bEwald=0)),
         ("zeld64_th05", ics.periodic_box(64), 0.5, GravityParams(nReps=1, bPeriodic=1)),
         ("plummer100k", ics.plummer(100000), 0.7, GravityParams()),
         ]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
for name, p, theta, g in cases:
    o = oracle.OracleGravity(p); o.build_tree(8, theta, 4)
    ref = o.gravity(g.nReps, g.bPeriodic, g.iOrder, g.bEwald, g.iEwOrder, g.fEwCut, g.fEwhCut); o.close()
    pkd = PKD(fPeriod=p.period); pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h); pkd.pkdBuildBinary(8, theta, 4)
    out = pkd.pkdGravAll(g); counts = pkd.pkdBucketCounts()
    rms, mx = acc_errors(out["acc"], ref["acc"]); prms, pmx = pot_errors(out["pot"], ref["pot"])
    dt = (np.abs(out["dtGrav"] - ref["dtGrav"]) / ref["dtGrav"]).max()
    print(f"{name:16s} counts_equal={np.array_equal(counts, ref['counts'])} flop_equal={out['dFlop']==ref['dFlop']} "
          f"acc rms {rms:.2e} max {mx:.2e} | pot rms {prms:.2e} max {pmx:.2e} | dt max {dt:.2e} | "
          f"|a|rms {np.sqrt((ref['acc']**2).sum(1).mean()):.3g} phi rms {np.sqrt((ref['pot']**2).mean()):.3g} "
          f"tree {out['msTree']:.2f}ms ewald {out['msEwald']:.2f}ms cpu-oracle {ref['seconds']:.2f}s", flush=True)
    pkd.close()
