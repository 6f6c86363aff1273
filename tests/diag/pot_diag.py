"""Potential-error diagnostic (development aid; lives under tests/ because it uses the oracle): signed error statistics of the CUDA path vs the oracle."""
import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
import numpy as np
from gasoline_b200 import ics
from gasoline_b200.pkd import PKD, GravityParams
from oracle import oracle
for name, p, g in [("jitter16", ics.periodic_box(16, mode="jitter"), GravityParams(nReps=1, bPeriodic=1)),
                   ("zeld32", ics.periodic_box(32), GravityParams(nReps=1, bPeriodic=1)),
                   ("jitter16_noewald", ics.periodic_box(16, mode="jitter"), GravityParams(nReps=1, bPeriodic=1, bEwald=0))]:
    o = oracle.OracleGravity(p); o.build_tree(8, 0.7, 4)
    ref = o.gravity(g.nReps, g.bPeriodic, 4, g.bEwald, 4); o.close()
    pkd = PKD(fPeriod=p.period); pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h); pkd.pkdBuildBinary(8, 0.7, 4)
    out = pkd.pkdGravAll(g); pkd.close()
    d = out["pot"] - ref["pot"]
    print(f"{name}: phi mean {ref['pot'].mean():.4e} rms {np.sqrt((ref['pot']**2).mean()):.4e} | dphi mean {d.mean():.3e} std {d.std():.3e} "
          f"| mean/|phi_tree| est {d.mean()/35:.2e}")
