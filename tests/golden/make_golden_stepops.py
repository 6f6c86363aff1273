"""Golden vectors for kick / drift / grav-step from the COMPILED REFERENCE (oracle/_ref/libgasref.so, i.e. the
reference's own pkdKick pkd.c:3780, pkdDrift pkd.c:3686, pkdGravStep pkd.c:4609).  Run in the container that has
/root/reference:   python tests/golden/make_golden_stepops.py   -> tests/golden/stepops.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import reflib  # noqa: E402
from oracle.oracle import DRIFT, GRAVSTEP, KICK  # noqa: E402


def inputs(seed=77, n=4096):
    rng = np.random.default_rng(seed)
    r = rng.uniform(-0.5, 0.5, size=(n, 3))
    r[::97] = np.nextafter(0.5, 0.0)  # particles about to leave through the upper face
    r[1::97] = -0.5
    v = rng.normal(0, 0.3, size=(n, 3))
    a = rng.normal(0, 5.0, size=(n, 3))
    active = (rng.random(n) < 0.6).astype(np.int32)
    dtGrav = rng.uniform(1.0, 1e4, size=n)
    dt = np.full(n, 0.05)
    return r, v, a, active, dtGrav, dt


PARAMS = dict(dvFacOne=0.98, dvFacTwo=0.0123, dDelta=0.0371, fCenter=(0.0, 0.0, 0.0), bPeriodic=1, fPeriod=(1.0, 1.0, 1.0),
              dEta=0.2)

if __name__ == "__main__":
    r, v, a, active, dtGrav, dt = inputs()
    out = {}
    for name, what in (("kick", KICK), ("drift", DRIFT), ("gravstep", GRAVSTEP), ("all", KICK | DRIFT | GRAVSTEP)):
        r2, v2, dt2, _ = reflib.ref_step_ops(r, v, a, active, dtGrav, dt, what=what, **PARAMS)
        out[name + "_r"], out[name + "_v"], out[name + "_dt"] = r2, v2, dt2
    r2, v2, dt2, _ = reflib.ref_step_ops(r * 3.0, v, a, None, dtGrav, dt, what=KICK | DRIFT, **dict(PARAMS, bPeriodic=0))
    out["open_r"], out["open_v"] = r2, v2
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stepops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
