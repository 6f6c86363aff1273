"""Golden vectors for time-step selection and rungs from the COMPILED REFERENCE (oracle/_ref/libgasref.so: the
reference's own pkdInitDt pkd.c:4818, pkdAccelStep pkd.c:4625, pkdGravStep pkd.c:4609, pkdDtToRung pkd.c:4715,
pkdActiveRung pkd.c:4569).  Run where /root/reference exists:  python tests/golden/make_golden_rungops.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import reflib  # noqa: E402
from oracle.oracle import ACCELSTEP, ACTIVERUNG, DTTORUNG, GRAVSTEP_R, INITDT  # noqa: E402


def inputs(seed=91, n=3000):
    rng = np.random.default_rng(seed)
    v = rng.normal(0, 1, (n, 3))
    a = rng.normal(0, 30, (n, 3))
    a[::50] = 0.0  # zero acceleration: no criterion applies (acc > 0 tests)
    pot = -rng.uniform(0.1, 5, n)
    h = rng.uniform(0.001, 0.02, n)
    dtg = rng.uniform(1, 1e5, n)
    rung = rng.integers(0, 3, n).astype(np.int32)
    active = (rung >= 1).astype(np.int32)  # pkdDtToRung asserts rung >= iRung => ACTIVE
    dt = np.full(n, 0.02)
    dt[::7] = 0.02 / 4  # exact sub-multiples of dDelta: the "integer boundary goes to the lower rung" branch
    return v, a, pot, h, dtg, active, dt, rung


CASES = {
    "accel_all": (INITDT | ACCELSTEP | DTTORUNG, dict(iRung=1, dDelta=0.02)),
    "everything": (INITDT | ACCELSTEP | GRAVSTEP_R | DTTORUNG | ACTIVERUNG, dict(iRung=1, bSqrtPhi=1, iRungActive=2, dDelta=0.02)),
    "symplectic": (DTTORUNG, dict(iRung=1, bAll=0, dDelta=0.015)),
    "clamped": (GRAVSTEP_R | DTTORUNG, dict(iRung=1, iMaxRung=3, dDelta=0.02)),
    "exact_rung": (DTTORUNG | ACTIVERUNG, dict(iRung=1, dDelta=0.02, iRungActive=3, bGreater=0)),
}

if __name__ == "__main__":
    v, a, pot, h, dtg, active, dt, rung = inputs()
    out = {}
    for name, (what, kw) in CASES.items():
        act2, dt2, rung2, o = reflib.ref_rung_ops(v, a, pot, h, dtg, active, dt, rung, what=what, **kw)
        out[name + "_active"], out[name + "_dt"], out[name + "_rung"], out[name + "_out"] = act2, dt2, rung2, o
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rungops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
