"""Generate tests/golden/opentypes.npz from the reference's OWN compiled code (oracle/_ref/libgasref.so): trees built with
the opening criteria of pkdCalcOpen other than OPEN_JOSH (pkd.c:2228-2264; opentype.h:5-9) and the forces those trees give.

    make -C oracle ref && python tests/golden/make_golden_opentypes.py

OPEN_ABSPAR (the distance at which the truncation-error estimate of the expansion falls to dAbsPartial, dRootBracket
pkd.c:2182-2224) does not terminate in the reference on a cell without extent -- a one-particle bucket has Bmax = B_k = 0
and the estimate is 0/0 (observed: Plummer 500, nBucket 8, 8 one-particle buckets: no return).  Its cases are therefore
particle sets whose trees hold no such cell (found by building with OPEN_JOSH first, same geometry).  The other three
criteria take the reference's "minimal" radius Bmax."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from gasoline_b200 import ics  # noqa: E402

OPEN_JOSH, OPEN_ABSPAR, OPEN_RELPAR, OPEN_ABSTOT, OPEN_RELTOT = 1, 2, 3, 4, 5

# name -> (generator, args, nBucket, iOpenType, dCrit, iOrder, gravity kwargs)
CASES = {
    "abspar_hex_plummer": ("plummer", dict(N=400), 16, OPEN_ABSPAR, 1e-3, 4, dict(nReps=0, bPeriodic=0, bEwald=0)),
    "abspar_oct_plummer": ("plummer", dict(N=300), 12, OPEN_ABSPAR, 1e-2, 3, dict(nReps=0, bPeriodic=0, bEwald=0)),
    "abspar_quad_plummer": ("plummer", dict(N=250), 8, OPEN_ABSPAR, 3e-2, 2, dict(nReps=0, bPeriodic=0, bEwald=0)),
    "abspar_mono_plummer": ("plummer", dict(N=200), 8, OPEN_ABSPAR, 1e-1, 1, dict(nReps=0, bPeriodic=0, bEwald=0)),
    "abspar_hex_periodic": ("periodic_box", dict(n=6, mode="jitter"), 8, OPEN_ABSPAR, 1e-4, 4,
                            dict(nReps=1, bPeriodic=1, bEwald=1)),
    "relpar_plummer": ("plummer", dict(N=1200, seed=2), 8, OPEN_RELPAR, 0.7, 4, dict(nReps=0, bPeriodic=0, bEwald=0)),
    "abstot_periodic": ("periodic_box", dict(n=8, mode="jitter"), 8, OPEN_ABSTOT, 0.7, 4,
                        dict(nReps=1, bPeriodic=1, bEwald=1)),
    "reltot_plummer": ("plummer", dict(N=900, seed=4), 5, OPEN_RELTOT, 0.7, 3, dict(nReps=0, bPeriodic=0, bEwald=0)),
}


def particles(name):
    """The case's particles.  OPEN_ABSPAR cases: the first seed whose tree has no cell without extent (see above)."""
    gen, args, nBucket, iOpenType, dCrit, iOrder, kw = CASES[name]
    if iOpenType != OPEN_ABSPAR:
        return getattr(ics, gen)(**args), None
    z = np.load(os.path.join(HERE, "opentypes.npz")) if os.path.exists(os.path.join(HERE, "opentypes.npz")) else None
    if z is not None and f"{name}_seed" in z:
        return getattr(ics, gen)(**dict(args, seed=int(z[f"{name}_seed"]))), int(z[f"{name}_seed"])
    return None, None


def main():
    from oracle import reflib
    assert reflib.available(), "build oracle/_ref first (make -C oracle ref)"
    L = reflib.lib()
    L.ref_build_tree_open.restype = C.c_double
    L.ref_build_tree_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
    out = {}
    for name, (gen, args, nBucket, iOpenType, dCrit, iOrder, kw) in CASES.items():
        seed = None
        if iOpenType == OPEN_ABSPAR:
            for seed in range(1, 400):
                p = getattr(ics, gen)(**dict(args, seed=seed))
                r = reflib.RefGravity(p)
                r.build_tree(nBucket, 0.7, iOrder)
                t = r.tree()
                r.close()
                if t["bmom"][:, 0].min() > 0:
                    break
            else:
                raise SystemExit(f"{name}: no seed gives a tree without zero-extent cells")
            out[f"{name}_seed"] = seed
        else:
            p = getattr(ics, gen)(**args)
        r = reflib.RefGravity(p)
        L.ref_build_tree_open(r.h, nBucket, iOpenType, dCrit, iOrder)
        t = r.tree()
        res = r.gravity(kw["nReps"], kw["bPeriodic"], iOrder, kw["bEwald"], iOrder)
        r.close()
        for k in ("fOpen2", "bmom", "mom", "r", "pLower", "pUpper", "iLower", "iUpper", "iOrder"):
            out[f"{name}_tree_{k}"] = t[k]
        out[f"{name}_counts"] = res["counts"]
        out[f"{name}_acc"], out[f"{name}_pot"] = res["acc"], res["pot"]
        out[f"{name}_sums"] = np.array([res["nActive"], res["dPartSum"], res["dCellSum"], res["dSoftSum"], res["dFlop"]])
        print(name, "seed", seed, p.n, "particles", t["nNodes"], "nodes; sqrt(fOpen2)/Bmax of the root",
              np.sqrt(t["fOpen2"][0]) / t["bmom"][0, 0], "interactions", res["dPartSum"] + res["dCellSum"] + res["dSoftSum"])
    np.savez_compressed(os.path.join(HERE, "opentypes.npz"), **out)
    print(os.path.getsize(os.path.join(HERE, "opentypes.npz")), "bytes")


if __name__ == "__main__":
    main()
