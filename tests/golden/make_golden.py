"""Generate tests/golden/*.npz from the reference's OWN compiled code (oracle/_ref/libgasref.so, built by
oracle/Makefile from /root/reference).  Run in the build container (the GPU box has no /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

Each fixture holds the inputs' generator arguments and the reference's outputs for one small case: its tree
(pkdBuildBinary), per-bucket interaction-list counts (pkdBucketWalk), per-particle a / fPot / dtGrav / fWeight and
the scalar sums of pkdGravAll, the Ewald k-space table (pkdEwaldInit) and the full lists of a few buckets."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from gasoline_b200 import ics  # noqa: E402
from oracle import reflib  # noqa: E402

# name -> (generator, args, theta, gravity kwargs, active fraction or None, seed for the active mask)
CASES = {
    "periodic8_jitter_ewald": ("periodic_box", dict(n=8, mode="jitter"), 0.7, dict(nReps=1, bPeriodic=1, bEwald=1), None),
    "periodic10_zeldovich_ewald_theta05": ("periodic_box", dict(n=10), 0.5, dict(nReps=1, bPeriodic=1, bEwald=1), None),
    "periodic8_noewald_order2": ("periodic_box", dict(n=8, seed=5), 0.7,
                                 dict(nReps=1, bPeriodic=1, bEwald=0, iOrder=2), None),
    "plummer3000": ("plummer", dict(N=3000), 0.7, dict(nReps=0, bPeriodic=0, bEwald=0), None),
    "plummer2000_active30": ("plummer", dict(N=2000, seed=7), 0.7, dict(nReps=0, bPeriodic=0, bEwald=0), 0.3),
    "plummer1500_bigsoft": ("plummer", dict(N=1500, seed=5, eps=0.4), 0.7, dict(nReps=0, bPeriodic=0, bEwald=0), None),
}


def make_particles(gen, args):
    return getattr(ics, gen)(**args)


def make_active(n, frac, seed=3):
    if frac is None:
        return None
    return (np.random.default_rng(seed).random(n) < frac).astype(np.int32)


def main():
    assert reflib.available(), "build oracle/_ref first (make -C oracle ref)"
    for name, (gen, args, theta, kw, frac) in CASES.items():
        p = make_particles(gen, args)
        active = make_active(p.n, frac)
        r = reflib.RefGravity(p, active=active)
        r.build_tree(8, theta, 4)
        t = r.tree()
        order = kw.get("iOrder", 4)
        res = r.gravity(kw["nReps"], kw["bPeriodic"], order, kw["bEwald"], order)
        out = {f"tree_{k}": t[k] for k in ("bnd", "r", "fMass", "fSoft", "fOpen2", "mom", "pLower", "pUpper", "iLower",
                                          "iUpper", "iOrder", "root")}
        out.update(nNodes=t["nNodes"], iRoot=t["iRoot"], counts=res["counts"], acc=res["acc"], pot=res["pot"],
                   dtGrav=res["dtGrav"], fWeight=res["fWeight"],
                   sums=np.array([res["nActive"], res["dPartSum"], res["dCellSum"], res["dSoftSum"], res["dFlop"]]))
        if kw["bPeriodic"]:
            out["ewt"] = r.ewald_table(2.8, order)
        bk = np.where(t["iLower"] == -1)[0]
        pick = bk[:: max(1, len(bk) // 3)][:3]
        out["list_buckets"] = pick.astype(np.int32)
        for i, b in enumerate(pick):
            if active is not None and not t["active"][t["pLower"][b]:t["pUpper"][b] + 1].any():
                continue
            ilp, ilcs, ilcn = r.bucket_lists(int(b), kw["nReps"], order)
            out[f"ilp{i}"], out[f"ilcs{i}"], out[f"ilcn{i}"] = ilp, ilcs, ilcn
        r.close()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, p.n, "particles", t["nNodes"], "nodes", os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()
