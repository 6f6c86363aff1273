"""Generate tests/golden/multirank_*.npz from MULTI-RANK runs of the reference's own binary (oracle/_ref/gasoline_ref:
all reference objects + main.c, ranks = threads of the pthread MDL stand-in oracle/ref_shim/mdl.c).  The dump hooks
in oracle/ref_api.c (linker --wrap around pkdBucketWalk / pkdGravAll, env REF_DUMP) record, per rank: which particles
the reference's domain decomposition gave it (tree order), the top tree kdTop and ilcnRoot every rank holds, the
per-bucket interaction-list counts and the per-particle results.  Run in the build container:

    make -C oracle ref && python tests/golden/make_golden_multirank.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from gasoline_b200 import ics  # noqa: E402
from oracle import reflib  # noqa: E402

# name -> (generator, args, theta, nThreads)
CASES = {
    "multirank_periodic10_r2": ("periodic_box", dict(n=10), 0.7, 2),
    "multirank_periodic8_jitter_r3": ("periodic_box", dict(n=8, mode="jitter"), 0.7, 3),
    "multirank_plummer3000_r4": ("plummer", dict(N=3000), 0.7, 4),
}


def parse_dump(path):
    b = open(path, "rb").read()
    hdr = np.frombuffer(b, np.int32, 6)
    nThreads, idSelf, nLocal, nNodes, iRoot, nTop = (int(v) for v in hdr)
    off = 24
    iOrder = np.frombuffer(b, np.int32, nLocal, off).copy(); off += 4 * nLocal
    top_i = np.zeros((nTop, 2), np.int32); top_d = np.zeros((nTop, 37))
    for i in range(nTop):
        top_i[i] = np.frombuffer(b, np.int32, 2, off); off += 8
        top_d[i] = np.frombuffer(b, np.float64, 37, off); off += 37 * 8
    top_i[0] = 0  # cell 0 is never used (ROOT = 1); cells with pUpper == 0 are unused malloc garbage (pst.c:3928)
    top_d[top_i[:, 1] == 0] = 0.0
    top_i[top_i[:, 1] == 0, 0] = 0
    root = np.frombuffer(b, np.float64, 35, off).copy(); off += 35 * 8
    recs = []
    while True:
        r = np.frombuffer(b, np.int32, 6, off); off += 24
        if r[0] == -1:
            break
        recs.append(r.copy())
    sums = np.frombuffer(b, np.float64, 4, off).copy(); off += 32
    res = np.frombuffer(b, np.float64, 6 * nLocal, off).reshape(nLocal, 6).copy(); off += 48 * nLocal
    pos = None
    if len(b) >= off + 24 * nLocal:  # REF_DUMP_ALL: the particles' positions at this force evaluation
        pos = np.frombuffer(b, np.float64, 3 * nLocal, off).reshape(nLocal, 3).copy()
    return dict(pos=pos, nThreads=nThreads, idSelf=idSelf, nNodes=nNodes, iRoot=iRoot, iOrder=iOrder, top_i=top_i, top_d=top_d,
                root=root, buckets=np.array(recs, np.int32), sums=sums, res=res)


def main():
    assert os.path.exists(reflib.BIN_PATH), "build oracle/_ref first (make -C oracle ref)"
    for name, (gen, args, theta, nThreads) in CASES.items():
        p = getattr(ics, gen)(**args)
        with tempfile.TemporaryDirectory() as tmp:
            ics.write_tipsy_native(os.path.join(tmp, "ic.tipsy"), p)
            periodic = 1 if p.periodic else 0
            open(os.path.join(tmp, "run.param"), "w").write(
                f"achInFile = {tmp}/ic.tipsy\nachOutName = {tmp}/out\nbPeriodic = {periodic}\ndPeriod = 1\n"
                f"nReplicas = {periodic}\nbEwald = {periodic}\ndTheta = {theta}\nnSteps = 0\nbVStep = 1\n"
                "bDoDensity = 0\niBinaryOutput = 0\nbParaRead = 0\nbParaWrite = 0\nbOverwrite = 1\n")
            env = dict(os.environ, MDL_NTHREADS=str(nThreads), REF_DUMP=os.path.join(tmp, "dump"))
            # the dumps are complete once every rank's pkdGravAll has returned; what the binary does afterwards
            # (re-ordering particles for the output files, which the MDL stand-in's mdlSwap only supports when the
            # receiver has room for everything) is of no interest here, so its exit status is not checked
            subprocess.run([reflib.BIN_PATH, "run.param"], cwd=tmp, env=env, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=600)
            ranks = [parse_dump(os.path.join(tmp, f"dump.rank{r}")) for r in range(nThreads)]
        out = dict(nThreads=nThreads, theta=theta)
        for r, d in enumerate(ranks):
            assert d["idSelf"] == r and d["nThreads"] == nThreads
            for k in ("iOrder", "buckets", "sums", "res"):
                out[f"r{r}_{k}"] = d[k]
            out[f"r{r}_nNodes"], out[f"r{r}_iRoot"] = d["nNodes"], d["iRoot"]
            assert np.array_equal(d["top_i"], ranks[0]["top_i"]) and np.array_equal(d["top_d"], ranks[0]["top_d"])
            assert np.array_equal(d["root"], ranks[0]["root"])
        out["top_i"], out["top_d"], out["root"] = ranks[0]["top_i"], ranks[0]["top_d"], ranks[0]["root"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, p.n, "particles", nThreads, "ranks", [len(d["iOrder"]) for d in ranks],
              os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()
