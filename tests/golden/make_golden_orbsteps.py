"""Generate tests/golden/orbsteps_*.npz: the domain decompositions of a TIME-STEPPING multi-rank run of the reference
binary (oracle/_ref/gasoline_ref on pthread-MDL ranks, REF_DUMP_ALL=1: one dump per force evaluation).  Step 0 is the
first decomposition of a run (pst->iSplitDim == -1, unit weights); steps 1.. are LATER decompositions: positions after
kick + drift, work weights = the fWeight the previous force evaluation left (pkd.c:2851-2861), split axis chosen with the
NEWSPLITDIMCUT hysteresis against the previous axis (pst.c:1900-1910), bDoRootFind = bDoSplitDimFind = 1 (single rung,
everything active: master.c:4176-4177).  Per step the fixture holds, indexed by iOrder: the positions, the rank each
particle went to, and the fWeight after the step's force evaluation.  Run in the build container:

    make -C oracle ref && python tests/golden/make_golden_orbsteps.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from gasoline_b200 import ics  # noqa: E402
from make_golden_multirank import parse_dump  # noqa: E402
from oracle import reflib  # noqa: E402

# name -> (generator, args, theta, nThreads, nSteps, dDelta, dExtraStore)
# dExtraStore = 1: every rank's store has room for what the work-weighted splits send it, so the store-overflow branch of
# _pstRootSplit (pst.c:1049-1270) never moves a boundary.  The *_overflow cases run with small stores (the first with the
# default dExtraStore = 0.1): some of their decompositions send a side more particles than its ranks' stores hold and the
# reference bisects its second boundary (fSplitInactive) into the cell -- on 2, 3, 5 and 6 ranks, open and periodic, up to
# four cells of one decomposition.  The restatements reproduce them when given the ranks' stores (rank_stores).
CASES = {
    "orbsteps_plummer2000_r4": ("plummer", dict(N=2000, seed=5), 0.7, 4, 3, 0.02, 1.0),
    "orbsteps_plummer1500_r3": ("plummer", dict(N=1500, seed=8), 0.7, 3, 2, 0.05, 1.0),
    "orbsteps_periodic10_r2": ("periodic_box", dict(n=10), 0.7, 2, 2, 0.01, 1.0),
    "orbsteps_plummer1500_r3_overflow": ("plummer", dict(N=1500, seed=8), 0.7, 3, 2, 0.05, 0.1),
    "orbsteps_plummer3000_r5_overflow": ("plummer", dict(N=3000, seed=11), 0.7, 5, 3, 0.05, 0.05),
    "orbsteps_periodic12_r3_overflow": ("periodic_box", dict(n=12), 0.7, 3, 2, 0.01, 0.01),
    "orbsteps_plummer1800_r2_overflow": ("plummer", dict(N=1800, seed=2), 0.7, 2, 3, 0.05, 0.02),
    "orbsteps_plummer2500_r6_overflow": ("plummer", dict(N=2500, seed=9), 0.7, 6, 2, 0.05, 0.06),
}


def run_case(gen, args, theta, nThreads, nSteps, dDelta, dExtraStore):
    p = getattr(ics, gen)(**args)
    with tempfile.TemporaryDirectory() as tmp:
        ics.write_tipsy_native(os.path.join(tmp, "ic.tipsy"), p)
        periodic = 1 if p.periodic else 0
        open(os.path.join(tmp, "run.param"), "w").write(
            f"achInFile = {tmp}/ic.tipsy\nachOutName = {tmp}/out\nbPeriodic = {periodic}\ndPeriod = 1\n"
            f"nReplicas = {periodic}\nbEwald = {periodic}\ndTheta = {theta}\nnSteps = {nSteps}\ndDelta = {dDelta}\n"
            f"iOutInterval = {10 * nSteps}\niLogInterval = 1\nbVStep = 1\nbDoDensity = 0\niBinaryOutput = 0\n"
            f"bParaRead = 0\nbParaWrite = 0\nbOverwrite = 1\niCheckInterval = 0\ndExtraStore = {dExtraStore}\n")
        env = dict(os.environ, MDL_NTHREADS=str(nThreads), REF_DUMP=os.path.join(tmp, "dump"), REF_DUMP_ALL="1")
        r = subprocess.run([reflib.BIN_PATH, "run.param"], cwd=tmp, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "Integration complete" in r.stdout, r.stdout[-800:] + r.stderr[-800:]
        steps = []
        k = 0
        while os.path.exists(os.path.join(tmp, f"dump.s{k}.rank0")):
            steps.append([parse_dump(os.path.join(tmp, f"dump.s{k}.rank{q}")) for q in range(nThreads)])
            k += 1
    return p, steps


def main():
    assert os.path.exists(reflib.BIN_PATH), "build oracle/_ref first (make -C oracle ref)"
    only = sys.argv[1:]  # (names: regenerate just those)
    for name, (gen, args, theta, nThreads, nSteps, dDelta, dExtraStore) in CASES.items():
        if only and name not in only:
            continue
        p, steps = run_case(gen, args, theta, nThreads, nSteps, dDelta, dExtraStore)
        assert len(steps) == nSteps + 1, (name, len(steps))
        out = dict(nThreads=nThreads, theta=theta, nSteps=nSteps, dDelta=dDelta, dExtraStore=dExtraStore)
        for k, ranks in enumerate(steps):
            pos = np.zeros((p.n, 3)); rank = np.full(p.n, -1, np.int32); w = np.zeros(p.n)
            for q, d in enumerate(ranks):
                pos[d["iOrder"]] = d["pos"]
                rank[d["iOrder"]] = q
                w[d["iOrder"]] = d["res"][:, 5]
            assert (rank >= 0).all()
            out[f"s{k}_pos"], out[f"s{k}_rank"], out[f"s{k}_fWeight"] = pos, rank, w
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, p.n, "particles", nThreads, "ranks", len(steps), "decompositions:",
              [np.bincount(out[f"s{k}_rank"]).tolist() for k in range(len(steps))])


if __name__ == "__main__":
    main()
