"""Time the reference's CPU implementation of the gravity path on the host cores.  TEST/BENCH INFRASTRUCTURE ONLY
(used by bench.py's cpu_baseline leg and by `bench.py --impl reference`; never imported by the product).

What runs: the reference's OWN compiled pkdGravAll (pkd.c:2868: pkdBucketWalk + pkdBucketInteract + pkdBucketEwald)
from oracle/_ref/libgasref.so when that library exists ("kind": "reference"), else our C restatement
oracle/liboracle.so ("kind": "port").  The image has no MPI and the reference parallelises only through MDL ranks,
so all host cores are used like this: the parent builds the reference tree once (one rank, msrBuildTree sequence),
then fork()s P workers that share it copy-on-write; worker k marks a disjoint 1/P share of the sink buckets ACTIVE
(the reference's own partial-active mechanism, pkd.c:2916-2944) and calls pstGravity -> pkdGravAll.  Sources are
always the full tree, so every worker computes exactly the forces the 1-rank run computes for its sinks.  Time =
slowest worker (like max over MPI ranks); tree build and IC generation are outside the timed region, as they are
for the GPU arm.

A bounded SAMPLE of the sink buckets (every k-th run of 64 consecutive buckets in tree order, shared round-robin
between workers) keeps the run within seconds; throughput = interactions of the sampled sinks / time.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from gasoline_b200 import ics  # noqa: E402  (input generation only)
from oracle import oracle, reflib  # noqa: E402

CHUNK = 64  # consecutive buckets per work unit

_G = {}


def make_workload(spec: str):
    """'plummer:N:theta' | 'periodic:n:theta' (n^3 particles, Ewald on) -> (Particles, theta, gravity kwargs)."""
    kind, n, theta = spec.split(":")
    n, theta = int(n), float(theta)
    if kind == "plummer":
        return ics.plummer(n), theta, dict(nReps=0, bPeriodic=0, bEwald=0)
    if kind == "periodic":
        return ics.periodic_box(n), theta, dict(nReps=1, bPeriodic=1, bEwald=1)
    raise ValueError(spec)


def _worker(k):
    eng, t, units, P, kw = _G["eng"], _G["tree"], _G["units"], _G["P"], _G["kw"]
    act = np.zeros(len(t["x"]), dtype=np.int32)
    for lo, hi in units[k::P]:
        act[lo:hi] = 1
    if act.sum() == 0:
        return 0.0, 0.0, 0, 0.0
    eng.set_active_tree(act)
    r = eng.gravity(kw["nReps"], kw["bPeriodic"], 4, kw["bEwald"], 4, **_G["extra"])
    return r["seconds"], r["dPartSum"] + r["dCellSum"] + r["dSoftSum"], r["nActive"], r["dFlop"]


def run(spec: str, target_seconds: float = 15.0, cores: int = 0, kind: str = "auto", repeats: int = 1):
    p, theta, kw = make_workload(spec)
    use_ref = reflib.available() if kind == "auto" else kind == "reference"
    P = cores or len(os.sched_getaffinity(0))
    t0 = time.time()
    if use_ref:
        eng = reflib.RefGravity(p)
    else:
        eng = oracle.OracleGravity(p)
    eng.build_tree(8, theta, 4)
    t = eng.tree()
    t_build = time.time() - t0
    bk = np.where(t["iLower"] == -1)[0]
    bk = bk[np.argsort(t["pLower"][bk])]
    lo, hi = t["pLower"][bk], t["pUpper"][bk] + 1
    chunks = [(int(lo[i]), int(hi[min(i + CHUNK, len(bk)) - 1])) for i in range(0, len(bk), CHUNK)]
    _G.update(eng=eng, tree=t, kw=kw, extra={} if use_ref else dict(threads=1))
    # calibration: one chunk from the middle on one core -> seconds per sink particle
    _G.update(units=[chunks[len(chunks) // 2]], P=1)
    s, inter, nact, _ = _worker(0)
    per_particle = max(s, 1e-4) / max(nact, 1)
    want_particles = target_seconds * P / per_particle
    stride = max(1, int(np.ceil(p.n / want_particles)))
    units = chunks[::stride]
    if len(units) < P:  # tiny problems: fewer workers than cores
        P = max(1, len(units))
    _G.update(units=units, P=P)
    best = None
    for _ in range(repeats):
        if P == 1:
            res = [_worker(0)]
        else:
            with mp.get_context("fork").Pool(P) as pool:
                res = pool.map(_worker, range(P))
        secs = max(r[0] for r in res)
        inter = sum(r[1] for r in res)
        val = inter / secs
        if best is None or val > best["value"]:
            best = dict(value=val, seconds=secs, interactions=inter, nActive=int(sum(r[2] for r in res)),
                        dFlop=sum(r[3] for r in res))
    best.update(unit="interactions/s", cores=P, kind="reference" if use_ref else "port",
                sample=f"{best['nActive']} of {p.n} sink particles (every {stride}th run of {CHUNK} buckets in tree order), "
                       f"{spec}, full source tree; slowest of {P} forked workers {best['seconds']:.2f} s; "
                       f"tree build {t_build:.1f} s outside the timed region",
                workload=spec)
    eng.close()
    return best


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="plummer:1000000:0.7")
    ap.add_argument("--seconds", type=float, default=15.0)
    ap.add_argument("--cores", type=int, default=0)
    ap.add_argument("--kind", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--repeats", type=int, default=1)
    a = ap.parse_args()
    print(json.dumps(run(a.workload, a.seconds, a.cores, a.kind, a.repeats)))
