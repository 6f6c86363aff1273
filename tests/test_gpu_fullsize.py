"""GPU parity at BASELINE.json's full sizes (configs[1] = Plummer 1 M, configs[2] = periodic 128^3 + Ewald,
configs[3] = periodic 256^3 = 16.8 M particles at theta = 0.5 on ONE GPU).

The oracle cannot evaluate millions of sinks in seconds, so the full-size run is checked three ways:
  * a random sample of COMPLETE sink buckets is re-evaluated by the oracle on the same tree (the oracle marks just
    those buckets' particles ACTIVE; their sink boxes -- hence their interaction lists -- are the full run's): counts
    bit-exact, accelerations / potentials within the north-star tolerance;
  * size-independent properties of the whole run: the per-bucket counts add up to the sums pkdGravAll reports
    (pkd.c:2945-2949), a second evaluation is bit-identical (deterministic summation order), total momentum change
    is small against the summed force magnitudes;
  * the reference's flop score recomputed from the counts equals dFlop (grav.c:246-247, ewald.c:175-176)."""
import numpy as np
import pytest

from gasoline_b200 import ics
from gasoline_b200.pkd import PKD, GravityParams
from oracle import oracle
from parity import MAX_TOL, RMS_TOL, acc_errors, pot_errors

pytestmark = pytest.mark.gpu

CASES = {
    "c2_plummer_1m": (lambda: ics.plummer(1_000_000), 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0), 300),
    "c3_periodic128_ewald": (lambda: ics.periodic_box(128), 0.7, GravityParams(nReps=1, bPeriodic=1, bEwald=1), 150),
    "c4_periodic256_theta05_ewald": (lambda: ics.periodic_box(256), 0.5, GravityParams(nReps=1, bPeriodic=1, bEwald=1), 100),
}


@pytest.mark.parametrize("name", list(CASES))
def test_full_size_sampled_buckets_and_invariants(name, gpu_lib):
    mk, theta, g, nSample = CASES[name]
    p = mk()
    pkd = PKD(fPeriod=p.period)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    pkd.pkdBuildBinary(8, theta, 4)
    out = pkd.pkdGravAll(g)
    counts = pkd.pkdBucketCounts()
    t = pkd.tree
    n = pkd.nLocal

    # ---- invariants of the whole run
    again = pkd.pkdGravAll(g)
    assert np.array_equal(out["acc"], again["acc"]) and np.array_equal(out["pot"], again["pot"])
    bk = np.where(t.iLower == -1)[0]
    nb = (t.pUpper[bk] - t.pLower[bk] + 1).astype(np.int64)
    c = counts[bk].astype(np.int64)
    assert np.all(c >= 0)
    assert out["nActive"] == n == int(nb.sum())
    assert out["dPartSum"] == float((nb * c[:, 0] + nb * (nb - 1) // 2).sum())
    assert out["dCellSum"] == float((nb * c[:, 2]).sum()) and out["dSoftSum"] == float((nb * c[:, 1]).sum())
    flop_tree = (nb * ((c[:, 0] + nb) * 38 + c[:, 1] * 82 + c[:, 2] * (35 + 277))).sum()
    assert out["dFlop"] - out["dFlopEwald"] == float(flop_tree)
    f = pkd.fMass[:, None] * out["acc"]
    assert np.linalg.norm(f.sum(axis=0)) <= 2e-3 * np.linalg.norm(f, axis=1).sum()

    # ---- a sample of complete buckets against the oracle on the same tree
    rng = np.random.default_rng(2026)
    pick = rng.choice(bk, size=nSample, replace=False)
    active = np.zeros(n, dtype=np.int32)
    for b in pick:
        active[t.pLower[b]:t.pUpper[b] + 1] = 1
    td = t.as_dict()
    td.update(x=pkd.x, y=pkd.y, z=pkd.z, m=pkd.fMass, h=pkd.fSoft, active=active, root=pkd.ilcnRoot,
              period=np.array(p.period), iOrder=pkd.iOrderMap)
    o = oracle.OracleGravity(None, tree=td)
    ref = o.gravity(g.nReps, g.bPeriodic, g.iOrder, g.bEwald, g.iEwOrder, g.fEwCut, g.fEwhCut)
    o.close()
    assert np.array_equal(counts[pick], ref["counts"][pick]), "per-bucket interaction-list counts differ"
    act = active.astype(bool)
    rms, mx = acc_errors(out["acc"][act], ref["acc"][act])
    prms, pmx = pot_errors(out["pot"][act], ref["pot"][act])
    print(f"{name}: {n} particles, {len(bk)} buckets, {int(act.sum())} sampled sinks: acc rms {rms:.3e} max {mx:.3e}; "
          f"pot rms {prms:.3e} max {pmx:.3e}; device {out['msTotal']:.2f} ms "
          f"(walk {out['msWalk']:.2f}, eval {out['msEval']:.2f}, Ewald {out['msEwald']:.2f})")
    assert rms <= RMS_TOL and mx <= MAX_TOL
    assert prms <= RMS_TOL and pmx <= MAX_TOL
    assert np.array_equal(out["fWeight"][act], ref["fWeight"][act])
    pkd.close()


@pytest.mark.parametrize("name", list(CASES))
def test_full_size_device_tree_build_identical(name, gpu_lib):
    """pkdBuildBinary on the device at BASELINE.json's sizes: every tree array and the particle order equal the host
    build's bit for bit (the host build is pinned to the reference's in the CPU suite)."""
    mk, theta, g, _ = CASES[name]
    p = mk()
    host = PKD(fPeriod=p.period)
    host.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    th = host.pkdBuildBinary(8, theta, 4)
    dev = PKD(fPeriod=p.period)
    dev.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    assert dev.pkdBuildBinaryDevice(8, theta) == th.nNodes
    dev.pkdBuildBinaryDevice(8, theta)  # second build on warm buffers: the timing below is steady state
    td, pd_ = dev.pkdFetchTree(with_mom=False)
    for k in ("bnd", "r", "fMass", "fSoft", "fOpen2", "pLower", "pUpper", "iLower", "iUpper"):
        assert np.array_equal(getattr(td, k), getattr(th, k)), f"{name}: tree field {k} differs from the host build"
    assert np.array_equal(dev.treeOrder, host.iOrderMap)
    assert np.array_equal(pd_["x"], host.x) and np.array_equal(pd_["z"], host.z)
    nn, nl, ms = dev.pkdBuildInfo()
    print(f"{name}: device tree build {nn} cells, {nl} levels, {ms:.2f} ms")
    host.close()
    dev.close()
