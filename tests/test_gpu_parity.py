"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from gasoline_b200 import ics
from gasoline_b200.pkd import PKD, GravityParams, Tree
from oracle import oracle
from parity import MAX_TOL, RMS_TOL, acc_errors, pot_errors

pytestmark = pytest.mark.gpu

CASES = {
    "c1_periodic32_ewald": (lambda: ics.periodic_box(32), 0.7, GravityParams(nReps=1, bPeriodic=1, bEwald=1)),
    "periodic16_jitter": (lambda: ics.periodic_box(16, mode="jitter"), 0.7, GravityParams(nReps=1, bPeriodic=1)),
    # nReplicas = 2 (125 images, 7 image bits) and 3 (343 images, 9 bits; Ewald's hole follows nReps, ewald.c:60-70)
    "periodic10_nreps2_ewald": (lambda: ics.periodic_box(10), 0.7, GravityParams(nReps=2, bPeriodic=1, bEwald=1)),
    "periodic8_nreps3_ewald": (lambda: ics.periodic_box(8), 0.7, GravityParams(nReps=3, bPeriodic=1, bEwald=1)),
    "periodic8_nreps3_noewald": (lambda: ics.periodic_box(8, mode="jitter"), 0.6, GravityParams(nReps=3, bPeriodic=1, bEwald=0)),
    # nReplicas 4 (729 images, 10 bits) and 5 (1331 images, 11 bits): the walk seeds its frontier in batches of 343 roots
    "periodic6_nreps4_ewald": (lambda: ics.periodic_box(6, mode="jitter"), 0.7, GravityParams(nReps=4, bPeriodic=1, bEwald=1)),
    "periodic5_nreps5_noewald": (lambda: ics.periodic_box(5, seed=3), 0.6, GravityParams(nReps=5, bPeriodic=1, bEwald=0)),
    # lower multipole orders of the lists (QEVAL fallthrough, qeval.h:21-64) and of the Ewald root expansion (meval.h:21-81)
    "plummer8k_order1": (lambda: ics.plummer(8000, seed=6), 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0, iOrder=1)),
    "plummer8k_order3": (lambda: ics.plummer(8000, seed=6), 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0, iOrder=3)),
    "periodic10_order2_ewald2": (lambda: ics.periodic_box(10), 0.7, GravityParams(nReps=1, bPeriodic=1, bEwald=1, iOrder=2, iEwOrder=2)),
    "periodic10_order4_ewald3": (lambda: ics.periodic_box(10), 0.7, GravityParams(nReps=1, bPeriodic=1, bEwald=1, iOrder=4, iEwOrder=3)),
    "plummer20k": (lambda: ics.plummer(20000), 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "plummer50k_theta05": (lambda: ics.plummer(50000), 0.5, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
}


def run_oracle(p, theta, g, active=None):
    o = oracle.OracleGravity(p, active=active)
    o.build_tree(8, theta, 4)
    t = o.tree()
    res = o.gravity(g.nReps, g.bPeriodic, g.iOrder, g.bEwald, g.iEwOrder, g.fEwCut, g.fEwhCut)
    o.close()
    return t, res


def run_gpu(p, theta, g, active=None, tree=None):
    pkd = PKD(fPeriod=p.period)
    if tree is None:
        pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, active)
        pkd.pkdBuildBinary(8, theta, 4)
    else:
        t = tree
        pkd.pkdSetTree(Tree(t["nNodes"], t["iRoot"], **{k: t[k] for k in Tree.FIELDS}), t["x"], t["y"], t["z"],
                       t["m"], t["h"], active=t["active"], ilcnRoot=t["root"], iOrderMap=t["iOrder"])
    out = pkd.pkdGravAll(g)
    counts = pkd.pkdBucketCounts()
    return pkd, out, counts


@pytest.mark.parametrize("name", list(CASES))
def test_counts_bit_exact_and_forces_within_tolerance(name, gpu_lib):
    mk, theta, g = CASES[name]
    p = mk()
    t, ref = run_oracle(p, theta, g)
    pkd, out, counts = run_gpu(p, theta, g)
    # the product's own host tree builder must reproduce the oracle's tree exactly
    for k in ("pLower", "pUpper", "iLower", "iUpper", "bnd", "r", "fOpen2", "fSoft", "fMass", "mom"):
        assert np.array_equal(getattr(pkd.tree, k), t[k]), k
    assert np.array_equal(pkd.iOrderMap, t["iOrder"])
    # per-bucket interaction-list counts: bit-exact
    assert np.array_equal(counts, ref["counts"])
    for k in ("nActive", "dPartSum", "dCellSum", "dSoftSum"):
        assert out[k] == ref[k], k
    assert out["dFlop"] == ref["dFlop"]
    assert np.array_equal(out["fWeight"], ref["fWeight"])
    rms, mx = acc_errors(out["acc"], ref["acc"])
    prms, pmx = pot_errors(out["pot"], ref["pot"])
    print(f"{name}: acc rms {rms:.3e} max {mx:.3e}; pot rms {prms:.3e} max {pmx:.3e}")
    assert rms <= RMS_TOL and mx <= MAX_TOL
    assert prms <= RMS_TOL and pmx <= MAX_TOL
    dt = np.abs(out["dtGrav"] - ref["dtGrav"]) / ref["dtGrav"]
    assert dt.max() <= MAX_TOL
    pkd.close()


def test_partial_active_set(gpu_lib):
    """Multistepping calls gravity with few active sinks: the sink box is the bbox of ACTIVE particles only
    (pkd.c:2916-2944) and inactive particles are sources but get no force."""
    p = ics.plummer(20000, seed=7)
    rng = np.random.default_rng(3)
    active = (rng.random(p.n) < 0.3).astype(np.int32)
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    t, ref = run_oracle(p, 0.7, g, active)
    pkd, out, counts = run_gpu(p, 0.7, g, active)
    assert np.array_equal(counts, ref["counts"])
    assert out["nActive"] == ref["nActive"] == int(active.sum())
    act = t["active"].astype(bool)
    assert np.all(out["acc"][~act] == 0.0) and np.all(out["pot"][~act] == 0.0)
    rms, mx = acc_errors(out["acc"][act], ref["acc"][act])
    assert rms <= RMS_TOL and mx <= MAX_TOL
    assert np.array_equal(out["fWeight"], ref["fWeight"])
    pkd.close()


def test_accumulate_semantics(gpu_lib):
    """a, fPot are += onto the caller's values, dtGrav is a running max (SURVEY.md 8b / grav.c:100,192-195)."""
    p = ics.plummer(5000, seed=11)
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    pkd = PKD(fPeriod=p.period)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    pkd.pkdBuildBinary(8, 0.7, 4)
    base = pkd.pkdGravAll(g)
    a = np.full((p.n, 3), 2.0); pot = np.full(p.n, -1.0); dt = np.full(p.n, 1e30); w = np.zeros(p.n)
    pkd.pkdGravAll(g, a, pot, dt, w)
    assert np.allclose(a, 2.0 + base["acc"], rtol=0, atol=1e-12)
    assert np.allclose(pot, -1.0 + base["pot"], rtol=0, atol=1e-12)
    assert np.all(dt == 1e30)
    assert np.array_equal(w, base["fWeight"])
    pkd.close()


def test_bucket_walk_seam_and_lower_orders(gpu_lib):
    p = ics.periodic_box(16)
    for order in (1, 2, 3):
        g = GravityParams(nReps=1, bPeriodic=1, bEwald=1, iOrder=order, iEwOrder=order)
        t, ref = run_oracle(p, 0.7, g)
        pkd, out, counts = run_gpu(p, 0.7, g)
        assert np.array_equal(counts, ref["counts"])
        rms, mx = acc_errors(out["acc"], ref["acc"])
        assert rms <= RMS_TOL and mx <= MAX_TOL, (order, rms, mx)
        assert out["dFlop"] == ref["dFlop"]
        buckets = np.where(t["iLower"] == -1)[0][:5]
        for b in buckets:
            assert pkd.pkdBucketWalk(int(b), g) == tuple(int(v) for v in ref["counts"][b])
        pkd.close()


def test_soft_cells_and_big_softening(gpu_lib):
    """Huge softening pushes accepted cells onto the softened-cell list (walk.c:118-160, SPLINEQ path)."""
    p = ics.plummer(8000, seed=5, eps=0.4)
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    t, ref = run_oracle(p, 0.7, g)
    assert ref["dSoftSum"] > 0
    pkd, out, counts = run_gpu(p, 0.7, g)
    assert np.array_equal(counts, ref["counts"])
    assert out["dSoftSum"] == ref["dSoftSum"]
    rms, mx = acc_errors(out["acc"], ref["acc"])
    prms, pmx = pot_errors(out["pot"], ref["pot"])
    print(f"soft: acc rms {rms:.3e} max {mx:.3e}; pot rms {prms:.3e} max {pmx:.3e}")
    assert rms <= RMS_TOL and mx <= MAX_TOL and prms <= RMS_TOL and pmx <= MAX_TOL
    pkd.close()


def test_ewald_table_matches_oracle(gpu_lib):
    p = ics.periodic_box(16)
    o = oracle.OracleGravity(p); o.build_tree(8, 0.7, 4)
    want = o.ewald_table(2.8, 4); t = o.tree(); o.close()
    pkd = PKD(fPeriod=p.period)
    pkd.pkdDistribRoot(t["root"])
    got = pkd.pkdEwaldInit(2.8, 4)
    assert got.shape == want.shape == (80, 5)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-300)
    pkd.close()


def test_edge_cases_no_active_one_active_coincident(gpu_lib):
    """Ragged inputs: a call with no ACTIVE particle (a rung with nobody on it), with exactly one, and a particle set
    with coincident particles (zero-extent buckets, zero-distance softened pairs) -- each against the oracle."""
    p = ics.plummer(4000, seed=14)
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    none = np.zeros(p.n, np.int32)
    pkd, out, counts = run_gpu(p, 0.7, g, none)
    assert out["nActive"] == 0 and out["dPartSum"] == 0 and out["dCellSum"] == 0 and out["dFlop"] == 0
    assert np.all(out["acc"] == 0) and np.all(out["pot"] == 0) and np.all(counts == -1)
    pkd.close()
    one = none.copy()
    one[1234] = 1
    t, ref = run_oracle(p, 0.7, g, one)
    pkd, out, counts = run_gpu(p, 0.7, g, one)
    assert np.array_equal(counts, ref["counts"]) and out["nActive"] == 1 and out["dFlop"] == ref["dFlop"]
    act = t["active"].astype(bool)
    rms, mx = acc_errors(out["acc"][act], ref["acc"][act])
    assert mx <= MAX_TOL and np.all(out["acc"][~act] == 0)
    pkd.close()
    x, y, z = p.x.copy(), p.y.copy(), p.z.copy()
    idx = np.arange(0, p.n - 1, 5)
    x[idx], y[idx], z[idx] = x[idx + 1], y[idx + 1], z[idx + 1]
    q = ics.Particles(x, y, z, p.m, p.h, p.period, "coincident")
    t, ref = run_oracle(q, 0.7, g)
    pkd, out, counts = run_gpu(q, 0.7, g)
    assert np.array_equal(counts, ref["counts"]) and out["dFlop"] == ref["dFlop"]
    assert np.all(np.isfinite(out["acc"])) and np.all(np.isfinite(out["pot"]))
    rms, mx = acc_errors(out["acc"], ref["acc"])
    prms, pmx = pot_errors(out["pot"], ref["pot"])
    print(f"coincident pairs: acc rms {rms:.3e} max {mx:.3e}; pot rms {prms:.3e} max {pmx:.3e}")
    assert rms <= RMS_TOL and mx <= MAX_TOL and prms <= RMS_TOL and pmx <= MAX_TOL
    pkd.close()


def test_per_bucket_seams_add_up_to_pkdGravAll(gpu_lib):
    """gg_bucket_interact (pkdBucketWalk + pkdBucketInteract, walk.h:32 / grav.h:100) and gg_bucket_ewald (pkdBucketEwald,
    ewald.h:8) for single buckets: tree part + Ewald part = what pkdGravAll leaves for the bucket's particles."""
    from gasoline_b200 import ics
    from gasoline_b200.pkd import PKD, GravityParams
    p = ics.periodic_box(12)
    g = GravityParams(nReps=1, bPeriodic=1, bEwald=1)
    pkd = PKD(fPeriod=p.period)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    pkd.pkdBuildBinary(8, 0.7, 4)
    full = pkd.pkdGravAll(g)
    counts = pkd.pkdBucketCounts()
    t = pkd.tree
    bk = np.where(t.iLower == -1)[0]
    for b in bk[:: max(1, len(bk) // 7)]:
        lo, n = int(t.pLower[b]), int(t.pUpper[b] - t.pLower[b] + 1)
        a, ph, dt, n3 = pkd.pkdBucketInteract(int(b), g)
        ae, pe, nflop = pkd.pkdBucketEwald(int(b), g)
        assert n3 == tuple(counts[b])
        tot = a[:n] + ae[:n]
        rel = np.linalg.norm(tot - full["acc"][lo:lo + n], axis=1) / np.linalg.norm(full["acc"][lo:lo + n], axis=1)
        # (a bucket walked alone gets its lists in another ORDER than as one of a walk group of ten: the FP32 pair terms
        #  then add up in another order -- 1e-7, not 1e-16)
        assert rel.max() < 2e-6, rel.max()
        # (tree part and Ewald part of the potential nearly cancel in a periodic box: the bar is relative to the parts)
        assert np.allclose(ph[:n] + pe[:n], full["pot"][lo:lo + n], rtol=0, atol=5e-6 * np.abs(ph[:n]).max())
        assert np.allclose(dt[:n], full["dtGrav"][lo:lo + n], rtol=1e-5)
        assert nflop > 0
    pkd.close()
