"""Result delivery of gg_gravity (pkd.c:2851-2861 semantics): zero-copy stores into mapped pinned host arrays must be
bit-identical to the staged device->host copy, for open and periodic (Ewald-first ordering) runs, and must leave
inactive particles untouched."""
import numpy as np
import pytest

from gasoline_b200 import ics
from gasoline_b200.pkd import PKD, GravityParams, pinned_empty

pytestmark = pytest.mark.gpu

CASES = {
    "plummer20k": (lambda: ics.plummer(20000), GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "periodic16_ewald": (lambda: ics.periodic_box(16), GravityParams(nReps=1, bPeriodic=1, bEwald=1)),
    "plummer_comove": (lambda: ics.plummer(5000), GravityParams(nReps=0, bPeriodic=0, bEwald=0, bComove=1, dRhoFac=0.37)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_zero_copy_equals_staged(name, gpu_lib):
    mk, g = CASES[name]
    p = mk()
    pkd = PKD(fPeriod=p.period, pinned=True)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    pkd.pkdBuildBinary(8, 0.7, 4)
    n = pkd.nLocal
    staged = pkd.pkdGravAll(g)  # fresh pageable arrays -> staged copy
    pin = [pinned_empty((n, 3)), pinned_empty(n), pinned_empty(n), pinned_empty(n)]
    for v in pin:
        v[...] = np.nan
    pkd.pkdGravAll(g, *pin, accumulate=False)  # mapped pinned arrays -> stored by the kernels
    for nm, u in zip(("acc", "pot", "dtGrav", "fWeight"), pin):
        assert np.array_equal(u, staged[nm]), nm
    pkd.close()


def test_zero_copy_leaves_inactive_untouched(gpu_lib):
    p = ics.plummer(8000, seed=11)
    active = (np.random.default_rng(5).random(p.n) < 0.4).astype(np.int32)
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    pkd = PKD(fPeriod=p.period, pinned=True)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, active)
    pkd.pkdBuildBinary(8, 0.7, 4)
    n = pkd.nLocal
    act = pkd.active.astype(bool)
    staged = pkd.pkdGravAll(g)
    pin = [pinned_empty((n, 3)), pinned_empty(n), pinned_empty(n), pinned_empty(n)]
    for v in pin:
        v[...] = -7.0
    pkd.pkdGravAll(g, *pin, accumulate=False)
    for nm, u in zip(("acc", "pot", "dtGrav", "fWeight"), pin):
        assert np.array_equal(u[act], staged[nm][act]), nm
        assert np.all(u[~act] == -7.0), nm
    pkd.close()


def test_set_active_equals_fresh_upload(gpu_lib):
    """gg_set_active: new ACTIVE flags on the loaded domain give exactly what a fresh gg_set_local with those flags gives
    (sinks = active only, sources = all; inactive particles untouched), and NULL restores the all-active evaluation."""
    from gasoline_b200 import ics
    from gasoline_b200.pkd import PKD, GravityParams
    p = ics.plummer(12000, seed=17)
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    a = PKD()
    a.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    a.pkdBuildBinary(8, 0.7, 4)
    full = a.pkdGravAll(g)
    rng = np.random.default_rng(4)
    act = (rng.random(p.n) < 0.25).astype(np.int32)  # tree order
    a.pkdSetActive(act)
    part = a.pkdGravAll(g)
    counts = a.pkdBucketCounts()
    b = PKD()
    b.pkdSetTree(a.tree, a.x, a.y, a.z, a.fMass, a.fSoft, active=act)
    ref = b.pkdGravAll(g)
    assert np.array_equal(counts, b.pkdBucketCounts())
    for k in ("nActive", "dPartSum", "dCellSum", "dSoftSum", "dFlop"):
        assert part[k] == ref[k]
    for k in ("acc", "pot", "dtGrav", "fWeight"):
        assert np.array_equal(part[k], ref[k])
    m = act.astype(bool)
    assert np.all(part["acc"][~m] == 0)  # (active sinks see lists built for the ACTIVE bounding boxes: not `full`'s)
    a.pkdSetActive(None)
    again = a.pkdGravAll(g)
    assert np.array_equal(again["acc"], full["acc"]) and again["nActive"] == p.n
    a.close(); b.close()


@pytest.mark.parametrize("name,frac", [("plummer20k", 1.0), ("periodic16_ewald", 1.0), ("plummer20k", 0.5)])
def test_chunked_evaluation_hands_over_final_ranges(name, frac, gpu_lib):
    """gg_gravity_chunked: the list evaluation in several launches; every callback's particle range is final at the time
    of the call (copied there and compared with the one-launch result), the ranges are disjoint, ascending and cover all
    particles; the totals are gg_gravity's."""
    mk, g = CASES[name]
    p = mk()
    active = None if frac == 1.0 else (np.random.default_rng(2).random(p.n) < frac).astype(np.int32)
    pkd = PKD(fPeriod=p.period, pinned=True)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, active)
    pkd.pkdBuildBinary(8, 0.7, 4)
    n = pkd.nLocal
    ref = pkd.pkdGravAll(g)
    act = np.ones(n, bool) if pkd.active is None else pkd.active.astype(bool)
    pin = [pinned_empty((n, 3)), pinned_empty(n), pinned_empty(n), pinned_empty(n)]
    for v in pin:
        v[...] = np.nan
    snap = {}

    def on_chunk(first, count):
        snap[first] = [v[first:first + count].copy() for v in pin]

    out = pkd.pkdGravAllChunked(g, *pin, nChunks=6, on_chunk=on_chunk)
    ch = out["chunks"]
    assert len(ch) > 1 and ch[0][0] == 0 and sum(c for _, c in ch) == n
    assert all(ch[i][0] + ch[i][1] == ch[i + 1][0] for i in range(len(ch) - 1))
    for first, count in ch:
        a = act[first:first + count]
        for nm, v in zip(("acc", "pot", "dtGrav", "fWeight"), snap[first]):
            assert np.array_equal(v[a], ref[nm][first:first + count][a]), (nm, first)
    for nm in ("nActive", "dPartSum", "dCellSum", "dSoftSum", "dFlop"):
        assert out[nm] == ref[nm], nm
    # too few tasks / no callback: one call for everything, same numbers
    small = pkd.pkdGravAllChunked(g, *pin, nChunks=64)
    assert small["chunks"] == [(0, n)]
    for nm, v in zip(("acc", "pot", "dtGrav", "fWeight"), pin):
        assert np.array_equal(v[act], ref[nm][act]), nm
    pkd.close()
