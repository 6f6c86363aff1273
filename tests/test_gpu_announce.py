"""gg_announce: the Ewald correction launched DURING gg_set_local (particles first, slice by slice, side stream) must give
bit for bit what the ordinary order gives -- it is the same kernel on the same inputs (ewald.c:15-178) -- and every
sequence the announcement does not cover (other parameters, a new sink set, a second evaluation) must fall back cleanly."""
import numpy as np
import pytest

from gasoline_b200 import ics
from gasoline_b200.pkd import PKD, GravityParams

pytestmark = pytest.mark.gpu
KEYS = ("acc", "pot", "dtGrav", "fWeight")


def _pkd(p, active=None):
    k = PKD(fPeriod=p.period, pinned=True)
    k.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, active)
    k.pkdBuildBinary(8, 0.7, 4)
    return k


@pytest.mark.parametrize("device_moments", [False, True])
@pytest.mark.parametrize("frac", [1.0, 0.45])
def test_announced_ewald_equals_ordinary_order(frac, device_moments, gpu_lib):
    p = ics.periodic_box(20, seed=7)
    active = None if frac == 1.0 else (np.random.default_rng(3).random(p.n) < frac).astype(np.int32)
    g = GravityParams(nReps=1, bPeriodic=1, bEwald=1)
    k = _pkd(p, active)
    k.device_moments = device_moments
    k.upload()
    ref = k.pkdGravAll(g)
    assert ref["dFlopEwald"] > 0
    k.upload(announce=g)
    out = k.pkdGravAll(g)
    for nm in KEYS:
        assert np.array_equal(out[nm], ref[nm]), nm
    for nm in ("nActive", "dPartSum", "dCellSum", "dSoftSum", "dFlop", "dFlopEwald"):
        assert out[nm] == ref[nm], nm
    # a second evaluation of the same upload computes the correction itself
    again = k.pkdGravAll(g)
    for nm in KEYS:
        assert np.array_equal(again[nm], ref[nm]), nm
    k.close()


def test_announcement_not_matching_the_evaluation_falls_back(gpu_lib):
    p = ics.periodic_box(16, seed=2)
    g = GravityParams(nReps=1, bPeriodic=1, bEwald=1)
    g3 = GravityParams(nReps=1, bPeriodic=1, bEwald=1, iEwOrder=3)
    gno = GravityParams(nReps=1, bPeriodic=1, bEwald=0)
    k = _pkd(p)
    k.upload()
    ref, ref3, refno = k.pkdGravAll(g), k.pkdGravAll(g3), k.pkdGravAll(gno)
    for ann, ev, want in ((g, g3, ref3), (g3, g, ref), (g, gno, refno), (gno, g, ref)):
        k.upload(announce=ann)
        out = k.pkdGravAll(ev)
        for nm in KEYS:
            assert np.array_equal(out[nm], want[nm]), nm
    # a new sink set between the upload and the evaluation: the early correction was for the old one
    act = (np.random.default_rng(9).random(p.n) < 0.5).astype(np.int32)
    k.upload(announce=g)
    k.pkdSetActive(act)
    out = k.pkdGravAll(g)
    k2 = _pkd(p)
    k2.upload()
    k2.pkdSetActive(act)
    want = k2.pkdGravAll(g)
    a = act.astype(bool)  # (flags are in tree order; both trees are the same build)
    for nm in KEYS:
        assert np.array_equal(out[nm][a], want[nm][a]), nm
    k.close(); k2.close()


def test_announced_upload_in_slices(gpu_lib):
    """More particles than one slice of the early upload (GG_EARLY_SLICE = 2^20): 128^3 = 2.1 M particles, three slices."""
    p = ics.periodic_box(128, seed=1)
    g = GravityParams(nReps=1, bPeriodic=1, bEwald=1)
    k = _pkd(p)
    k.device_moments = True
    k.upload()
    ref = k.pkdGravAll(g)
    k.upload(announce=g)
    out = k.pkdGravAll(g)
    for nm in KEYS:
        assert np.array_equal(out[nm], ref[nm]), nm
    print(f"128^3: Ewald {ref['msEwald']:.2f} ms in the ordinary order, {out['msEwald']:.2f} ms beside the upload; "
          f"device total {ref['msTotal']:.2f} -> {out['msTotal']:.2f} ms")
    k.close()


@pytest.mark.parametrize("case", ["plummer", "periodic_ewald", "periodic_ewald_active"])
def test_sliced_upload_equals_gg_set_local(case, gpu_lib):
    """gg_local_begin / _particles / _nodes / _end (slices in any order; early Ewald per slice when announced) load the very
    domain gg_set_local loads: every result bit for bit."""
    if case == "plummer":
        p, g, active = ics.plummer(30000, seed=5), GravityParams(nReps=0, bPeriodic=0, bEwald=0), None
    else:
        p, g = ics.periodic_box(24, seed=4), GravityParams(nReps=1, bPeriodic=1, bEwald=1)
        active = (np.random.default_rng(8).random(p.n) < 0.6).astype(np.int32) if case.endswith("active") else None
    k = _pkd(p, active)
    k.device_moments = True
    k.upload()
    ref = k.pkdGravAll(g)
    counts = k.pkdBucketCounts()
    for nSlices, ann in ((1, None), (5, None), (3, g)):
        k.upload_sliced(nSlices, announce=ann)
        out = k.pkdGravAll(g)
        for nm in KEYS:
            assert np.array_equal(out[nm], ref[nm]), (nm, nSlices)
        assert np.array_equal(k.pkdBucketCounts(), counts)
        for nm in ("nActive", "dPartSum", "dCellSum", "dSoftSum", "dFlop"):
            assert out[nm] == ref[nm], nm
    k.close()


def test_sliced_upload_argument_checks(gpu_lib):
    import ctypes as C
    from gasoline_b200.pkd import GasolineB200Error
    p = ics.plummer(2000, seed=1)
    k = _pkd(p)
    L, ctx = k._L, k._ctx
    assert L.gg_local_end(ctx) != 0  # nothing begun
    assert L.gg_local_begin(ctx, 0, k.tree.nNodes, k.tree.iRoot, k.nLocal, None, 0) == 0
    assert L.gg_local_particles(ctx, 10, k.nLocal, *([C.c_void_p(k.x.ctypes.data)] * 5), None) != 0  # runs past the end
    assert L.gg_local_end(ctx) != 0  # incomplete
    with pytest.raises(GasolineB200Error):
        k._uploaded = True
        k.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0))  # no domain loaded after the failed sequence
    k.upload()
    assert k.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0))["nActive"] == p.n
    k.close()
