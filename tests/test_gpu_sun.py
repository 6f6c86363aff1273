"""bDoSun (pkd.c:3003-3041): the indirect acceleration at the origin -- a dummy sink of softening dSunSoft in a cell of
+-1e-14 that walks the tree and is evaluated like any bucket -- against golden vectors produced by the compiled
reference (tests/golden/make_golden_sun.py): the dummy bucket's interaction-list counts bit-exact, aSun within the
north-star tolerance, and the particles' own results untouched by the extra pass."""
import os
import sys

import numpy as np
import pytest

from gasoline_b200.pkd import PKD, GasolineB200Error, GravityParams
from oracle import reflib
from parity import MAX_TOL

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_sun import NAMES, case  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sun.npz"))


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("build", ["host", "device"])
def test_sun_indirect_term(name, build, gpu_lib):
    p, theta, soft = case(name)
    pkd = PKD()
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    if build == "host":
        pkd.pkdBuildBinary(8, theta, 4)
    else:
        pkd.pkdBuildBinaryDevice(8, theta)
    plain = pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0))
    counts_plain = pkd.pkdBucketCounts()
    out = pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0, bDoSun=1, dSunSoft=soft))
    assert (out["nSunPart"], out["nSunCellSoft"], out["nSunCellNewt"]) == tuple(GOLD[name + "_counts"])
    ref = GOLD[name + "_aSun"]
    err = np.linalg.norm(out["aSun"] - ref) / np.linalg.norm(ref)
    print(f"sun {name} ({build} tree): aSun {out['aSun']}, relative error {err:.2e}")
    assert err <= MAX_TOL
    # the dummy pass is not counted and leaves the particles alone (pkd.c:3003-3041 runs after the bucket loop)
    assert (out["nActive"], out["dPartSum"], out["dCellSum"], out["dSoftSum"], out["dFlop"]) == tuple(GOLD[name + "_sums"])
    for k in ("acc", "pot", "dtGrav", "fWeight"):
        assert np.array_equal(out[k], plain[k])
    assert np.array_equal(pkd.pkdBucketCounts(), counts_plain)
    assert np.all(plain["aSun"] == 0)
    pkd.close()


def test_sun_rejects_periodic(gpu_lib):
    from gasoline_b200 import ics
    p = ics.periodic_box(8)
    pkd = PKD(fPeriod=p.period)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    pkd.pkdBuildBinary(8, 0.7, 4)
    with pytest.raises(GasolineB200Error):  # the reference asserts nReps == 0 && !bPeriodic (pkd.c:3013-3014)
        pkd.pkdGravAll(GravityParams(nReps=1, bPeriodic=1, bEwald=1, bDoSun=1, dSunSoft=0.01))
    pkd.close()


@pytest.mark.skipif(not reflib.gpu_host_available(), reason="oracle/_ref/libgasref_gpu.so not built")
def test_sun_through_the_reference_host(gpu_lib):
    """pstGravity with bDoSun = 1 on the reference host whose pkdGravAll is ours: aSun as the reference returns it."""
    p, theta, soft = case("inside")
    r = reflib.RefGravity(p, gpu_host=True)
    r.build_tree(8, theta, 4)
    a, _ = r.gravity_sun(soft)
    r.close()
    ref = GOLD["inside_aSun"]
    assert np.linalg.norm(a - ref) / np.linalg.norm(ref) <= MAX_TOL


def test_comove_background_term(gpu_lib):
    """bComove && !bPeriodic (pkd.c:2967-2991): a += dRhoFac r, fPot -= dRhoFac r^2 / 2 on ACTIVE particles, against the
    compiled reference's pstGravity (golden), partially active."""
    from parity import RMS_TOL, acc_errors, pot_errors
    p, theta, _ = case("inside")
    active = (np.arange(p.n) % 3 != 0).astype(np.int32)
    pkd = PKD()
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, active)
    pkd.pkdBuildBinary(8, theta, 4)
    assert np.array_equal(pkd.active, GOLD["comove_active"])
    out = pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0, bComove=1, dRhoFac=0.37))
    plain = pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0))
    act = pkd.active != 0
    rms, mx = acc_errors(out["acc"][act], GOLD["comove_acc"][act])
    prms, pmx = pot_errors(out["pot"][act], GOLD["comove_pot"][act])
    print(f"comove: acc rms {rms:.2e} max {mx:.2e}; pot rms {prms:.2e} max {pmx:.2e}")
    assert rms <= RMS_TOL and mx <= MAX_TOL and prms <= RMS_TOL and pmx <= MAX_TOL
    r = np.stack([pkd.x, pkd.y, pkd.z], axis=1)
    assert np.allclose((out["acc"] - plain["acc"])[act], 0.37 * r[act], rtol=1e-12, atol=1e-13)
    assert np.all(out["acc"][~act] == 0) and np.all(GOLD["comove_acc"][~act] == 0)
    # bDoSun on top of bComove: the reference adds the background term once per bucket of the main loop, before the Sun
    # pass (pkd.c:2967-2991 vs 3003-3041) -- the dummy sink's pass must not apply it a second time
    both = pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0, bComove=1, dRhoFac=0.37, bDoSun=1, dSunSoft=0.01))
    assert np.array_equal(both["acc"], out["acc"]) and np.array_equal(both["pot"], out["pot"])
    sun_only = pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0, bDoSun=1, dSunSoft=0.01))
    assert np.array_equal(both["aSun"], sun_only["aSun"])
    pkd.close()


def test_sun_with_remote_domains(gpu_lib):
    """bDoSun on a rank that also holds remote domains (pst.c:3258-3277 hands it to the one rank whose domain contains the
    origin; that rank's dummy sink walks the top tree, its own tree and the remote trees).  With an opening angle so small
    that every cell is opened the result does not depend on how the particles are split into trees: the dummy's list is
    all N particles and aSun equals the one-domain value to rounding; at theta = 0.7 it agrees within the tree error."""
    from gasoline_b200 import domain, ics
    p = ics.plummer(4000, seed=21)
    soft = 0.01
    one = PKD()
    one.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    # theta 0.7: the net pull at the centre of the sphere nearly cancels (|aSun| is a few per cent of a typical particle's
    # acceleration), so two different trees agree only to the TREE error of the force -- measured against the typical
    # acceleration, not against |aSun| itself
    # (the pruned trees a rank receives are complete only for sinks inside ITS box: like the reference, which hands bDoSun
    #  to the rank holding the origin, theta 0.7 is checked on that rank; with theta 0.02 every cell is opened anyway)
    for theta, tol in ((0.02, 2e-6), (0.7, 2e-2)):
        one.pkdBuildBinary(8, theta, 4)
        ref = one.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0, bDoSun=1, dSunSoft=soft))
        scale = np.linalg.norm(ref["aSun"]) if theta < 0.1 else np.sqrt((ref["acc"] ** 2).sum(axis=1).mean())
        parts = domain.orb_decompose(p.x, p.y, p.z, 3)
        doms = [domain.Domain(r, 3, p.x[ix], p.y[ix], p.z[ix], p.m[ix], p.h[ix], p.period, theta, device=0)
                for r, ix in enumerate(parts)]
        g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
        domain.run_in_process(doms, let=g)
        for r, d in enumerate(doms):
            ix = parts[r]
            holds_origin = all(v[ix].min() <= 0.0 <= v[ix].max() for v in (p.x, p.y, p.z))
            if theta > 0.1 and not holds_origin:
                continue
            plain = d.pkd.pkdGravAll(g)
            out = d.pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0, bDoSun=1, dSunSoft=soft))
            err = np.linalg.norm(out["aSun"] - ref["aSun"]) / scale
            print(f"sun with remote domains, theta {theta}, rank {r}: aSun diff to the one-domain run {err:.2e} "
                  f"(|aSun| = {np.linalg.norm(ref['aSun']) / scale:.2e} of the scale), "
                  f"lists {out['nSunPart']}/{out['nSunCellSoft']}/{out['nSunCellNewt']}")
            assert err <= tol
            if theta < 0.1:
                assert out["nSunPart"] == p.n and out["nSunCellNewt"] == 0
            for k in ("acc", "pot", "dtGrav", "fWeight"):  # the particles' own results are untouched by the extra pass
                assert np.array_equal(out[k], plain[k])
        for d in doms:
            d.pkd.close()
    one.close()
