"""gg_tree.mom = NULL: the device forms the cells' multipole moments itself (gg_moments.cu) instead of receiving
pkd->kdNodes[].mom.  The opening decisions read only r / fOpen2 / fSoft (still the host's), so the per-bucket list
counts stay bit-exact; forces must meet the same tolerance against the oracle and be practically identical to the
run with uploaded moments (the moments are rounded to FP32 either way)."""
import numpy as np
import pytest

from gasoline_b200 import ics
from gasoline_b200.pkd import PKD, GravityParams
from oracle import oracle
from parity import MAX_TOL, RMS_TOL, acc_errors, pot_errors

pytestmark = pytest.mark.gpu

CASES = {
    "plummer20k": (lambda: ics.plummer(20000), 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "periodic16_ewald": (lambda: ics.periodic_box(16), 0.7, GravityParams(nReps=1, bPeriodic=1, bEwald=1)),
    "plummer30k_theta05_order3": (lambda: ics.plummer(30000, seed=5), 0.5, GravityParams(nReps=0, bPeriodic=0, bEwald=0, iOrder=3)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_device_moments(name, gpu_lib):
    mk, theta, g = CASES[name]
    p = mk()
    o = oracle.OracleGravity(p)
    o.build_tree(8, theta, 4)
    ref = o.gravity(g.nReps, g.bPeriodic, g.iOrder, g.bEwald, g.iEwOrder, g.fEwCut, g.fEwhCut)
    o.close()
    res = {}
    for dm in (False, True):
        pkd = PKD(fPeriod=p.period, device_moments=dm)
        pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
        pkd.pkdBuildBinary(8, theta, 4)
        out = pkd.pkdGravAll(g)
        counts = pkd.pkdBucketCounts()
        if dm:
            assert pkd.upload_bytes() < 0.5 * up_host
        else:
            up_host = pkd.upload_bytes()
        pkd.close()
        assert np.array_equal(counts, ref["counts"])
        assert out["dFlop"] == ref["dFlop"]
        rms, mx = acc_errors(out["acc"], ref["acc"])
        prms, pmx = pot_errors(out["pot"], ref["pot"])
        print(f"{name} device_moments={dm}: acc rms {rms:.3e} max {mx:.3e}; pot rms {prms:.3e} max {pmx:.3e}")
        assert rms <= RMS_TOL and mx <= MAX_TOL and prms <= RMS_TOL and pmx <= MAX_TOL
        res[dm] = out
    rms, mx = acc_errors(res[True]["acc"], res[False]["acc"])
    print(f"{name}: device vs uploaded moments: acc rms {rms:.3e} max {mx:.3e}")
    assert rms <= 1e-7 and mx <= 1e-5


def test_device_moments_softened_cells(gpu_lib):
    """Large softening: the softened-cell path (ILCS, FP64) reads the raw quadrupole the device computed."""
    p = ics.plummer(6000, seed=9, eps=0.4)
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    o = oracle.OracleGravity(p)
    o.build_tree(8, 0.7, 4)
    ref = o.gravity(g.nReps, g.bPeriodic, g.iOrder, g.bEwald, g.iEwOrder, g.fEwCut, g.fEwhCut)
    o.close()
    pkd = PKD(fPeriod=p.period, device_moments=True)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    pkd.pkdBuildBinary(8, 0.7, 4)
    out = pkd.pkdGravAll(g)
    assert out["dSoftSum"] == ref["dSoftSum"] and ref["dSoftSum"] > 0
    assert np.array_equal(pkd.pkdBucketCounts(), ref["counts"])
    rms, mx = acc_errors(out["acc"], ref["acc"])
    assert rms <= RMS_TOL and mx <= MAX_TOL
    pkd.close()
