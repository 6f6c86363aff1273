"""GPU: seeded random sweep of configurations (tests/random_cases.py -- sizes 1..2500, nBucket 1..33, theta 0.3..1, open
and periodic boxes with 1-2 replicas, Ewald on/off, list orders 1-4, partial active sets, exact duplicates, softenings
over 2.5 decades) -- the CUDA path against the oracle, with the tree built by the host builder and on the device.
Bar: tree, per-bucket list counts, sums, flops and weights bit-exact; accelerations and potentials within the
north_star tolerance measured against max(|value_i|, rms(value)) (random point sets contain particles whose net force
nearly cancels; the strict per-particle figures are printed)."""
import numpy as np
import pytest

from gasoline_b200.pkd import PKD, GravityParams
from oracle import oracle
from parity import MAX_TOL, RMS_TOL, pot_errors
from random_cases import random_case

pytestmark = pytest.mark.gpu


def _acc_errors_floored(a, ref):
    d = np.linalg.norm(a - ref, axis=1)
    nrm = np.linalg.norm(ref, axis=1)
    floor = np.sqrt(np.mean(nrm ** 2))
    rel = d / np.maximum(nrm, floor)
    strict = d[nrm > 0] / nrm[nrm > 0]
    return float(np.sqrt(np.mean(rel ** 2))), float(rel.max()), float(strict.max()) if strict.size else 0.0


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(seed, gpu_lib):
    p, active, nBucket, theta, kw = random_case(seed)
    g = GravityParams(nReps=kw["nReps"], bPeriodic=kw["bPeriodic"], bEwald=kw["bEwald"], iOrder=kw["iOrder"],
                      iEwOrder=kw["iEwOrder"])
    o = oracle.OracleGravity(p, active=active)
    o.build_tree(nBucket, theta, 4)
    t = o.tree()
    ref = o.gravity(g.nReps, g.bPeriodic, g.iOrder, g.bEwald, g.iEwOrder, g.fEwCut, g.fEwhCut)
    o.close()
    act = t["active"].astype(bool)
    for device_build in (False, True):
        pkd = PKD(fPeriod=p.period)
        pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, active)
        if device_build:
            pkd.pkdBuildBinaryDevice(nBucket, theta)
            assert np.array_equal(pkd.treeOrder, t["iOrder"])
        else:
            pkd.pkdBuildBinary(nBucket, theta, 4)
            for k in ("pLower", "pUpper", "iLower", "iUpper", "bnd", "r", "fOpen2", "fSoft", "fMass", "mom"):
                assert np.array_equal(getattr(pkd.tree, k), t[k]), k
            assert np.array_equal(pkd.iOrderMap, t["iOrder"])
        out = pkd.pkdGravAll(g)
        assert np.array_equal(pkd.pkdBucketCounts(), ref["counts"]), (device_build, p.n, nBucket, theta, kw)
        for k in ("nActive", "dPartSum", "dCellSum", "dSoftSum", "dFlop"):
            assert out[k] == ref[k], (k, device_build)
        assert np.array_equal(out["fWeight"][act], ref["fWeight"][act])
        if act.any() and np.linalg.norm(ref["acc"][act], axis=1).max() > 0:
            rms, mx, strict = _acc_errors_floored(out["acc"][act], ref["acc"][act])
            prms, pmx = pot_errors(out["pot"][act], ref["pot"][act])
            print(f"seed {seed} n={p.n} nBucket={nBucket} theta={theta:.2f} {kw} device_build={int(device_build)}: "
                  f"acc rms {rms:.2e} max {mx:.2e} (strict max {strict:.2e}); pot rms {prms:.2e} max {pmx:.2e}")
            assert rms <= RMS_TOL and mx <= MAX_TOL, (rms, mx)
            assert prms <= RMS_TOL and pmx <= MAX_TOL, (prms, pmx)
        pkd.close()
