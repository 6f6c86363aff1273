"""pkdBuildBinary on the device (gg_build_local, csrc/gg_tree_gpu.cu) against the host build (gg_tree_build, itself pinned
bit-for-bit to the reference's BuildBinary in tests/test_oracle_vs_reference.py and the golden trees).

The bar is the integer/byte one: every array of the tree -- cell numbering, links, particle ranges, squeezed bounds,
centres of mass, masses, softenings, opening radii -- and the order of the particles inside every bucket must be
IDENTICAL, bit for bit.  Only the multipole moments (bottom-up M2M on the device, particle-by-particle on the host) are
floating-point-close instead; the forces that follow must be identical to the host-tree + device-moments path."""
import numpy as np
import pytest

from gasoline_b200 import ics
from gasoline_b200.pkd import PKD, GravityParams

pytestmark = pytest.mark.gpu

EXACT = ("bnd", "r", "fMass", "fSoft", "fOpen2", "pLower", "pUpper", "iLower", "iUpper")


def _dup(p):
    """every 7th particle duplicated on top of its neighbour: zero-extent cells, ties on the split plane"""
    x, y, z = p.x.copy(), p.y.copy(), p.z.copy()
    idx = np.arange(0, p.n - 1, 7)
    x[idx], y[idx], z[idx] = x[idx + 1], y[idx + 1], z[idx + 1]
    return ics.Particles(x, y, z, p.m, p.h, p.period, p.name + "_dup")


CASES = {
    "plummer20k": (lambda: ics.plummer(20000), 8, 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "plummer50k_b5_theta05": (lambda: ics.plummer(50000, seed=4), 5, 0.5, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "periodic16_ewald": (lambda: ics.periodic_box(16), 8, 0.7, GravityParams(nReps=1, bPeriodic=1, bEwald=1)),
    "periodic24_jitter": (lambda: ics.periodic_box(24, mode="jitter"), 8, 0.7, GravityParams(nReps=1, bPeriodic=1, bEwald=1)),
    "plummer9k_duplicates": (lambda: _dup(ics.plummer(9000, seed=11)), 8, 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "tiny_1": (lambda: ics.plummer(1, seed=2), 8, 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "tiny_2_bucket1": (lambda: ics.plummer(2, seed=2), 1, 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "plummer3000_bucket1": (lambda: ics.plummer(3000, seed=12), 1, 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "plummer5000_bucket64": (lambda: ics.plummer(5000, seed=13), 64, 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "tiny_7": (lambda: ics.plummer(7, seed=2), 8, 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
    "tiny_9": (lambda: ics.plummer(9, seed=2), 8, 0.7, GravityParams(nReps=0, bPeriodic=0, bEwald=0)),
}


def _both(p, nBucket, theta, active=None):
    host = PKD(fPeriod=p.period, device_moments=True)
    host.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, active)
    th = host.pkdBuildBinary(nBucket, theta, 4)
    dev = PKD(fPeriod=p.period)
    dev.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, active)
    nn = dev.pkdBuildBinaryDevice(nBucket, theta)
    return host, th, dev, nn


@pytest.mark.parametrize("name", list(CASES))
def test_device_tree_identical(name, gpu_lib):
    mk, nBucket, theta, g = CASES[name]
    p = mk()
    host, th, dev, nn = _both(p, nBucket, theta)
    assert nn == th.nNodes
    td, pd_ = dev.pkdFetchTree()
    for k in EXACT:
        a, b = getattr(td, k), getattr(th, k)
        assert a.shape == b.shape and np.array_equal(a, b), f"{name}: tree field {k} differs from the host build"
    assert np.array_equal(dev.treeOrder, host.iOrderMap), f"{name}: particle order differs"
    for k, hv in (("x", host.x), ("y", host.y), ("z", host.z), ("fMass", host.fMass), ("fSoft", host.fSoft)):
        assert np.array_equal(pd_[k], hv)
    # moments: same definition, different summation order
    scale = np.abs(th.mom).max(axis=0) + 1e-20  # (a one-particle tree has moments of pure rounding noise)
    assert np.max(np.abs(td.mom - th.mom) / scale) < 1e-9
    nn_, nl, ms = dev.pkdBuildInfo()
    print(f"{name}: {nn_} cells, {nl} levels, device build {ms:.3f} ms")
    # forces: the same tree and the same device-formed moments -> the same bits
    oh, od = host.pkdGravAll(g), dev.pkdGravAll(g)
    assert np.array_equal(host.pkdBucketCounts(), dev.pkdBucketCounts())
    for k in ("dPartSum", "dCellSum", "dSoftSum", "nActive"):
        assert oh[k] == od[k]
    if g.bEwald:  # the Ewald root expansion is summed in another order on the device: FP64-close, not identical
        assert od["dFlop"] == oh["dFlop"]
        assert np.allclose(od["acc"], oh["acc"], rtol=1e-9, atol=1e-12 * np.abs(oh["acc"]).max())
        assert np.allclose(od["pot"], oh["pot"], rtol=1e-9, atol=1e-12 * np.abs(oh["pot"]).max())
    else:
        assert od["dFlop"] == oh["dFlop"]
        for k in ("acc", "pot", "dtGrav", "fWeight"):
            assert np.array_equal(od[k], oh[k]), f"{name}: {k} differs between host-built and device-built tree"
    host.close()
    dev.close()


def test_device_tree_partial_active_and_root(gpu_lib):
    """ACTIVE flags travel with the particles through the device partition; pkdCalcRoot's expansion from the device."""
    p = ics.periodic_box(16)
    rng = np.random.default_rng(3)
    active = (rng.random(p.n) < 0.3).astype(np.int32)
    host, th, dev, nn = _both(p, 8, 0.7, active)
    td, pd_ = dev.pkdFetchTree(with_mom=False)
    assert np.array_equal(pd_["active"], host.active)
    assert np.array_equal(td.fOpen2, th.fOpen2)
    g = GravityParams(nReps=1, bPeriodic=1, bEwald=1)
    oh, od = host.pkdGravAll(g), dev.pkdGravAll(g)
    assert oh["nActive"] == od["nActive"] == int(active.sum())
    assert np.array_equal(host.pkdBucketCounts(), dev.pkdBucketCounts())
    act = host.active != 0
    assert np.allclose(od["acc"][act], oh["acc"][act], rtol=1e-9, atol=1e-12 * np.abs(oh["acc"]).max())
    assert np.all(od["acc"][~act] == 0)
    dev2 = PKD(fPeriod=p.period)
    dev2.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    dev2.pkdBuildBinaryDevice(8, 0.7, want_root=True)
    host2 = PKD(fPeriod=p.period)
    host2.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    host2.pkdBuildBinary(8, 0.7, 4)
    scale = np.maximum(np.abs(host2.ilcnRoot), 1e-12 * np.abs(host2.ilcnRoot).max())
    assert np.max(np.abs(dev2.ilcnRoot - host2.ilcnRoot) / scale) < 1e-6
    assert np.array_equal(dev2.ilcnRoot[:4], host2.ilcnRoot[:4])
    for q in (host, dev, dev2, host2):
        q.close()


def test_device_tree_rebuild_reuses_context(gpu_lib):
    """Build, evaluate, move the particles, build again on the same context (what a time-stepping host does)."""
    p = ics.plummer(15000, seed=8)
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    dev = PKD()
    for step in range(3):
        x = p.x + 0.01 * step * p.y
        dev.pkdLoadParticles(x, p.y, p.z, p.m, p.h)
        dev.pkdBuildBinaryDevice(8, 0.7)
        od = dev.pkdGravAll(g)
        host = PKD(device_moments=True)
        host.pkdLoadParticles(x, p.y, p.z, p.m, p.h)
        host.pkdBuildBinary(8, 0.7, 4)
        oh = host.pkdGravAll(g)
        assert np.array_equal(dev.treeOrder, host.iOrderMap)
        assert np.array_equal(od["acc"], oh["acc"])
        host.close()
    dev.close()


# ---- pkdCalcOpen's other opening criteria (pkd.c:2228-2264), against fixtures from the compiled reference
import os  # noqa: E402
import sys  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_opentypes as _ot  # noqa: E402
from parity import MAX_TOL, RMS_TOL  # noqa: E402


@pytest.mark.parametrize("name", sorted(_ot.CASES))
def test_other_opening_criteria_match_the_reference(name, gpu_lib):
    """Trees with iOpenType = OPEN_ABSPAR (host builder: dRootBracket at orders 1-4) and OPEN_RELPAR / ABSTOT / RELTOT (host AND
    device builder) through the CUDA force path, against tests/golden/opentypes.npz: per-bucket list counts and sums
    bit-exact, forces within the tolerance."""
    gen, args, nBucket, iOpenType, dCrit, iOrder, kw = _ot.CASES[name]
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "opentypes.npz"))
    p, _ = _ot.particles(name)
    g = GravityParams(nReps=kw["nReps"], bPeriodic=kw["bPeriodic"], bEwald=kw["bEwald"], iOrder=iOrder, iEwOrder=iOrder)
    builds = ["host"] if iOpenType == _ot.OPEN_ABSPAR else ["host", "device"]
    for how in builds:
        pkd = PKD(fPeriod=p.period)
        pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, None)
        if how == "host":
            t = pkd.pkdBuildBinary(nBucket, dCrit, iOrder, iOpenType=iOpenType)
            assert np.array_equal(t.fOpen2, z[f"{name}_tree_fOpen2"])
            order = pkd.iOrderMap
        else:
            pkd.pkdBuildBinaryDevice(nBucket, dCrit, iOpenType=iOpenType)
            t, _p = pkd.pkdFetchTree()
            assert np.array_equal(t.fOpen2, z[f"{name}_tree_fOpen2"])
            order = pkd.treeOrder
        assert np.array_equal(order, z[f"{name}_tree_iOrder"])
        out = pkd.pkdGravAll(g)
        assert np.array_equal(pkd.pkdBucketCounts(), z[f"{name}_counts"])
        assert (out["nActive"], out["dPartSum"], out["dCellSum"], out["dSoftSum"], out["dFlop"]) == tuple(z[f"{name}_sums"])
        ref_a, ref_p = z[f"{name}_acc"], z[f"{name}_pot"]
        d = np.linalg.norm(out["acc"] - ref_a, axis=1) / np.maximum(np.linalg.norm(ref_a, axis=1),
                                                                    np.sqrt(np.mean(np.sum(ref_a * ref_a, axis=1))))
        dp = np.abs(out["pot"] - ref_p) / np.maximum(np.abs(ref_p), np.sqrt(np.mean(ref_p ** 2)))
        print(f"{name} [{how}]: acc rms {np.sqrt(np.mean(d * d)):.2e} max {d.max():.2e}; pot max {dp.max():.2e}")
        assert np.sqrt(np.mean(d * d)) <= RMS_TOL and d.max() <= MAX_TOL
        assert np.sqrt(np.mean(dp * dp)) <= RMS_TOL and dp.max() <= MAX_TOL
        pkd.close()


def test_abspar_is_not_built_on_the_device(gpu_lib):
    from gasoline_b200.pkd import GasolineB200Error
    p = ics.plummer(400, seed=3)
    pkd = PKD(fPeriod=p.period)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h, None)
    with pytest.raises(GasolineB200Error, match="OPEN_ABSPAR"):
        pkd.pkdBuildBinaryDevice(16, 1e-3, iOpenType=_ot.OPEN_ABSPAR)
    pkd.close()
