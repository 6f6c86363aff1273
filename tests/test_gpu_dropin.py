"""The drop-in itself: the REFERENCE HOST (the reference's own objects: pstBuildTree -> pstGravity, oracle/_ref/
libgasref_gpu.so, built by oracle/Makefile) with its pkdGravAll link-substituted by gasoline_b200/csrc/pkd_gravall_shim.c,
so that the reference's pstGravity runs on the GPU through the C ABI.  Compared with the golden fixtures the PURE
reference produced: the scalars pkdGravAll returns (nActive, dPartSum, dCellSum, dSoftSum, dFlop) bit-exact, fWeight
bit-exact, accelerations / potentials / dtGrav within the north-star tolerance, inactive particles untouched."""
import time

import numpy as np
import pytest

from gasoline_b200 import ics
from golden_cases import NAMES, load
from oracle import reflib
from parity import MAX_TOL, RMS_TOL

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reflib.gpu_host_available(), reason="oracle/_ref/libgasref_gpu.so not built")]


@pytest.mark.parametrize("name", NAMES)
def test_reference_host_dropin(name, gpu_lib):
    p, active, theta, kw, z = load(name)
    order = kw.get("iOrder", 4)
    r = reflib.RefGravity(p, active=active, gpu_host=True)
    r.build_tree(8, theta, 4)  # the reference's own tree build
    t = r.tree()
    assert np.array_equal(t["iOrder"], z["tree_iOrder"])
    out = r.gravity(kw["nReps"], kw["bPeriodic"], order, kw["bEwald"], order)
    r.close()
    assert (out["nActive"], out["dPartSum"], out["dCellSum"], out["dSoftSum"], out["dFlop"]) == tuple(z["sums"])
    act = np.ones(p.n, bool) if active is None else t["active"].astype(bool)
    assert np.array_equal(out["fWeight"][act], z["fWeight"][act])
    d = np.linalg.norm(out["acc"] - z["acc"], axis=1)[act] / np.linalg.norm(z["acc"], axis=1)[act]
    rms, mx = float(np.sqrt(np.mean(d * d))), float(d.max())
    scale = np.sqrt(np.mean(z["pot"][act] ** 2))
    dp = np.abs(out["pot"] - z["pot"])[act] / np.maximum(np.abs(z["pot"][act]), scale)
    print(f"{name} (reference host + GPU pkdGravAll): acc rms {rms:.2e} max {mx:.2e}; pot max {dp.max():.2e}")
    assert rms <= RMS_TOL and mx <= MAX_TOL
    assert np.sqrt(np.mean(dp * dp)) <= RMS_TOL and dp.max() <= MAX_TOL
    assert (np.abs(out["dtGrav"] - z["dtGrav"])[act] / z["dtGrav"][act]).max() <= MAX_TOL
    if active is not None:  # pkdGravAll must not touch inactive particles (pkd.c:2851-2861)
        assert np.all(out["acc"][~act] == 0) and np.all(out["pot"][~act] == 0)


def test_reference_host_dropin_full_size_timing(gpu_lib):
    """BASELINE.json configs[1] through the reference host: what a Gasoline user sees per pstGravity call."""
    p = ics.plummer(1_000_000)
    r = reflib.RefGravity(p, gpu_host=True)
    tb = r.build_tree(8, 0.7, 4)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        out = r.gravity(0, 0, 4, 0, 4)
        best = min(best, out["seconds"])
    r.close()
    inter = out["dPartSum"] + out["dCellSum"] + out["dSoftSum"]
    print(f"reference host, 1 M Plummer: its own tree build {tb:.2f} s; pstGravity -> GPU pkdGravAll {best * 1e3:.1f} ms "
          f"({inter / best:.3g} interactions/s incl. flattening PARTICLE/KDN and writing back)")
    assert out["nActive"] == p.n and inter == 618322384.0
    f = p.m[0] * out["acc"]
    assert np.linalg.norm(f.sum(axis=0)) <= 2e-3 * np.linalg.norm(f, axis=1).sum()


@pytest.mark.parametrize("name", NAMES)
def test_reference_host_dropin_device_tree(name, gpu_lib, monkeypatch):
    """GG_SHIM_DEVICE_TREE=1: the reference host's pkdBuildBinary is ALSO served by the GPU (gg_build_local; pStore
    permuted and kdNodes filled by the shim).  The reference's own code then exports the tree it sees: every field the
    gravity path reads must equal the PURE reference's tree bit for bit (moments to FP64 rounding), and pstGravity on it
    must return the reference's sums."""
    monkeypatch.setenv("GG_SHIM_DEVICE_TREE", "1")
    p, active, theta, kw, z = load(name)
    order = kw.get("iOrder", 4)
    r = reflib.RefGravity(p, active=active, gpu_host=True)
    r.build_tree(8, theta, 4)
    t = r.tree()
    assert t["nNodes"] == int(z["nNodes"]) and t["iRoot"] == int(z["iRoot"])
    for k in ("bnd", "r", "fMass", "fSoft", "fOpen2", "pLower", "pUpper", "iLower", "iUpper"):
        assert np.array_equal(t[k], z["tree_" + k]), f"{name}: kdNodes field {k} differs from the reference's tree"
    assert np.array_equal(t["iOrder"], z["tree_iOrder"])
    scale = np.abs(z["tree_mom"]).max(axis=0) + 1e-300
    assert np.max(np.abs(t["mom"] - z["tree_mom"]) / scale) < 1e-9
    assert np.allclose(t["root"], z["tree_root"], rtol=1e-9, atol=1e-12 * np.abs(z["tree_root"]).max())
    out = r.gravity(kw["nReps"], kw["bPeriodic"], order, kw["bEwald"], order)
    r.close()
    assert (out["nActive"], out["dPartSum"], out["dCellSum"], out["dSoftSum"], out["dFlop"]) == tuple(z["sums"])
    act = np.ones(p.n, bool) if active is None else t["active"].astype(bool)
    d = np.linalg.norm(out["acc"] - z["acc"], axis=1)[act] / np.linalg.norm(z["acc"], axis=1)[act]
    assert float(np.sqrt(np.mean(d * d))) <= RMS_TOL and float(d.max()) <= MAX_TOL
    assert np.array_equal(out["fWeight"][act], z["fWeight"][act])


def test_reference_host_dropin_device_tree_full_size_timing(gpu_lib, monkeypatch):
    """The reference host's whole force step (pstBuildTree + pstGravity) at 1 M particles with both entry points on the GPU."""
    monkeypatch.setenv("GG_SHIM_DEVICE_TREE", "1")
    p = ics.plummer(1_000_000)
    r = reflib.RefGravity(p, gpu_host=True)
    tb = min(r.build_tree(8, 0.7, 4) for _ in range(3))
    out = r.gravity(0, 0, 4, 0, 4)
    out = r.gravity(0, 0, 4, 0, 4)
    r.close()
    inter = out["dPartSum"] + out["dCellSum"] + out["dSoftSum"]
    print(f"reference host, 1 M Plummer, GG_SHIM_DEVICE_TREE=1: msrBuildTree sequence {tb * 1e3:.1f} ms (pkdBuildBinary on "
          f"the GPU + pStore permutation + kdNodes fill + the host's pstColCells/pstCalcRoot), pstGravity {out['seconds'] * 1e3:.1f} ms")
    assert inter == 618322384.0 and out["nActive"] == p.n
