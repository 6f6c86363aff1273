"""CPU: the C-ABI library loads, exports every symbol include/gasoline_b200.h declares, refuses to run without a
GPU (no CPU fallback), and its host-side tree builder reproduces the reference's tree bit for bit."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

from golden_cases import NAMES, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gasoline_b200 import build, pkd
    build.build()
    return pkd.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gasoline_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gasoline_b200.h but not exported"
    assert lib.gg_version() >= 100


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_error_not_fallback(lib):
    from gasoline_b200.pkd import PKD, GasolineB200Error
    ctx = C.c_void_p()
    rc = lib.gg_create(C.byref(ctx), 0)
    assert rc == -1 and not ctx.value  # GG_ERR_CUDA
    assert b"no CPU path" in lib.gg_last_error()
    with pytest.raises(GasolineB200Error):
        PKD()


def test_missing_library_raises(tmp_path, monkeypatch):
    from gasoline_b200 import pkd
    monkeypatch.setattr(pkd, "_lib", None)
    with pytest.raises(pkd.GasolineB200Error):
        pkd.load_library(str(tmp_path / "nope.so"))


@pytest.mark.parametrize("name", NAMES)
def test_host_tree_builder_matches_reference_tree(lib, name):
    """gg_tree_build (pkdBuildBinary semantics) vs the reference's tree in the golden fixture: bit-exact."""
    from gasoline_b200 import pkd as pk
    p, active, theta, kw, z = load(name)
    x, y, zz, m, h = (np.array(a, dtype=np.float64) for a in (p.x, p.y, p.z, p.m, p.h))
    order = np.zeros(p.n, np.int32)
    act = None if active is None else pk._i(active.copy())
    for nthreads in (1, 3):
        xs, ys, zs, ms, hs = x.copy(), y.copy(), zz.copy(), m.copy(), h.copy()
        bt = C.c_void_p()
        assert lib.gg_tree_build(p.n, pk._d(xs), pk._d(ys), pk._d(zs), pk._d(ms), pk._d(hs), act, pk._i(order), 8,
                                 theta, 4, nthreads, C.byref(bt)) == 0
        v = pk.gg_tree()
        root = np.zeros(35)
        assert lib.gg_tree_view(bt, C.byref(v), pk._d(root)) == 0
        nn = v.nNodes
        assert nn == int(z["nNodes"]) and v.iRoot == int(z["iRoot"])
        arr = lambda ptr, shape: np.ctypeslib.as_array(ptr, shape=shape)
        assert np.array_equal(arr(v.bnd, (nn, 6)), z["tree_bnd"])
        assert np.array_equal(arr(v.r, (nn, 3)), z["tree_r"])
        assert np.array_equal(arr(v.fMass, (nn,)), z["tree_fMass"])
        assert np.array_equal(arr(v.fSoft, (nn,)), z["tree_fSoft"])
        assert np.array_equal(arr(v.fOpen2, (nn,)), z["tree_fOpen2"])
        assert np.array_equal(arr(v.mom, (nn, 31)), z["tree_mom"])
        for k in ("pLower", "pUpper", "iLower", "iUpper"):
            assert np.array_equal(arr(getattr(v, k), (nn,)), z["tree_" + k]), k
        assert np.array_equal(order, z["tree_iOrder"])
        assert np.array_equal(root, z["tree_root"])
        lib.gg_tree_free(bt)


@pytest.mark.parametrize("name", NAMES)
def test_bottom_up_moments_match_particle_sums(lib, name):
    """gg_tree_moments_m2m = the device's moment algorithm (gg_moments.cu / gg_m2m.h: raw bucket moments, children
    translated and summed, reduced as pkd.c:2056-2131) run on the host, against the REFERENCE's particle-by-particle
    pkdCalcCell sums stored in the golden fixture.  Same quantity, different summation order: agreement to FP64
    rounding of the largest term, M * Bmax^l."""
    from gasoline_b200 import pkd as pk
    p, active, theta, kw, z = load(name)
    nn = int(z["nNodes"])
    order = z["tree_iOrder"]
    cols = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)[order]) for a in (p.x, p.y, p.z, p.m, p.h)]
    ints = {k: np.ascontiguousarray(z["tree_" + k], dtype=np.int32) for k in ("pLower", "pUpper", "iLower", "iUpper")}
    dbl = {k: np.ascontiguousarray(z["tree_" + k], dtype=np.float64) for k in ("bnd", "r", "fMass", "fSoft", "fOpen2", "mom")}
    tv = pk.gg_tree(nn, int(z["iRoot"]), pk._d(dbl["bnd"]), pk._d(dbl["r"]), pk._d(dbl["fMass"]), pk._d(dbl["fSoft"]),
                    pk._d(dbl["fOpen2"]), None, pk._i(ints["pLower"]), pk._i(ints["pUpper"]), pk._i(ints["iLower"]),
                    pk._i(ints["iUpper"]))
    pv = pk.gg_particles(p.n, *[pk._d(c) for c in cols], None)
    out = np.zeros((nn, 31))
    assert lib.gg_tree_moments_m2m(C.byref(tv), C.byref(pv), pk._d(out)) == 0
    ref = dbl["mom"]
    # scale of order l: M * Bmax^l, with Bmax recovered from fOpen2 = max(Bmax, 2/sqrt(3) Bmax/theta)^2
    bmax = np.sqrt(dbl["fOpen2"]) / max(1.0, 2.0 / np.sqrt(3.0) / theta)
    M = dbl["fMass"]
    for lo, hi, l in ((0, 6, 2), (6, 16, 3), (16, 31, 4)):
        scale = (M * bmax ** l)[:, None]
        ok = scale[:, 0] > 0
        err = np.abs(out[ok, lo:hi] - ref[ok, lo:hi]) / scale[ok]
        assert err.max() < 2e-13, (l, err.max())
    # single-particle / zero-extent cells: both are exactly zero
    assert np.all(out[bmax == 0] == 0)


@pytest.mark.parametrize("seed", range(24))
def test_host_tree_builder_random_configurations(lib, seed):
    """gg_tree_build (the product's host-side pkdBuildBinary) against the oracle's tree on the seeded random sweep
    (tests/random_cases.py: 1..2500 particles, nBucket 1..33, duplicates, partial active sets, open and periodic);
    the oracle is pinned to the compiled reference on the same sweep (tests/test_oracle_vs_reference.py).  Bit-exact,
    single- and multi-threaded."""
    from gasoline_b200 import pkd as pk
    from oracle import oracle
    from random_cases import random_case
    p, active, nBucket, theta, kw = random_case(seed)
    o = oracle.OracleGravity(p, active=active)
    o.build_tree(nBucket, theta, 4)
    t = o.tree()
    o.close()
    for nthreads in (1, 4):
        cols = [np.array(a, dtype=np.float64) for a in (p.x, p.y, p.z, p.m, p.h)]
        order = np.zeros(p.n, np.int32)
        act = None if active is None else active.astype(np.int32).copy()
        bt = C.c_void_p()
        assert lib.gg_tree_build(p.n, *[pk._d(c) for c in cols], None if act is None else pk._i(act), pk._i(order), nBucket,
                                 theta, 4, nthreads, C.byref(bt)) == 0
        v = pk.gg_tree()
        root = np.zeros(35)
        assert lib.gg_tree_view(bt, C.byref(v), pk._d(root)) == 0
        nn = v.nNodes
        assert nn == t["nNodes"] and v.iRoot == t["iRoot"]
        arr = lambda ptr, shape: np.ctypeslib.as_array(ptr, shape=shape)
        for k, shape in (("bnd", (nn, 6)), ("r", (nn, 3)), ("fMass", (nn,)), ("fSoft", (nn,)), ("fOpen2", (nn,)), ("mom", (nn, 31)),
                         ("pLower", (nn,)), ("pUpper", (nn,)), ("iLower", (nn,)), ("iUpper", (nn,))):
            assert np.array_equal(arr(getattr(v, k), shape), t[k]), (k, p.n, nBucket, nthreads)
        assert np.array_equal(order, t["iOrder"])
        assert np.array_equal(root, t["root"])
        assert np.array_equal(cols[0], t["x"]) and np.array_equal(cols[4], t["h"])  # particles permuted into tree order
        if act is not None:
            assert np.array_equal(act, t["active"])
        lib.gg_tree_free(bt)


# ---- the opening criteria of pkdCalcOpen other than OPEN_JOSH (pkd.c:2228-2264), pinned to the compiled reference
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_opentypes as _ot  # noqa: E402


@pytest.mark.parametrize("name", sorted(_ot.CASES))
def test_host_tree_builder_other_opening_criteria(lib, name):
    """gg_tree_build_open against tests/golden/opentypes.npz (the reference's pstBuildTree with iOpenType = OPEN_ABSPAR at
    orders 1-4 -- dRootBracket's hunt and bisection, pkd.c:2182-2224 -- and OPEN_RELPAR / ABSTOT / RELTOT): fOpen2 and the
    radial moments Bmax, B2..B6 of every cell bit-exact, tree and multipoles as before."""
    from gasoline_b200 import pkd as pk
    gen, args, nBucket, iOpenType, dCrit, iOrder, kw = _ot.CASES[name]
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "opentypes.npz"))
    p, _ = _ot.particles(name)
    for nthreads in (1, 4):
        cols = [np.array(a, dtype=np.float64) for a in (p.x, p.y, p.z, p.m, p.h)]
        order = np.zeros(p.n, np.int32)
        bt = C.c_void_p()
        assert lib.gg_tree_build_open(p.n, *[pk._d(c) for c in cols], None, pk._i(order), nBucket, iOpenType, dCrit, iOrder,
                                      nthreads, C.byref(bt)) == 0
        v = pk.gg_tree()
        assert lib.gg_tree_view(bt, C.byref(v), None) == 0
        nn = v.nNodes
        assert nn == len(z[f"{name}_tree_fOpen2"])
        arr = lambda ptr, shape: np.ctypeslib.as_array(ptr, shape=shape)
        assert np.array_equal(arr(v.fOpen2, (nn,)), z[f"{name}_tree_fOpen2"])
        bm = np.zeros((nn, 6))
        assert lib.gg_tree_bnumbers(bt, pk._d(bm)) == 0
        nb = {1: 3, 2: 4, 3: 5, 4: 6}[iOrder]  # pkdCalcCell initialises B5 / B6 only from octopole / hexadecapole order on
        assert np.array_equal(bm[:, :nb], z[f"{name}_tree_bmom"][:, :nb])
        nm = {1: 6, 2: 6, 3: 16, 4: 31}[iOrder]
        assert np.array_equal(arr(v.mom, (nn, 31))[:, :nm], z[f"{name}_tree_mom"][:, :nm])
        assert np.array_equal(arr(v.r, (nn, 3)), z[f"{name}_tree_r"])
        for k in ("pLower", "pUpper", "iLower", "iUpper"):
            assert np.array_equal(arr(getattr(v, k), (nn,)), z[f"{name}_tree_{k}"]), k
        assert np.array_equal(order, z[f"{name}_tree_iOrder"])
        lib.gg_tree_free(bt)
    if iOpenType == _ot.OPEN_ABSPAR:  # the criterion really differs from theta's
        josh = (2 / np.sqrt(3.0) * z[f"{name}_tree_bmom"][:, 0] / 0.7) ** 2
        assert not np.allclose(josh, z[f"{name}_tree_fOpen2"], rtol=1e-3)
    else:
        assert np.array_equal(z[f"{name}_tree_fOpen2"], z[f"{name}_tree_bmom"][:, 0] ** 2)


def test_abspar_on_a_cell_without_extent_is_an_error_not_a_hang(lib):
    """A one-particle bucket has Bmax = B_k = 0; the reference's dRootBracket (pkd.c:2182-2224) evaluates 0/0 there and
    never leaves its hunt loop (observed with oracle/_ref).  The builder reports GG_ERR_UNSUPPORTED instead."""
    from gasoline_b200 import ics, pkd as pk
    p = ics.plummer(500, seed=3)
    cols = [np.array(a, dtype=np.float64) for a in (p.x, p.y, p.z, p.m, p.h)]
    bt = C.c_void_p()
    rc = lib.gg_tree_build_open(p.n, *[pk._d(c) for c in cols], None, None, 8, _ot.OPEN_ABSPAR, 1e-3, 4, 2, C.byref(bt))
    assert rc == -3 and not bt.value
    assert lib.gg_tree_build_open(p.n, *[pk._d(c) for c in cols], None, None, 8, 6, 0.7, 4, 2, C.byref(bt)) == -2
