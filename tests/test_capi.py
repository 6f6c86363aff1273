"""CPU: the C-ABI library loads, exports every symbol include/gasoline_b200.h declares, refuses to run without a
GPU (no CPU fallback), and its host-side tree builder reproduces the reference's tree bit for bit."""
import ctypes as C
import os
import re

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
