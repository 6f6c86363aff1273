"""CPU: the oracle (oracle/gravity_oracle.c, our restatement) against the golden fixtures produced by the
reference's own compiled code.  This is what pins the oracle; the GPU parity tests then compare against the oracle."""
import numpy as np
import pytest

from golden_cases import NAMES, load, sorted_rows
from oracle import oracle

TREE_KEYS = ("bnd", "r", "fMass", "fSoft", "fOpen2", "mom", "pLower", "pUpper", "iLower", "iUpper", "iOrder", "root")


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_fixture(name):
    p, active, theta, kw, z = load(name)
    order = kw.get("iOrder", 4)
    o = oracle.OracleGravity(p, active=active)
    o.build_tree(8, theta, 4)
    t = o.tree()
    assert t["nNodes"] == int(z["nNodes"]) and t["iRoot"] == int(z["iRoot"])
    for k in TREE_KEYS:  # the tree is reproduced bit for bit (same partition order => same summation order)
        assert np.array_equal(t[k], z["tree_" + k]), k
    res = o.gravity(kw["nReps"], kw["bPeriodic"], order, kw["bEwald"], order)
    assert np.array_equal(res["counts"], z["counts"])  # per-bucket list counts: bit-exact
    sums = z["sums"]
    assert (res["nActive"], res["dPartSum"], res["dCellSum"], res["dSoftSum"], res["dFlop"]) == tuple(sums)
    # forces: the reference's default build uses the v_sqrt1 approximation (<= 2.1e-10 relative),
    # the oracle exact 1/sqrt -- FP64 otherwise
    act = np.ones(p.n, bool) if active is None else t["active"].astype(bool)
    assert np.array_equal(res["fWeight"][act], z["fWeight"][act])  # inactive: untouched by the reference (pkd.c:2851)
    d = np.linalg.norm(res["acc"] - z["acc"], axis=1)[act] / np.linalg.norm(z["acc"], axis=1)[act]
    assert d.max() < 2e-7, d.max()
    scale = np.sqrt(np.mean(z["pot"][act] ** 2))
    assert np.abs(res["pot"] - z["pot"]).max() / scale < 1e-7
    assert np.allclose(res["dtGrav"], z["dtGrav"], rtol=1e-8, atol=0)
    assert np.all(res["acc"][~act] == 0) and np.all(res["pot"][~act] == 0)
    if kw["bPeriodic"]:
        ewt = o.ewald_table(2.8, order)
        assert ewt.shape == z["ewt"].shape
        assert np.allclose(ewt, z["ewt"], rtol=1e-12, atol=1e-300)
    for i, b in enumerate(z["list_buckets"]):
        if f"ilp{i}" not in z:
            continue
        ilp, ilcs, ilcn = o.bucket_lists(int(b), kw["nReps"], order)
        assert ilp.shape == z[f"ilp{i}"].shape and ilcs.shape == z[f"ilcs{i}"].shape and ilcn.shape == z[f"ilcn{i}"].shape
        assert np.array_equal(sorted_rows(ilp), sorted_rows(z[f"ilp{i}"]))
        assert np.array_equal(sorted_rows(ilcs), sorted_rows(z[f"ilcs{i}"]))
        got, want = sorted_rows(ilcn), sorted_rows(z[f"ilcn{i}"])
        nz = {1: 4, 2: 10, 3: 20, 4: 35}[order]  # SETILIST copies moments up to iOrder only (walk.c:10-56)
        assert np.array_equal(got[:, :nz], want[:, :nz])
    o.close()


def test_oracle_threads_do_not_change_results():
    p, active, theta, kw, z = load("plummer3000")
    o = oracle.OracleGravity(p)
    o.build_tree(8, theta, 4)
    a = o.gravity(0, 0, threads=1)
    b = o.gravity(0, 0, threads=4)
    o.close()
    for k in ("acc", "pot", "dtGrav", "fWeight", "counts"):
        assert np.array_equal(a[k], b[k]), k


def test_oracle_import_tree_roundtrip():
    """The oracle can adopt a tree built elsewhere (used to check hosts that bring their own kdNodes)."""
    p, active, theta, kw, z = load("periodic8_jitter_ewald")
    o = oracle.OracleGravity(p)
    o.build_tree(8, theta, 4)
    t = o.tree()
    a = o.gravity(1, 1)
    o.close()
    o2 = oracle.OracleGravity(None, tree=t)
    b = o2.gravity(1, 1)
    o2.close()
    assert np.array_equal(a["counts"], b["counts"]) and np.array_equal(a["acc"], b["acc"])


# ---- the opening criteria of pkdCalcOpen other than OPEN_JOSH (pkd.c:2228-2264)
import make_golden_opentypes as _ot  # noqa: E402  (tests/golden is on sys.path through golden_cases)


@pytest.mark.parametrize("name", sorted(_ot.CASES))
def test_oracle_opening_criteria_match_reference_fixture(name):
    """The oracle's tree with iOpenType = OPEN_ABSPAR (dRootBracket at orders 1-4) / OPEN_RELPAR / ABSTOT / RELTOT and the
    forces that tree gives, against tests/golden/opentypes.npz (the compiled reference): fOpen2 and B2..B6 of every cell,
    per-bucket list counts and the sums bit-exact."""
    import os
    from oracle import oracle
    gen, args, nBucket, iOpenType, dCrit, iOrder, kw = _ot.CASES[name]
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "opentypes.npz"))
    p, _ = _ot.particles(name)
    o = oracle.OracleGravity(p)
    o.build_tree(nBucket, dCrit, iOrder, iOpenType=iOpenType)
    t = o.tree()
    assert np.array_equal(t["fOpen2"], z[f"{name}_tree_fOpen2"])
    nb = {1: 2, 2: 3, 3: 4, 4: 5}[iOrder]  # (the reference initialises B5 / B6 only from octopole / hexadecapole order on)
    assert np.array_equal(o.bnumbers()[:, :nb], z[f"{name}_tree_bmom"][:, 1:1 + nb])
    assert np.array_equal(t["bmax"], z[f"{name}_tree_bmom"][:, 0])
    assert np.array_equal(t["iOrder"], z[f"{name}_tree_iOrder"])
    res = o.gravity(kw["nReps"], kw["bPeriodic"], iOrder, kw["bEwald"], iOrder)
    o.close()
    assert np.array_equal(res["counts"], z[f"{name}_counts"])
    assert (res["nActive"], res["dPartSum"], res["dCellSum"], res["dSoftSum"], res["dFlop"]) == tuple(z[f"{name}_sums"])
    d = np.linalg.norm(res["acc"] - z[f"{name}_acc"], axis=1)
    assert d.max() <= 1e-9 * np.linalg.norm(z[f"{name}_acc"], axis=1).max()
