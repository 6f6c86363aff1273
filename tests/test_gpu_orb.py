"""GPU: the ORB domain decomposition services (gg_orb_*, csrc/gg_orb.cu) and the decomposition they drive
(domain.pst_domain_decomp) against the oracle restatement of pstDomainDecomp/_pstRootSplit (oracle/orb_oracle.py, pinned
to the compiled reference's domains in the CPU suite) and against the reference's own domains (golden fixtures).
Bar: splits, counts and every particle's destination bit-exact."""
import time

import numpy as np
import pytest

from gasoline_b200 import domain, ics
from gasoline_b200.pkd import PKD, GasolineB200Error
from multirank_cases import NAMES, load
from oracle import orb_oracle
from orb_stub import HostOrbRank

pytestmark = pytest.mark.gpu


def _decompose(p, nThreads, nCtx=1, weights=None, split_work=True, resident=False):
    owner = np.arange(p.n) % nCtx if nCtx > 1 else np.zeros(p.n, np.int64)
    idx = [np.nonzero(owner == s)[0] for s in range(nCtx)]
    pkds = [PKD(device=0, fPeriod=p.period) for _ in range(nCtx)]
    for k, i in zip(pkds, idx):
        w = None if weights is None else weights[i]
        if resident:
            zero = np.zeros(len(i))
            k.pkdLoadResident(p.x[i], p.y[i], p.z[i], zero, zero, zero, p.m[i], p.h[i])
            k.pkdOrbLoad(fWeight=w)
        else:
            k.pkdOrbLoad(p.x[i], p.y[i], p.z[i], fWeight=w)
    nodes = domain.pst_domain_decomp(pkds, nThreads, split_work=split_work)
    dest = np.zeros(p.n, np.int32)
    lr = domain.leaf_rank(nThreads)
    for i, k in zip(idx, pkds):
        dest[i] = lr[k.pkdOrbCells()]
        k.close()
    return nodes, dest


def test_services_equal_host_stand_in():
    p = ics.plummer(30000, seed=9)
    rng = np.random.default_rng(2)
    w = rng.uniform(0.5, 4.0, p.n)
    k = PKD(device=0, fPeriod=p.period)
    k.pkdOrbLoad(p.x, p.y, p.z, fWeight=w)
    h = HostOrbRank(p.x, p.y, p.z, w)
    b, n = k.pkdCalcBound([1])
    hb, hn = h.pkdCalcBound([1])
    assert np.array_equal(b, hb) and np.array_equal(n, hn)
    got, ref = k.pkdWeight([1], [1], [0.03]), h.pkdWeight([1], [1], [0.03])
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    assert np.allclose(got[2], ref[2], rtol=1e-13, atol=0) and np.allclose(got[3], ref[3], rtol=1e-13, atol=0)
    again = k.pkdWeight([1], [1], [0.03])
    assert np.array_equal(got[2], again[2]) and np.array_equal(got[3], again[3])  # fixed summation order
    k.pkdOrbSplit([1], [1], [0.03]); h.pkdOrbSplit([1], [1], [0.03])
    assert np.array_equal(k.pkdOrbCells(), h.pkdOrbCells())
    # two cells of the next level in one request, plus a cell that holds nothing (heap index 7 does not exist yet)
    cells, dims, splits = [2, 3, 7], [0, 2, 0], [-0.2, 0.4, 0.0]
    b, n = k.pkdCalcBound(cells)
    hb, hn = h.pkdCalcBound(cells)
    assert np.array_equal(b, hb) and np.array_equal(n, hn) and n[2] == 0
    got, ref = k.pkdWeight(cells, dims, splits), h.pkdWeight(cells, dims, splits)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    assert np.allclose(got[2], ref[2], rtol=1e-13, atol=0) and np.allclose(got[3], ref[3], rtol=1e-13, atol=0)
    k.pkdOrbSplit(cells[:2], dims[:2], splits[:2]); h.pkdOrbSplit(cells[:2], dims[:2], splits[:2])
    assert np.array_equal(k.pkdOrbCells(), h.pkdOrbCells())
    with pytest.raises(GasolineB200Error):
        k.pkdWeight([2, 2], [0, 0], [0.0, 0.0])
    k.close()


@pytest.mark.parametrize("name", NAMES)
def test_device_decomposition_reproduces_the_reference_domains(name):
    p, theta, nThreads, z = load(name)
    for sw in (True, False):
        nodes, dest = _decompose(p, nThreads, nCtx=2, split_work=sw)
        for r in range(nThreads):
            assert np.array_equal(np.nonzero(dest == r)[0], np.sort(z[f"r{r}_iOrder"]))


@pytest.mark.parametrize("nThreads,nCtx,resident", [(2, 1, False), (3, 2, False), (8, 3, False), (5, 1, True), (8, 1, True)])
def test_device_decomposition_equals_oracle(nThreads, nCtx, resident):
    p = ics.plummer(50000, seed=13)
    doms, ref_nodes = orb_oracle.domain_decomp(p.x, p.y, p.z, nThreads)
    nodes, dest = _decompose(p, nThreads, nCtx=nCtx, resident=resident)
    for r in range(nThreads):
        assert np.array_equal(np.nonzero(dest == r)[0], doms[r])
    ref = {n[0]: n for n in ref_nodes}
    for n in nodes:
        assert n["iDim"] == ref[n["iCell"]][1] and n["fSplit"] == ref[n["iCell"]][2]
        assert np.array_equal(n["bnd"], ref[n["iCell"]][3])


def test_device_decomposition_weights_periodic_and_empty_rank():
    p = ics.periodic_box(20)
    w = np.ones(p.n); w[p.z > 0.1] = 5.0  # integer-valued work weights: all sums exact
    doms, _ = orb_oracle.domain_decomp(p.x, p.y, p.z, 8, weights=w)
    nodes, dest = _decompose(p, 8, nCtx=2, weights=w)
    for r in range(8):
        assert np.array_equal(np.nonzero(dest == r)[0], doms[r])
    # a service rank without particles takes part (the reference's pkdCalcBound on an empty store)
    a, b = PKD(device=0, fPeriod=p.period), PKD(device=0, fPeriod=p.period)
    a.pkdOrbLoad(p.x, p.y, p.z); b.pkdOrbLoad(p.x[:0], p.y[:0], p.z[:0])
    domain.pst_domain_decomp([a, b], 4)
    doms4, _ = orb_oracle.domain_decomp(p.x, p.y, p.z, 4)
    dest = domain.leaf_rank(4)[a.pkdOrbCells()]
    for r in range(4):
        assert np.array_equal(np.nonzero(dest == r)[0], doms4[r])
    assert len(b.pkdOrbCells()) == 0
    a.close(); b.close()


def test_device_decomposition_full_size():
    """1 M Plummer particles into 8 domains: equal to the oracle; time of the whole decomposition on the device."""
    p = ics.plummer(1000000, seed=12345)
    k = PKD(device=0, fPeriod=p.period)
    k.pkdOrbLoad(p.x, p.y, p.z)
    t0 = time.perf_counter()
    nodes = domain.pst_domain_decomp([k], 8)
    cells = k.pkdOrbCells()
    ms = (time.perf_counter() - t0) * 1e3
    dest = domain.leaf_rank(8)[cells]
    k.close()
    t0 = time.perf_counter()
    doms, _ = orb_oracle.domain_decomp(p.x, p.y, p.z, 8)
    ms_cpu = (time.perf_counter() - t0) * 1e3
    for r in range(8):
        assert np.array_equal(np.nonzero(dest == r)[0], doms[r])
    print(f"ORB 1 M particles -> 8 domains: device {ms:.1f} ms ({sum(n['ittr'] for n in nodes)} bisection steps in "
          f"3 levels), numpy restatement {ms_cpu:.0f} ms; domain sizes {np.bincount(dest).tolist()}")


def test_later_decompositions_on_the_device_match_the_stepping_reference():
    """Second, third, ... decomposition of a run (tests/golden/orbsteps_*.npz: a TIME-STEPPING multi-rank run of the reference
    binary): the device services under domain.pst_domain_decomp with `prev` = the previous cells' axis and split and the
    work weights the previous force evaluation left -- every particle on the reference's rank."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_orbsteps import CASES
    for name in sorted(CASES):
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
        nThreads, nSteps = int(z["nThreads"]), int(z["nSteps"])
        # *_overflow: runs with small particle stores; where a split would overfill a side the reference bisects a second
        # boundary into the cell (pst.c:1049-1270) -- counted by gg_orb_weight, applied by gg_orb_split_wrap
        stores = domain.rank_stores(nThreads, len(z["s0_pos"]), float(z["dExtraStore"])) if name.endswith("_overflow") else None
        nFixed = 0
        prev = None
        for k in range(nSteps + 1):
            pos, want = z[f"s{k}_pos"], z[f"s{k}_rank"]
            w = None if k == 0 else z[f"s{k - 1}_fWeight"]
            idx = [np.nonzero(np.arange(len(pos)) % 2 == s)[0] for s in range(2)]
            pkds = [PKD(device=0) for _ in idx]
            for q, i in zip(pkds, idx):
                q.pkdOrbLoad(pos[i, 0], pos[i, 1], pos[i, 2], fWeight=None if w is None else w[i])
            cells = domain.pst_domain_decomp(pkds, nThreads, prev=prev, stores=stores)
            dest = np.zeros(len(pos), np.int32)
            for i, q in zip(idx, pkds):
                dest[i] = domain.leaf_rank(nThreads)[q.pkdOrbCells()]
                q.close()
            assert np.array_equal(dest, want), f"{name}: decomposition {k} differs from the reference's"
            nFixed += sum(1 for c in cells if c.get("fixed"))
            prev = cells
        assert stores is None or nFixed >= 1, f"{name}: no boundary moved"


def test_bisection_on_the_device_equals_the_host_driven_loop():
    """gg_orb_bisect (the root finder's state on the device, no host round trip per trial) against the host-driven loop over
    gg_orb_weight: same split axis, split, trial count for every PST cell and the same cell for every particle -- counts
    mode, integer and non-integer work weights (the weight sums are order-fixed in both), 2-13 ranks, and the later
    decompositions of a stepping run (prev axis / split kept or re-found)."""
    import os
    import sys

    def both(load, nThreads, **kw):
        res = []
        for dev in (False, True):
            k = PKD(device=0)
            load(k)
            nodes = domain.pst_domain_decomp([k], nThreads, device_bisect=dev, **kw)
            res.append((nodes, k.pkdOrbCells().copy()))
            k.close()
        (na, ca), (nb, cb) = res
        assert len(na) == len(nb)
        for u, v in zip(na, nb):
            assert (u["iCell"], u["iDim"], u["ittr"]) == (v["iCell"], v["iDim"], v["ittr"]) and u["fSplit"] == v["fSplit"], (u, v)
        assert np.array_equal(ca, cb)
        return nb

    p = ics.plummer(30000, seed=3)
    rng = np.random.default_rng(4)
    for nThreads in (2, 3, 5, 8, 13):
        both(lambda k: k.pkdOrbLoad(p.x, p.y, p.z), nThreads)
        both(lambda k: k.pkdOrbLoad(p.x, p.y, p.z), nThreads, split_work=False)
    w_int = rng.integers(1, 9, p.n).astype(np.float64)
    w_real = rng.uniform(0.5, 40.0, p.n)
    for w in (w_int, w_real):
        both(lambda k: k.pkdOrbLoad(p.x, p.y, p.z, fWeight=w), 8)
    q = ics.periodic_box(24)
    prev = both(lambda k: k.pkdOrbLoad(q.x, q.y, q.z), 6)
    x2 = q.x + rng.normal(0, 0.01, q.n)
    for flags in (dict(), dict(bDoRootFind=False), dict(bDoRootFind=False, bDoSplitDimFind=False)):
        both(lambda k: k.pkdOrbLoad(x2, q.y, q.z, fWeight=rng.uniform(1, 3, q.n) if False else None), 6, prev=prev, **flags)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_orbsteps import CASES
    name = sorted(n for n in CASES if not n.endswith("_overflow"))[0]
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    nThreads, prev = int(z["nThreads"]), None
    for s in range(int(z["nSteps"]) + 1):
        pos, want = z[f"s{s}_pos"], z[f"s{s}_rank"]
        w = None if s == 0 else z[f"s{s - 1}_fWeight"]
        k = PKD(device=0)
        k.pkdOrbLoad(pos[:, 0], pos[:, 1], pos[:, 2], fWeight=w)
        prev = domain.pst_domain_decomp([k], nThreads, prev=prev, device_bisect=True)
        assert np.array_equal(domain.leaf_rank(nThreads)[k.pkdOrbCells()], want), f"{name}: decomposition {s}"
        k.close()


def test_bisection_on_the_device_timing():
    p = ics.plummer(1000000, seed=12345)
    k = PKD(device=0, fPeriod=p.period)
    out = {}
    for dev in (False, True, True):
        k.pkdOrbLoad(p.x, p.y, p.z)
        t0 = time.perf_counter()
        nodes = domain.pst_domain_decomp([k], 8, device_bisect=dev)
        out[dev] = ((time.perf_counter() - t0) * 1e3, sum(n["ittr"] for n in nodes))
    k.close()
    print(f"ORB 1 M particles -> 8 domains, {out[True][1]} trials: host-driven loop {out[False][0]:.2f} ms, bisection on the "
          f"device {out[True][0]:.2f} ms")
    assert out[True][1] == out[False][1]


def _collective_decomp(p, nThreads, shares, weights=None, **kw):
    """Every rank a thread with its own context and share of the particles (in-process group): pst_domain_decomp with the
    bounds through gg_comm_allgather and each level's root finder as ONE gg_orb_bisect_all."""
    import threading
    from gasoline_b200 import pkd as _pkd
    grp = _pkd.Group(nThreads)
    res, errs = [None] * nThreads, []

    def work(r):
        try:
            k = PKD(device=0, fPeriod=p.period)
            i = shares[r]
            k.pkdOrbLoad(p.x[i], p.y[i], p.z[i], fWeight=None if weights is None else weights[i])
            k.commInitLocal(grp, r)
            nodes = domain.pst_domain_decomp([k], nThreads, reduce=domain.orb_reduce_lib(k), collective_bisect=True, **kw)
            res[r] = (nodes, k.pkdOrbCells().copy())
            k.close()
        except Exception as e:  # noqa: BLE001 -- surfaced below
            errs.append(e)

    th = [threading.Thread(target=work, args=(r,)) for r in range(nThreads)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    if errs:
        raise errs[0]
    assert not any(t.is_alive() for t in th), "a rank did not return from the collective decomposition"
    return res


def test_collective_bisection_equals_the_host_driven_loop():
    """gg_orb_bisect_all (ranks = threads with an in-process group here; NCCL in a multi-process job) against the host-driven
    loop over the SAME shares (one thread, the ranks' answers added in rank order on the host): every PST cell's axis, split
    and trial count and every particle's cell identical -- counts, integer and real work weights, an empty rank, uneven
    shares, and a later decomposition that starts from the previous axes."""
    p = ics.plummer(40000, seed=5)
    rng = np.random.default_rng(8)
    w_real = rng.uniform(0.5, 40.0, p.n)
    for nThreads, weights, kw in ((2, None, {}), (3, None, dict(split_work=False)), (4, w_real, {}), (5, w_real, {}),
                                  (3, rng.integers(1, 9, p.n).astype(np.float64), {})):
        cuts = np.sort(rng.choice(np.arange(1, p.n), nThreads - 1, replace=False))
        if nThreads == 5:
            cuts[1] = cuts[0]  # rank 1 starts with no particles at all
        shares = np.split(np.arange(p.n), cuts)
        pkds = [PKD(device=0, fPeriod=p.period) for _ in range(nThreads)]
        for k, i in zip(pkds, shares):
            k.pkdOrbLoad(p.x[i], p.y[i], p.z[i], fWeight=None if weights is None else weights[i])
        ref_nodes = domain.pst_domain_decomp(pkds, nThreads, **kw)
        ref_cells = [k.pkdOrbCells().copy() for k in pkds]
        for k in pkds:
            k.close()
        got = _collective_decomp(p, nThreads, shares, weights, **kw)
        for r in range(nThreads):
            nodes, cells = got[r]
            assert len(nodes) == len(ref_nodes)
            for u, v in zip(ref_nodes, nodes):
                assert (u["iCell"], u["iDim"], u["ittr"]) == (v["iCell"], v["iDim"], v["ittr"]) and u["fSplit"] == v["fSplit"], (r, u, v)
            assert np.array_equal(cells, ref_cells[r]), f"rank {r} of {nThreads}"
        if nThreads == 4:  # a later decomposition: previous axes / splits, particles moved
            q = ics.Particles(p.x + rng.normal(0, 0.02, p.n), p.y, p.z, p.m, p.h, p.period, "moved")
            pk2 = [PKD(device=0, fPeriod=p.period) for _ in range(nThreads)]
            for k, i in zip(pk2, shares):
                k.pkdOrbLoad(q.x[i], q.y[i], q.z[i], fWeight=weights[i])
            ref2 = domain.pst_domain_decomp(pk2, nThreads, prev=ref_nodes, bDoRootFind=False)
            ref2_cells = [k.pkdOrbCells().copy() for k in pk2]
            for k in pk2:
                k.close()
            got2 = _collective_decomp(q, nThreads, shares, weights, prev=ref_nodes, bDoRootFind=False)
            for r in range(nThreads):
                assert [(u["iDim"], u["fSplit"], u["ittr"]) for u in got2[r][0]] == [(u["iDim"], u["fSplit"], u["ittr"]) for u in ref2]
                assert np.array_equal(got2[r][1], ref2_cells[r])
