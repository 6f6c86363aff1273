"""GPU: multi-domain force evaluation (top tree + remote trees, the path bench.py --gpus N takes) against MULTI-RANK
runs of the reference binary: per-bucket interaction-list counts bit-exact, per-particle results within tolerance.
The ranks are separate contexts on the one GPU of the test box; the exchange is the in-process driver."""
import numpy as np
import pytest

from gasoline_b200 import domain
from gasoline_b200.pkd import GravityParams
from multirank_cases import NAMES, load, make_domains
from parity import MAX_TOL, RMS_TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["host-arrays", "device-records", "pruned-let"])
@pytest.mark.parametrize("name", NAMES)
def test_multidomain_matches_multirank_reference(name, mode, gpu_lib):
    p, theta, nThreads, z = load(name)
    doms = make_domains(p, theta, nThreads, z, device=0)
    g = GravityParams(nReps=1, bPeriodic=1, bEwald=1) if p.periodic else GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    domain.run_in_process(doms, packed=mode == "device-records", let=g if mode == "pruned-let" else None)
    if mode == "pruned-let":
        sent = sum(v[0] for v in domain.run_in_process.let_stats.values())
        full = sum(v[1] for v in domain.run_in_process.let_stats.values())
        print(f"{name}: pruned trees are {100.0 * sent / full:.1f} % of the whole domains")
        assert sent <= full
    for r, d in enumerate(doms):
        out = d.pkd.pkdGravAll(g)
        counts = d.pkd.pkdBucketCounts()
        bk = z[f"r{r}_buckets"]
        assert np.array_equal(counts[bk[:, 0]], bk[:, 3:6]), f"rank {r}: per-bucket list counts differ"
        walked = np.zeros(len(counts), bool); walked[bk[:, 0]] = True
        assert np.all(counts[~walked] == -1)
        sums = z[f"r{r}_sums"]
        assert (out["dPartSum"], out["dCellSum"], out["dSoftSum"], out["dFlop"]) == tuple(sums)
        res = z[f"r{r}_res"]
        rel = np.linalg.norm(out["acc"] - res[:, 0:3], axis=1) / np.linalg.norm(res[:, 0:3], axis=1)
        rms, mx = float(np.sqrt(np.mean(rel ** 2))), float(rel.max())
        floor = np.sqrt(np.mean(res[:, 3] ** 2))
        dp = np.abs(out["pot"] - res[:, 3]) / np.maximum(np.abs(res[:, 3]), floor)
        print(f"{name} rank {r}: acc rms {rms:.2e} max {mx:.2e}; pot max {dp.max():.2e}")
        assert rms <= RMS_TOL and mx <= MAX_TOL
        assert np.sqrt(np.mean(dp ** 2)) <= RMS_TOL and dp.max() <= MAX_TOL
        assert (np.abs(out["dtGrav"] - res[:, 4]) / res[:, 4]).max() <= MAX_TOL
        assert np.array_equal(out["fWeight"], res[:, 5])
    for d in doms:
        d.pkd.close()


@pytest.mark.parametrize("case", ["plummer200k_r4", "periodic32_r3"])
def test_pruned_let_equals_whole_domains(case, gpu_lib):
    """The locally essential trees (gg_let_export) must give every bucket exactly the lists the whole remote domains
    give -- same entries in the same order, hence bit-identical forces -- while shipping a fraction of the bytes."""
    from gasoline_b200 import ics
    if case == "plummer200k_r4":
        p, world, g = ics.plummer(200_000, seed=9), 4, GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    else:
        p, world, g = ics.periodic_box(32), 3, GravityParams(nReps=1, bPeriodic=1, bEwald=1)
    parts = domain.orb_decompose(p.x, p.y, p.z, world)
    results = {}
    for mode in ("whole", "let"):
        doms = [domain.Domain(r, world, p.x[ix], p.y[ix], p.z[ix], p.m[ix], p.h[ix], p.period, 0.7, device=0)
                for r, ix in enumerate(parts)]
        domain.run_in_process(doms, packed=mode == "whole", let=g if mode == "let" else None)
        results[mode] = [(d.pkd.pkdGravAll(g), d.pkd.pkdBucketCounts()) for d in doms]
        if mode == "let":
            st = domain.run_in_process.let_stats
            sent, full = sum(v[0] for v in st.values()), sum(v[1] for v in st.values())
            print(f"{case}: pruned trees are {100.0 * sent / full:.1f} % of the whole domains ({sent / 1e6:.1f} of {full / 1e6:.1f} MB)")
            assert sent < 0.7 * full
        for d in doms:
            d.pkd.close()
    for (a, ca), (b, cb) in zip(results["whole"], results["let"]):
        assert np.array_equal(ca, cb)
        for k in ("dPartSum", "dCellSum", "dSoftSum", "dFlop"):
            assert a[k] == b[k]
        assert np.array_equal(a["acc"], b["acc"]) and np.array_equal(a["pot"], b["pot"])
        assert np.array_equal(a["fWeight"], b["fWeight"])


@pytest.mark.parametrize("case", ["plummer200k_r4", "periodic32_r3"])
def test_device_built_domains_equal_host_built(case, gpu_lib):
    """Multi-domain run with every rank's tree built ON THE GPU (gg_build_local; root summaries and ancestor sums from
    gg_domain_summary / gg_domain_moments_about) against the host-built run: identical local trees and identical top-tree
    geometry (r, fMass, fSoft, fOpen2 -- Bmax is a maximum) give identical interaction lists; the top cells' moments are
    summed in another order (translated root record instead of particle by particle), so forces agree to FP64 rounding
    of those few cells' moments, far inside the tolerance."""
    from gasoline_b200 import ics
    if case == "plummer200k_r4":
        p, world, g = ics.plummer(200_000, seed=9), 4, GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    else:
        p, world, g = ics.periodic_box(32), 3, GravityParams(nReps=1, bPeriodic=1, bEwald=1)
    parts = domain.orb_decompose(p.x, p.y, p.z, world)
    results, tops = {}, {}
    for mode in ("host", "device"):
        doms = [domain.Domain(r, world, p.x[ix], p.y[ix], p.z[ix], p.m[ix], p.h[ix], p.period, 0.7, device=0,
                              device_build=mode == "device") for r, ix in enumerate(parts)]
        domain.run_in_process(doms, let=g)
        tops[mode] = (doms[0].kdTop, doms[0].ilcnRoot)
        results[mode] = [(d.pkd.pkdGravAll(g), d.pkd.pkdBucketCounts(),
                          d.pkd.treeOrder.copy() if mode == "device" else d.host.iOrderMap.copy()) for d in doms]
        for d in doms:
            d.pkd.close()
    th, td = tops["host"][0], tops["device"][0]
    for k in ("pLower", "bUsed", "r", "fMass", "fSoft", "fOpen2", "bnd"):
        assert np.array_equal(th[k], td[k]), f"top tree field {k} differs"
    scale = np.abs(th["mom"]).max(axis=0) + 1e-300
    assert np.max(np.abs(th["mom"] - td["mom"]) / scale) < 1e-9
    assert np.allclose(tops["host"][1], tops["device"][1], rtol=1e-6, atol=1e-9 * np.abs(tops["host"][1]).max())
    for (a, ca, oa), (b, cb, ob) in zip(results["host"], results["device"]):
        assert np.array_equal(oa, ob), "particle order differs"
        assert np.array_equal(ca, cb), "per-bucket list counts differ"
        for k in ("dPartSum", "dCellSum", "dSoftSum", "dFlop"):
            assert a[k] == b[k]
        rel = np.linalg.norm(a["acc"] - b["acc"], axis=1) / np.linalg.norm(a["acc"], axis=1)
        print(f"{case}: device-built vs host-built domains: acc max rel diff {rel.max():.2e}")
        assert rel.max() < 1e-6
        assert np.array_equal(a["fWeight"], b["fWeight"])
