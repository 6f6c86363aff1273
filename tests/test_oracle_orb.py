"""CPU: the ORB domain decomposition.
 * oracle/orb_oracle.py (restatement of pstDomainDecomp + _pstRootSplit) is PINNED against the domains the compiled
   reference produced on 2, 3 and 4 pthread-MDL ranks (tests/golden/multirank_*.npz, r<k>_iOrder), particle for particle;
 * the product's host-side driver (gasoline_b200/domain.py: pst_domain_decomp) gives the same splits and domains as the
   oracle when served by a numpy stand-in for the device services (tests/orb_stub.py), with the particles spread over
   1, 2 and 5 service ranks, and across two real processes (torch.distributed, gloo)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from gasoline_b200 import domain, ics
from multirank_cases import NAMES, load
from oracle import orb_oracle
from orb_stub import HostOrbRank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("split_work", [True, False])
def test_orb_oracle_reproduces_the_reference_domains(name, split_work):
    p, theta, nThreads, z = load(name)
    doms, nodes = orb_oracle.domain_decomp(p.x, p.y, p.z, nThreads, split_work=split_work)
    for r in range(nThreads):
        assert np.array_equal(np.sort(z[f"r{r}_iOrder"]), doms[r]), f"rank {r}: domain differs from the reference's"
    assert len(nodes) == nThreads - 1


def _run_driver(p, nThreads, nService, weights=None, split_work=True):
    rng = np.random.default_rng(7)
    owner = rng.integers(0, nService, p.n)
    idx = [np.nonzero(owner == s)[0] for s in range(nService)]
    ranks = [HostOrbRank(p.x[i], p.y[i], p.z[i], None if weights is None else weights[i]) for i in idx]
    nodes = domain.pst_domain_decomp(ranks, nThreads, split_work=split_work)
    dest = np.zeros(p.n, np.int32)
    lr = domain.leaf_rank(nThreads)
    for i, r in zip(idx, ranks):
        dest[i] = lr[r.pkdOrbCells()]
    return nodes, dest


@pytest.mark.parametrize("nThreads", [2, 3, 5, 8])
@pytest.mark.parametrize("nService", [1, 2, 5])
def test_driver_equals_oracle(nThreads, nService):
    p = ics.plummer(6000, seed=11)
    doms, ref_nodes = orb_oracle.domain_decomp(p.x, p.y, p.z, nThreads)
    nodes, dest = _run_driver(p, nThreads, nService)
    assert dest.min() >= 0
    for r in range(nThreads):
        assert np.array_equal(np.nonzero(dest == r)[0], doms[r])
    ref = {n[0]: n for n in ref_nodes}
    for n in nodes:
        assert n["iDim"] == ref[n["iCell"]][1] and n["fSplit"] == ref[n["iCell"]][2]
        assert np.array_equal(n["bnd"], ref[n["iCell"]][3])


def test_driver_equals_oracle_on_reference_cases_and_count_mode():
    for name in NAMES:
        p, theta, nThreads, z = load(name)
        for sw in (True, False):
            nodes, dest = _run_driver(p, nThreads, 2, split_work=sw)
            for r in range(nThreads):
                assert np.array_equal(np.nonzero(dest == r)[0], np.sort(z[f"r{r}_iOrder"]))


def test_driver_with_integer_weights_balances_work():
    p = ics.plummer(8000, seed=3)
    w = np.ones(p.n)
    w[p.x > 0] = 3.0  # integer-valued: every sum is exact in any order
    doms, _ = orb_oracle.domain_decomp(p.x, p.y, p.z, 4, weights=w)
    nodes, dest = _run_driver(p, 4, 3, weights=w)
    for r in range(4):
        assert np.array_equal(np.nonzero(dest == r)[0], doms[r])
    tot = [w[dest == r].sum() for r in range(4)]
    assert max(tot) - min(tot) <= 8.0


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from gasoline_b200 import domain, ics
from orb_stub import HostOrbRank
from oracle import orb_oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
p = ics.plummer(5000, seed=21)
mine = np.arange(p.n)[rank::world]            # any initial distribution: every other particle
svc = HostOrbRank(p.x[mine], p.y[mine], p.z[mine])
nodes = domain.pst_domain_decomp([svc], 4, reduce=domain.orb_reduce_dist("cpu"))
dest = domain.leaf_rank(4)[svc.pkdOrbCells()]
doms, ref_nodes = orb_oracle.domain_decomp(p.x, p.y, p.z, 4)
for r in range(4):
    assert np.array_equal(mine[dest == r], doms[r][np.isin(doms[r], mine)])
assert [n["fSplit"] for n in nodes] == [n[2] for n in sorted(ref_nodes, key=lambda n: n[0])]
# the particles travel to their ranks (2 processes drive ranks 0..1 here: destination modulo world)
cols = np.stack([p.x[mine], p.y[mine], p.z[mine], mine.astype(np.float64)], axis=1)
got = domain.orb_exchange(cols, dest % world, "cpu")
want = np.sort(np.concatenate([doms[r] for r in range(4) if r % world == rank]))
assert np.array_equal(np.sort(got[:, 3].astype(np.int64)), want)
assert np.array_equal(got[:, 0], p.x[got[:, 3].astype(np.int64)])
dist.destroy_process_group()
open(os.path.join(os.environ["GG_TEST_OUT"], f"rank{{rank}}.ok"), "w").write("ok")
"""


def test_driver_across_processes_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)],
                       capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, GG_TEST_OUT=str(tmp_path)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()


_WORKER_SHARE = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
from gasoline_b200 import domain, ics
from oracle import orb_oracle
import orb_stub
# CPU stand-in for the device services behind device_orb_share (test infrastructure): same interface as PKD
class _Svc(orb_stub.HostOrbRank):
    def __init__(self, device=None, fPeriod=None):
        pass
    def pkdOrbLoad(self, x, y, z, fWeight=None):
        orb_stub.HostOrbRank.__init__(self, x, y, z, fWeight)
    def close(self):
        pass
    def commInitNccl(self, ident, rank, world):  # no GPU here: the library communicator cannot be made on rank 1
        if rank == 1:
            raise domain._pkd.GasolineB200Error("gg_comm_init failed (test stand-in)")
domain.PKD = _Svc
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
p = ics.plummer(3001, seed=8)
t = {{}}
idx = domain.device_orb_share(p, rank, world, None, "cpu", timing=t)
doms, _ = orb_oracle.domain_decomp(p.x, p.y, p.z, world)
assert np.array_equal(idx, doms[rank]) and t["trials"] > 0 and t["collective"] is False
# collective asked for, but one rank cannot join the library communicator: ALL ranks agree on the host path
t2 = {{}}
idx2 = domain.device_orb_share(p, rank, world, None, "cpu", timing=t2, collective=True)
assert np.array_equal(idx2, doms[rank]) and t2["collective"] is False and t2["trials"] == t["trials"]
dist.destroy_process_group()
open(os.path.join(os.environ["GG_TEST_OUT"], f"rank{{rank}}.ok"), "w").write("ok")
"""


def test_orb_share_across_processes_gloo_world2(tmp_path):
    """device_orb_share's host logic (chunking, reduce, destinations, all-to-all) on two gloo processes, the device
    services replaced by the numpy stand-in: every process ends up with exactly its domain of the oracle's decomposition."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER_SHARE.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29553", str(script)],
                       capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, GG_TEST_OUT=str(tmp_path)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()


def test_repeated_decomposition_hysteresis_and_shortcuts():
    """Later decompositions of a run (restated from pst.c:1900-1910, :963; not pinned by execution): the driver equals
    the oracle; the split axis survives a small change of shape and flips on a large one; bDoRootFind = 0 keeps a split
    that is still inside the bounds."""
    p = ics.plummer(5000, seed=17)
    doms0, nodes0 = orb_oracle.domain_decomp(p.x, p.y, p.z, 4)
    prev = {n[0]: (n[1], n[2]) for n in nodes0}
    # the particles move a little: same axes, new roots
    rng = np.random.default_rng(5)
    x, y, z = (c + rng.normal(0, 1e-3, p.n) for c in (p.x, p.y, p.z))
    doms1, nodes1 = orb_oracle.domain_decomp(x, y, z, 4, prev=prev)
    assert [n[1] for n in nodes1] == [n[1] for n in nodes0]
    for nService in (1, 3):
        owner = np.arange(p.n) % nService
        idx = [np.nonzero(owner == s)[0] for s in range(nService)]
        first = [HostOrbRank(p.x[i], p.y[i], p.z[i]) for i in idx]
        got0 = domain.pst_domain_decomp(first, 4)
        ranks = [HostOrbRank(x[i], y[i], z[i]) for i in idx]
        got1 = domain.pst_domain_decomp(ranks, 4, prev=got0)
        dest = np.zeros(p.n, np.int32)
        for i, r in zip(idx, ranks):
            dest[i] = domain.leaf_rank(4)[r.pkdOrbCells()]
        for r in range(4):
            assert np.array_equal(np.nonzero(dest == r)[0], doms1[r])
        ref = {n[0]: n for n in nodes1}
        assert all(c["iDim"] == ref[c["iCell"]][1] and c["fSplit"] == ref[c["iCell"]][2] for c in got1)
        # bDoRootFind = bDoSplitDimFind = 0: the old axes and splits are kept (they are still inside the bounds)
        ranks = [HostOrbRank(x[i], y[i], z[i]) for i in idx]
        got2 = domain.pst_domain_decomp(ranks, 4, prev=got0, bDoRootFind=False, bDoSplitDimFind=False)
        assert [(c["iDim"], c["fSplit"]) for c in got2] == [(c["iDim"], c["fSplit"]) for c in got0]
        d2, n2 = orb_oracle.domain_decomp(x, y, z, 4, prev=prev, do_root_find=False, do_split_dim_find=False)
        assert [(n[1], n[2]) for n in sorted(n2, key=lambda n: n[0])] == [(c["iDim"], c["fSplit"]) for c in got2]
    # the axis loop as written (pst.c:1900-1910) starts from 0.707 x the previous axis's extent but visits all three axes,
    # the previous one included, so it always ends on the first axis of strictly largest extent
    for ext in ((1.0, 1.2, 0.5), (1.2, 1.0, 0.5), (0.8, 0.75, 1.0), (1.0, 1.0, 1.0)):
        lo, hi = np.zeros(3), np.array(ext)
        for pd in (-1, 0, 1, 2):
            assert orb_oracle.split_dim(lo, hi, pd) == int(np.argmax(ext))


def _reference_domains_live(p, nThreads, tmp):
    """Run the compiled reference binary on nThreads pthread-MDL ranks (nSteps = 0) and return each rank's particles."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_multirank import parse_dump
    from oracle import reflib
    ics.write_tipsy_native(os.path.join(tmp, "ic.tipsy"), p)
    open(os.path.join(tmp, "run.param"), "w").write(
        f"achInFile = {tmp}/ic.tipsy\nachOutName = {tmp}/out\nbPeriodic = 0\ndTheta = 0.7\nnSteps = 0\nbVStep = 1\n"
        "bDoDensity = 0\niBinaryOutput = 0\nbParaRead = 0\nbParaWrite = 0\nbOverwrite = 1\n")
    env = dict(os.environ, MDL_NTHREADS=str(nThreads), REF_DUMP=os.path.join(tmp, "dump"))
    try:  # (what the binary does after the force evaluation is of no interest; the dumps are complete by then)
        subprocess.run([reflib.BIN_PATH, "run.param"], cwd=tmp, env=env, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, timeout=60)
    except subprocess.TimeoutExpired:
        pass
    return [np.sort(parse_dump(os.path.join(tmp, f"dump.rank{r}"))["iOrder"]) for r in range(nThreads)]


@pytest.mark.parametrize("n,nThreads,seed", [(4000, 5, 31), (5000, 7, 33), (3500, 8, 32), (2999, 6, 34)])
def test_orb_oracle_vs_reference_binary_live(n, nThreads, seed, tmp_path):
    """Build container only: rank counts beyond the committed fixtures (5-8 ranks: uneven nLower/nUpper on several
    levels, domain sizes that do not divide) -- the restatement and the driver reproduce the domains of the reference
    binary run live."""
    from oracle import reflib
    if not os.path.exists(reflib.BIN_PATH):
        pytest.skip("oracle/_ref/gasoline_ref not built (needs /root/reference)")
    p = ics.plummer(n, seed=seed)
    ref = _reference_domains_live(p, nThreads, str(tmp_path))
    doms, _ = orb_oracle.domain_decomp(p.x, p.y, p.z, nThreads)
    nodes, dest = _run_driver(p, nThreads, 3)
    for r in range(nThreads):
        assert np.array_equal(ref[r], doms[r]), f"rank {r}"
        assert np.array_equal(np.nonzero(dest == r)[0], ref[r])


# ---- later decompositions of a run, PINNED: a time-stepping multi-rank run of the reference binary (make_golden_orbsteps.py)
import sys as _sys  # noqa: E402

_sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_orbsteps import CASES as ORBSTEP_CASES  # noqa: E402


@pytest.mark.parametrize("name", sorted(ORBSTEP_CASES))
def test_later_decompositions_match_the_stepping_reference(name):
    """Step 0 = the first decomposition (unit weights, no previous split).  Steps >= 1 = what pstDomainDecomp does on every
    later call of a single-rung run: bounds of the moved particles, split axis by the NEWSPLITDIMCUT hysteresis against the
    previous axis (pst.c:1900-1910), bisection from the new bounds with the work weights pkdGravAll left in fWeight
    (bSplitWork, master.c:964).  The *_overflow runs have small particle stores: where a split sends a side more than its
    ranks' stores hold the reference bisects a second boundary into the cell and the lower ranks receive the wrapped
    interval between the two (pst.c:1049-1270) -- reproduced when the restatements are given the ranks' stores.  Oracle and
    device-service driver (host stub of the services) against the reference's particle-to-rank assignment, particle for
    particle."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    nThreads, nSteps = int(z["nThreads"]), int(z["nSteps"])
    overflow = name.endswith("_overflow")
    stores = orb_oracle.rank_stores(nThreads, len(z["s0_pos"]), float(z["dExtraStore"])) if overflow else None
    if overflow:
        assert np.array_equal(stores, domain.rank_stores(nThreads, len(z["s0_pos"]), float(z["dExtraStore"])))
    prev_o, prev_d = None, None
    nFixed = nDiffer = 0
    for k in range(nSteps + 1):
        pos, want = z[f"s{k}_pos"], z[f"s{k}_rank"]
        w = None if k == 0 else z[f"s{k - 1}_fWeight"]
        doms, nodes = orb_oracle.domain_decomp(pos[:, 0], pos[:, 1], pos[:, 2], nThreads, weights=w, prev=prev_o, stores=stores)
        got = np.full(len(pos), -1, np.int32)
        for r, ix in enumerate(doms):
            got[ix] = r
        nbad = int(np.count_nonzero(got != want))
        assert nbad == 0, f"{name} decomposition {k}: {nbad} particles on another rank than in the reference run"
        if overflow:
            assert np.all(np.bincount(want, minlength=nThreads) <= stores)
            nFixed += sum(1 for n in nodes if n[5])
            if any(n[5] for n in nodes):
                # without the stores the plain split (r[d] < fSplit) decomposes differently: some rank gets more than it holds
                plain, _ = orb_oracle.domain_decomp(pos[:, 0], pos[:, 1], pos[:, 2], nThreads, weights=w,
                                                    prev={c: v[:2] for c, v in prev_o.items()} if prev_o else None)
                assert any(not np.array_equal(a, b) for a, b in zip(plain, doms))
                nDiffer += 1
        prev_o = {n[0]: tuple(n[i] for i in ((1, 2, 4) if overflow else (1, 2))) for n in nodes}
        # the driver above the per-rank services (what bench.py / a non-Gasoline host runs), services on 2 host stubs
        owner = np.arange(len(pos)) % 2
        idx = [np.nonzero(owner == s)[0] for s in range(2)]
        ranks = [HostOrbRank(pos[i, 0], pos[i, 1], pos[i, 2], None if w is None else w[i]) for i in idx]
        cells = domain.pst_domain_decomp(ranks, nThreads, prev=prev_d, stores=stores)
        dest = np.zeros(len(pos), np.int32)
        for i, r in zip(idx, ranks):
            dest[i] = domain.leaf_rank(nThreads)[r.pkdOrbCells()]
        assert np.array_equal(dest, want), f"{name} decomposition {k}: the driver's domains differ from the reference's"
        if overflow:
            by = {n[0]: n for n in nodes}
            for c in cells:
                assert c["fSplitInactive"] == by[c["iCell"]][4] and c["fixed"] == by[c["iCell"]][5]
        prev_d = cells
    if overflow:
        assert nFixed >= 1 and nDiffer >= 1, "the fixture no longer enters the store-overflow branch"
    if nSteps >= 1:  # the work weights really moved the boundaries
        assert not np.array_equal(np.bincount(z["s0_rank"]), np.bincount(z["s1_rank"]))


def test_a_store_overflow_is_refused_by_ranks_without_a_wrap_split():
    """Ranks whose services have no pkdOrbSplitWrap (a host's own): pst_domain_decomp with `stores` runs the reverse split's
    counting on them and, when a boundary really moves into a cell, says so instead of decomposing differently from the
    reference.  (Here: the host stand-in with the method hidden; the device services have gg_orb_split_wrap.)"""
    from gasoline_b200.pkd import GasolineB200Error
    name = "orbsteps_plummer1800_r2_overflow"
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    nThreads = int(z["nThreads"])
    stores = domain.rank_stores(nThreads, len(z["s0_pos"]), float(z["dExtraStore"]))

    class NoWrap(HostOrbRank):
        def __getattribute__(self, a):
            if a == "pkdOrbSplitWrap":
                raise AttributeError(a)
            return super().__getattribute__(a)

    pos, w = z["s1_pos"], z["s0_fWeight"]
    with pytest.raises(GasolineB200Error, match="stores"):
        domain.pst_domain_decomp([NoWrap(pos[:, 0], pos[:, 1], pos[:, 2], w)], nThreads, stores=stores)
    # the first decomposition of the run fits: same call, no complaint, the reference's domains
    pos0 = z["s0_pos"]
    r0 = NoWrap(pos0[:, 0], pos0[:, 1], pos0[:, 2])
    domain.pst_domain_decomp([r0], nThreads, stores=stores)
    assert np.array_equal(domain.leaf_rank(nThreads)[r0.pkdOrbCells()], z["s0_rank"])
