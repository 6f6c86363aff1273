"""CPU: the oracle's multi-rank walk (oracle/multidomain.py: the ranks' trees hung under the top tree = one walk of one
combined tree) pinned to MULTI-RANK runs of the reference binary (tests/golden/multirank_*.npz): per-bucket
interaction-list counts of every rank bit-exact, the sums pkdGravAll returns bit-exact, forces to the v_sqrt1 tolerance."""
import numpy as np
import pytest

from gasoline_b200 import domain
from multirank_cases import NAMES, load, make_domains
from oracle import multidomain, oracle


@pytest.mark.parametrize("name", NAMES)
def test_combined_tree_oracle_equals_multirank_reference(name):
    p, theta, nThreads, z = load(name)
    doms = make_domains(p, theta, nThreads, z)
    domain.run_in_process(doms, exchange_trees=False)
    act = {d.idSelf: np.ones(d.host.nLocal, np.int32) for d in doms}
    T, nodeBase, partBase = multidomain.from_domains(doms, act)
    o = oracle.OracleGravity(None, tree=T)
    per = 1 if p.periodic else 0
    out = o.gravity(per, per, 4, per, 4)
    o.close()
    for r in range(nThreads):
        bk = z[f"r{r}_buckets"]
        got = out["counts"][nodeBase[r] + bk[:, 0]]
        assert np.array_equal(got, bk[:, 3:6]), f"{name} rank {r}: per-bucket list counts differ from the reference's"
        n = len(z[f"r{r}_iOrder"])
        sl = slice(partBase[r], partBase[r] + n)
        res = z[f"r{r}_res"]
        rel = np.linalg.norm(out["acc"][sl] - res[:, 0:3], axis=1) / np.linalg.norm(res[:, 0:3], axis=1)
        # (the reference's default build uses the v_sqrt1 approximation, <= 2.1e-10 per term; the oracle exact 1/sqrt --
        #  the tolerance of tests/test_oracle_golden.py)
        assert rel.max() < 2e-7, rel.max()
        scale = np.sqrt(np.mean(res[:, 3] ** 2))
        assert np.abs(out["pot"][sl] - res[:, 3]).max() < 1e-7 * scale
        assert np.array_equal(out["fWeight"][sl], res[:, 5])
    # the sums of all ranks together (the oracle counts every sink of the combined tree)
    tot = np.sum([z[f"r{r}_sums"] for r in range(nThreads)], axis=0)
    assert (out["dPartSum"], out["dCellSum"], out["dSoftSum"]) == tuple(tot[:3])
