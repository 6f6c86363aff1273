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


@pytest.mark.parametrize("packed", [False, True], ids=["host-arrays", "device-records"])
@pytest.mark.parametrize("name", NAMES)
def test_multidomain_matches_multirank_reference(name, packed, gpu_lib):
    p, theta, nThreads, z = load(name)
    doms = make_domains(p, theta, nThreads, z, device=0)
    domain.run_in_process(doms, packed=packed)
    g = GravityParams(nReps=1, bPeriodic=1, bEwald=1) if p.periodic else GravityParams(nReps=0, bPeriodic=0, bEwald=0)
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
