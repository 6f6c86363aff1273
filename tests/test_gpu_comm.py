"""GPU: the exchange BELOW the C ABI (csrc/gg_comm.cu: gg_comm_init_local / gg_comm_allgather / gg_exchange).  The ranks
are threads of this process sharing the one GPU of the test box through an in-process group -- the same collective
sequence a multi-process NCCL job runs (bench.py --gpus N; tests/test_dist_nccl.py covers the NCCL transport on a
multi-GPU box).  Checked against MULTI-RANK runs of the reference binary (per-bucket list counts bit-exact, results
within tolerance) and against the single-threaded driver of the same exchange (bit-identical forces)."""
import numpy as np
import pytest

from gasoline_b200 import domain, ics
from gasoline_b200.pkd import GravityParams
from multirank_cases import NAMES, load, make_domains
from parity import MAX_TOL, RMS_TOL

pytestmark = pytest.mark.gpu


def _params(p):
    return GravityParams(nReps=1, bPeriodic=1, bEwald=1) if p.periodic else GravityParams(nReps=0, bPeriodic=0, bEwald=0)


@pytest.mark.parametrize("name", NAMES)
def test_library_exchange_matches_multirank_reference(name, gpu_lib):
    p, theta, nThreads, z = load(name)
    doms = make_domains(p, theta, nThreads, z, device=0)
    g = _params(p)
    drivers = domain.run_threads(doms, g, rounds=2)  # twice: the second exchange must forget the first one's domains
    for r, d in enumerate(doms):
        info = d.pkd.commInfo()
        assert info["rank"] == r and info["nRanks"] == nThreads and info["transport"] == "group"
        st = drivers[r].stats
        assert st["bytesSent"] > 0 and st["bytesReceived"] > 0
        out = d.pkd.pkdGravAll(g)
        counts = d.pkd.pkdBucketCounts()
        bk = z[f"r{r}_buckets"]
        assert np.array_equal(counts[bk[:, 0]], bk[:, 3:6]), f"rank {r}: per-bucket list counts differ"
        assert (out["dPartSum"], out["dCellSum"], out["dSoftSum"], out["dFlop"]) == tuple(z[f"r{r}_sums"])
        res = z[f"r{r}_res"]
        rel = np.linalg.norm(out["acc"] - res[:, 0:3], axis=1) / np.linalg.norm(res[:, 0:3], axis=1)
        rms, mx = float(np.sqrt(np.mean(rel ** 2))), float(rel.max())
        print(f"{name} rank {r} (gg_exchange, group transport): acc rms {rms:.2e} max {mx:.2e}; "
              f"sent {st['bytesSent'] / 1e3:.0f} kB of a {st['bytesWholeDomain'] / 1e3:.0f} kB domain")
        assert rms <= RMS_TOL and mx <= MAX_TOL
        assert np.array_equal(out["fWeight"], res[:, 5])
    for d in doms:
        d.pkd.close()


@pytest.mark.parametrize("case", ["plummer200k_r4", "periodic32_r3", "plummer60k_r8"])
def test_library_exchange_equals_single_threaded_driver(case, gpu_lib):
    """gg_exchange (threads + group) against run_in_process's pruned-LET driver (gg_let_export / gg_set_remote_packed called
    rank by rank from one thread): same remote domains in the same order, hence bit-identical results."""
    if case == "plummer200k_r4":
        p, world, g = ics.plummer(200_000, seed=9), 4, GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    elif case == "plummer60k_r8":
        p, world, g = ics.plummer(60_000, seed=4), 8, GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    else:
        p, world, g = ics.periodic_box(32), 3, GravityParams(nReps=1, bPeriodic=1, bEwald=1)
    parts = domain.orb_decompose(p.x, p.y, p.z, world)
    results = {}
    for mode in ("single", "threads"):
        doms = [domain.Domain(r, world, p.x[ix], p.y[ix], p.z[ix], p.m[ix], p.h[ix], p.period, 0.7, device=0)
                for r, ix in enumerate(parts)]
        if mode == "single":
            domain.run_in_process(doms, let=g)
        else:
            domain.run_threads(doms, g)
        results[mode] = [(d.pkd.pkdGravAll(g), d.pkd.pkdBucketCounts()) for d in doms]
        for d in doms:
            d.pkd.close()
    for (a, ca), (b, cb) in zip(results["single"], results["threads"]):
        assert np.array_equal(ca, cb)
        for k in ("dPartSum", "dCellSum", "dSoftSum", "dFlop"):
            assert a[k] == b[k]
        assert np.array_equal(a["acc"], b["acc"]) and np.array_equal(a["pot"], b["pot"])
        assert np.array_equal(a["fWeight"], b["fWeight"])


def test_exchange_argument_checks(gpu_lib):
    from gasoline_b200.pkd import PKD, GasolineB200Error
    p = ics.plummer(2000, seed=1)
    pkd = PKD(device=0)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    pkd.pkdBuildBinary(8, 0.7, 4)
    pkd.upload()
    with pytest.raises(GasolineB200Error):  # no communicator
        pkd.pkdExchange(GravityParams(nReps=0, bPeriodic=0, bEwald=0))
    pkd.close()
