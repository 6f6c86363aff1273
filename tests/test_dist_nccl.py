"""Multi-GPU box only (skipped with fewer than 2 devices): the NCCL transport of the exchange below the C ABI.  One
PROCESS per GPU, no torch.distributed anywhere -- the 128-byte communicator id travels through a multiprocessing pipe
(the "host's own means"), everything else (top-tree all-gathers, sizes, pruned trees) through gg_comm_* / gg_exchange
over NCCL.  The results must equal, bit for bit, those of the same ranks run as threads of one process on GPU 0 through
the in-process group transport (which tests/test_gpu_comm.py pins to the multi-rank reference)."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")]


def _case(name):
    from gasoline_b200 import ics
    from gasoline_b200.pkd import GravityParams
    if name == "plummer":
        return ics.plummer(300_000, seed=9), GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    return ics.periodic_box(40), GravityParams(nReps=1, bPeriodic=1, bEwald=1)


def _rank_main(name, rank, world, ident, q):
    sys.path.insert(0, ROOT)
    from gasoline_b200 import domain
    p, g = _case(name)
    parts = domain.orb_decompose(p.x, p.y, p.z, world)
    ix = parts[rank]
    d = domain.Domain(rank, world, p.x[ix], p.y[ix], p.z[ix], p.m[ix], p.h[ix], p.period, 0.7, device=rank)
    d.pkd.commInitNccl(ident, rank, world)
    ex = domain.LibExchange(d)
    ex.exchange(g)
    ex.exchange(g)  # a second step: the first one's remote domains must be forgotten
    out = d.pkd.pkdGravAll(g)
    info = d.pkd.commInfo()
    q.put((rank, out["acc"], out["pot"], out["fWeight"], d.pkd.pkdBucketCounts(),
           (out["dPartSum"], out["dCellSum"], out["dSoftSum"], out["dFlop"]), info, ex.stats))
    d.pkd.close()


@pytest.mark.parametrize("name", ["plummer", "periodic"])
def test_nccl_exchange_equals_group_exchange(name, gpu_lib):
    from gasoline_b200 import domain
    from gasoline_b200.pkd import comm_unique_id
    world = min(_ngpu(), 8)
    ident = comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(name, r, world, ident, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = {}
    for _ in range(world):
        item = q.get(timeout=600)
        got[item[0]] = item[1:]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    # the same ranks as threads on one GPU
    p, g = _case(name)
    parts = domain.orb_decompose(p.x, p.y, p.z, world)
    doms = [domain.Domain(r, world, p.x[ix], p.y[ix], p.z[ix], p.m[ix], p.h[ix], p.period, 0.7, device=0)
            for r, ix in enumerate(parts)]
    domain.run_threads(doms, g)
    for r, d in enumerate(doms):
        out = d.pkd.pkdGravAll(g)
        acc, pot, fw, counts, sums, info, st = got[r]
        assert info["transport"] == "nccl" and info["nRanks"] == world and info["nccl_version"] > 0
        assert np.array_equal(counts, d.pkd.pkdBucketCounts())
        assert sums == (out["dPartSum"], out["dCellSum"], out["dSoftSum"], out["dFlop"])
        assert np.array_equal(acc, out["acc"]) and np.array_equal(pot, out["pot"]) and np.array_equal(fw, out["fWeight"])
        print(f"{name} rank {r}/{world} over NCCL {info['nccl_version']}: sent {st['bytesSent'] / 1e6:.2f} MB, export "
              f"{st['msExport']:.3f} ms, transfer {st['msTransfer']:.3f} ms, ingest {st['msIngest']:.3f} ms")
        d.pkd.close()
