"""GPU: the MULTI-RANK drop-in.  The reference's own binary (main.c, master.c, pst.c, ... compiled where they lie;
oracle/_ref/gasoline_ref_gpu) runs on k pthread-MDL ranks with its pkdGravAll link-substituted by the product's shim
(gasoline_b200/csrc/pkd_gravall_shim.c): the host's domain decomposition, tree builds and top tree are the reference's,
every rank's force evaluation runs through the C ABI, and the remote trees travel below it (gg_exchange; the ranks
share the test box's one GPU through the in-process group transport, GG_SHIM_COMM=local).  Compared with the dumps the
PURE reference binary produced on the same ranks (tests/golden/multirank_*.npz): the particles of every rank and the
top tree identical, per-bucket interaction-list counts and the sums bit-exact, results within the north-star tolerance."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from gasoline_b200 import ics
from multirank_cases import NAMES, load
from oracle import reflib
from parity import MAX_TOL, RMS_TOL

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_multirank import parse_dump  # noqa: E402

GPU_BIN = os.path.join(os.path.dirname(reflib.BIN_PATH), "gasoline_ref_gpu")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(GPU_BIN), reason="oracle/_ref/gasoline_ref_gpu not built")]


def run_host(p, theta, nThreads, tmp, extra_env=None):
    ics.write_tipsy_native(os.path.join(tmp, "ic.tipsy"), p)
    periodic = 1 if p.periodic else 0
    open(os.path.join(tmp, "run.param"), "w").write(
        f"achInFile = {tmp}/ic.tipsy\nachOutName = {tmp}/out\nbPeriodic = {periodic}\ndPeriod = 1\n"
        f"nReplicas = {periodic}\nbEwald = {periodic}\ndTheta = {theta}\nnSteps = 0\nbVStep = 1\n"
        "bDoDensity = 0\niBinaryOutput = 0\nbParaRead = 0\nbParaWrite = 0\nbOverwrite = 1\n")
    env = dict(os.environ, MDL_NTHREADS=str(nThreads), REF_DUMP=os.path.join(tmp, "dump"), GG_SHIM_COMM="local")
    env.update(extra_env or {})
    # (like the pure reference run that made the fixtures, the binary's exit status after the force evaluation -- it
    #  re-orders particles for output through an mdlSwap the MDL stand-in only half supports -- is of no interest)
    r = subprocess.run([GPU_BIN, "run.param"], cwd=tmp, env=env, capture_output=True, text=True, timeout=600)
    dumps = []
    for k in range(nThreads):
        path = os.path.join(tmp, f"dump.rank{k}")
        assert os.path.exists(path), f"rank {k} wrote no dump\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
        dumps.append(parse_dump(path))
    return dumps, r


@pytest.mark.parametrize("name", NAMES)
def test_multirank_reference_host_on_the_gpu(name, gpu_lib):
    p, theta, nThreads, z = load(name)
    with tempfile.TemporaryDirectory() as tmp:
        dumps, r = run_host(p, theta, nThreads, tmp)
    assert "Gravity Calculated" in r.stdout
    for k, d in enumerate(dumps):
        assert d["idSelf"] == k and d["nThreads"] == nThreads
        assert np.array_equal(d["iOrder"], z[f"r{k}_iOrder"]), "the host's decomposition / tree order changed"
        assert np.array_equal(d["top_i"], z["top_i"]) and np.array_equal(d["top_d"], z["top_d"])
        ref_b, got_b = z[f"r{k}_buckets"], d["buckets"]
        ref_map = {int(b[0]): tuple(b[1:6]) for b in ref_b}
        got_map = {int(b[0]): tuple(b[1:6]) for b in got_b}
        assert ref_map == got_map, f"rank {k}: per-bucket interaction-list counts differ from the reference's"
        assert tuple(d["sums"]) == tuple(z[f"r{k}_sums"])
        res, ref = d["res"], z[f"r{k}_res"]
        rel = np.linalg.norm(res[:, 0:3] - ref[:, 0:3], axis=1) / np.linalg.norm(ref[:, 0:3], axis=1)
        rms, mx = float(np.sqrt(np.mean(rel ** 2))), float(rel.max())
        floor = np.sqrt(np.mean(ref[:, 3] ** 2))
        dp = np.abs(res[:, 3] - ref[:, 3]) / np.maximum(np.abs(ref[:, 3]), floor)
        print(f"{name} rank {k} (reference binary on {nThreads} ranks + GPU pkdGravAll): acc rms {rms:.2e} max {mx:.2e}; "
              f"pot max {dp.max():.2e}")
        assert rms <= RMS_TOL and mx <= MAX_TOL
        assert np.sqrt(np.mean(dp ** 2)) <= RMS_TOL and dp.max() <= MAX_TOL
        assert (np.abs(res[:, 4] - ref[:, 4]) / ref[:, 4]).max() <= MAX_TOL
        assert np.array_equal(res[:, 5], ref[:, 5])  # fWeight


def test_multirank_reference_host_device_tree(gpu_lib):
    """The same with every rank's pkdBuildBinary ALSO served by its GPU (GG_SHIM_DEVICE_TREE=1)."""
    name = "multirank_plummer3000_r4"
    p, theta, nThreads, z = load(name)
    with tempfile.TemporaryDirectory() as tmp:
        dumps, r = run_host(p, theta, nThreads, tmp, {"GG_SHIM_DEVICE_TREE": "1"})
    for k, d in enumerate(dumps):
        assert np.array_equal(d["iOrder"], z[f"r{k}_iOrder"])
        ref_map = {int(b[0]): tuple(b[1:6]) for b in z[f"r{k}_buckets"]}
        got_map = {int(b[0]): tuple(b[1:6]) for b in d["buckets"]}
        assert ref_map == got_map
        assert tuple(d["sums"]) == tuple(z[f"r{k}_sums"])
        ref = z[f"r{k}_res"]
        rel = np.linalg.norm(d["res"][:, 0:3] - ref[:, 0:3], axis=1) / np.linalg.norm(ref[:, 0:3], axis=1)
        assert float(np.sqrt(np.mean(rel ** 2))) <= RMS_TOL and float(rel.max()) <= MAX_TOL
