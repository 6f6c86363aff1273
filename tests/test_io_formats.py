"""On-disk formats either side of the path: native Tipsy in, .accg / .pot / .dt array files out
(ASCII %.14g -- vectors particle by particle -- and binary double -- vectors component by component), checked against
files the REFERENCE BINARY itself writes (oracle/_ref/gasoline_ref run
with nSteps = 0, the diagnostic branch main.c:621-657): our readers parse them, our writers reproduce them byte for
byte, and their contents are the reference's forces (== the oracle's)."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

from gasoline_b200 import ics
from oracle import oracle, reflib

pytestmark = pytest.mark.skipif(not os.path.exists(reflib.BIN_PATH), reason="oracle/_ref/gasoline_ref not built")


def _run_reference(tmp, p, mode):
    ics.write_tipsy_native(os.path.join(tmp, "ic.tipsy"), p)
    open(os.path.join(tmp, "run.param"), "w").write(
        f"achInFile = ic.tipsy\nachOutName = out{mode}\nbPeriodic = 0\nnReplicas = 0\nbEwald = 0\ndTheta = 0.7\n"
        f"nSteps = 0\nbVStep = 1\nbDoDensity = 0\niBinaryOutput = {mode}\nbParaRead = 0\nbParaWrite = 0\nbOverwrite = 1\n")
    subprocess.run([reflib.BIN_PATH, "run.param"], cwd=tmp, env=dict(os.environ, MDL_NTHREADS="1"),
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300, check=True)


def test_tipsy_roundtrip(tmp_path):
    p = ics.plummer(777, seed=5)
    f = str(tmp_path / "a.tipsy")
    ics.write_tipsy_native(f, p)
    assert os.path.getsize(f) == 32 + 36 * p.n  # header + dark records (tipsydefs.h:17-23,38-45)
    q = ics.read_tipsy_native(f)
    for k in ("x", "y", "z", "m", "h"):
        assert np.array_equal(getattr(q, k), getattr(p, k))  # (the generators round to float32 like the file does)


@pytest.mark.parametrize("mode", [0, 2])
def test_array_files_match_the_reference_binary(tmp_path, mode):
    p = ics.plummer(600, seed=3)
    tmp = str(tmp_path)
    _run_reference(tmp, p, mode)
    rd = ics.read_array_ascii if mode == 0 else ics.read_array_binary
    wr = ics.write_array_ascii if mode == 0 else ics.write_array_binary
    acc = rd(os.path.join(tmp, f"out{mode}.accg"), 3)
    pot = rd(os.path.join(tmp, f"out{mode}.pot"), 1)[:, 0]
    dt = rd(os.path.join(tmp, f"out{mode}.dt"), 1)[:, 0]
    assert acc.shape == (p.n, 3) and pot.shape == (p.n,) and dt.shape == (p.n,)
    # our writers reproduce the reference's files byte for byte
    for name, a in (("accg", acc), ("pot", pot), ("dt", dt)):
        mine = os.path.join(tmp, f"mine.{name}")
        wr(mine, a)
        assert filecmp.cmp(mine, os.path.join(tmp, f"out{mode}.{name}"), shallow=False), name
    # and the content is the reference's force field in iOrder order (== the oracle's, in tree order)
    o = oracle.OracleGravity(p)
    o.build_tree(8, 0.7, 4)
    ref = o.gravity(0, 0, 4, 0, 4, 2.6, 2.8)
    order = o.tree()["iOrder"]
    o.close()
    a_ref = np.zeros((p.n, 3)); a_ref[order] = ref["acc"]
    p_ref = np.zeros(p.n); p_ref[order] = ref["pot"]
    # (the oracle's 1/sqrt differs from the reference's table-based v_sqrt1 by ~2e-10; %.14g keeps 14 digits)
    assert np.allclose(acc, a_ref, rtol=1e-8, atol=2e-9) and np.allclose(pot, p_ref, rtol=1e-8, atol=0)


def test_standard_tipsy_is_read_by_the_reference_like_the_native_file(tmp_path):
    """bStandard = 1 (XDR, big-endian; xdrHeader master.c:3563, pkdReadTipsy pkd.c:456-560): the reference binary reads
    our standard file and writes the same force files, byte for byte, as from our native file of the same particles;
    our reader inverts our writer."""
    p = ics.plummer(500, seed=8)
    tmp = str(tmp_path)
    f = os.path.join(tmp, "ic.std")
    ics.write_tipsy_standard(f, p)
    assert os.path.getsize(f) == 32 + 36 * p.n
    q = ics.read_tipsy_standard(f)
    for k in ("x", "y", "z", "m", "h"):
        assert np.array_equal(getattr(q, k), getattr(p, k))
    ics.write_tipsy_native(os.path.join(tmp, "ic.tipsy"), p)
    for name, infile, std in (("nat", "ic.tipsy", 0), ("std", "ic.std", 1)):
        open(os.path.join(tmp, f"{name}.param"), "w").write(
            f"achInFile = {infile}\nachOutName = out_{name}\nbStandard = {std}\nbPeriodic = 0\nnReplicas = 0\nbEwald = 0\n"
            "dTheta = 0.7\nnSteps = 0\nbVStep = 1\nbDoDensity = 0\niBinaryOutput = 0\nbParaRead = 0\nbParaWrite = 0\nbOverwrite = 1\n")
        subprocess.run([reflib.BIN_PATH, f"{name}.param"], cwd=tmp, env=dict(os.environ, MDL_NTHREADS="1"),
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300, check=True)
    for ext in ("accg", "pot"):
        assert filecmp.cmp(os.path.join(tmp, f"out_nat.{ext}"), os.path.join(tmp, f"out_std.{ext}"), shallow=False), ext
    acc = ics.read_array_ascii(os.path.join(tmp, "out_std.accg"), 3)
    assert acc.shape == (p.n, 3) and np.isfinite(acc).all() and np.abs(acc).max() > 0
