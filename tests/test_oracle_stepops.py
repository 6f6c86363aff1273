"""oracle_step_ops (pkdKick / pkdDrift / pkdGravStep restated, oracle/gravity_oracle.c) against the golden vectors made
by the compiled reference (tests/golden/make_golden_stepops.py) and, where /root/reference exists, against the
reference itself -- bit for bit."""
import os
import sys

import numpy as np
import pytest

from oracle import reflib
from oracle.oracle import DRIFT, GRAVSTEP, KICK, oracle_step_ops

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_stepops import PARAMS, inputs  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stepops.npz"))


@pytest.mark.parametrize("name,what", [("kick", KICK), ("drift", DRIFT), ("gravstep", GRAVSTEP), ("all", KICK | DRIFT | GRAVSTEP)])
def test_oracle_step_ops_golden(name, what):
    r, v, a, active, dtGrav, dt = inputs()
    r2, v2, dt2, nOut = oracle_step_ops(r, v, a, active, dtGrav, dt, what=what, **PARAMS)
    assert nOut == 0
    assert np.array_equal(r2, GOLD[name + "_r"]) and np.array_equal(v2, GOLD[name + "_v"]) and np.array_equal(dt2, GOLD[name + "_dt"])


def test_oracle_drift_counts_runaways():
    r, v, a, active, dtGrav, dt = inputs(n=64)
    v[5, 1] = -80.0
    _, _, _, nOut = oracle_step_ops(r, v, a, None, dtGrav, dt, what=DRIFT, **PARAMS)
    assert nOut == 1


@pytest.mark.skipif(not (reflib.available() and os.path.exists("/root/reference/pkd.c")), reason="compiled reference not present")
def test_oracle_step_ops_vs_reference_random():
    for seed in (1, 2, 3):
        r, v, a, active, dtGrav, dt = inputs(seed=seed, n=1000)
        kw = dict(PARAMS, dDelta=0.01 * seed, dvFacTwo=0.003 * seed)
        for what in (KICK, DRIFT, GRAVSTEP, KICK | DRIFT | GRAVSTEP):
            o = oracle_step_ops(r, v, a, active, dtGrav, dt, what=what, **kw)
            f = reflib.ref_step_ops(r, v, a, active, dtGrav, dt, what=what, **kw)
            assert all(np.array_equal(x, y) for x, y in zip(o[:3], f[:3]))


# ---- time-step selection and rungs (pkdInitDt, pkdAccelStep, pkdGravStep, pkdDtToRung, pkdActiveRung)
import make_golden_rungops as _rg  # noqa: E402
from oracle.oracle import oracle_rung_ops  # noqa: E402

RGOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rungops.npz"))


@pytest.mark.parametrize("name", list(_rg.CASES))
def test_oracle_rung_ops_golden(name):
    what, kw = _rg.CASES[name]
    act, dt, rung, out = oracle_rung_ops(*_rg.inputs(), what=what, **kw)
    assert np.array_equal(act, RGOLD[name + "_active"]) and np.array_equal(dt, RGOLD[name + "_dt"])
    assert np.array_equal(rung, RGOLD[name + "_rung"]) and np.array_equal(out, RGOLD[name + "_out"])


@pytest.mark.skipif(not (reflib.available() and os.path.exists("/root/reference/pkd.c")), reason="compiled reference not present")
def test_oracle_rung_ops_vs_reference_random():
    for seed in (5, 6):
        args = _rg.inputs(seed=seed, n=1500)
        for name, (what, kw) in _rg.CASES.items():
            kw = dict(kw, dEta=0.05 * seed)
            o = oracle_rung_ops(*args, what=what, **kw)
            f = reflib.ref_rung_ops(*args, what=what, **kw)
            assert all(np.array_equal(x, y) for x, y in zip(o, f)), name
