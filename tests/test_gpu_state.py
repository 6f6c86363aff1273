"""The device-resident particle store and the steps either side of the force path:
gg_state_kick / gg_state_drift / gg_state_gravstep against the reference's own pkdKick / pkdDrift / pkdGravStep
(golden vectors generated from the compiled reference, tests/golden/stepops.npz) and against the oracle restatement --
bit for bit -- and a kick-drift-kick integration on the resident store against the same integration with the oracle's
FP64 forces (tolerance: the force tolerance, amplified by the steps taken)."""
import os
import sys

import numpy as np
import pytest

from gasoline_b200 import ics
from gasoline_b200.pkd import PKD, GasolineB200Error, GravityParams
from oracle import oracle
from oracle.oracle import DRIFT, GRAVSTEP, KICK, oracle_step_ops

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_stepops import PARAMS, inputs  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stepops.npz"))


def _resident(r, v, active, period=(1.0, 1.0, 1.0), dt0=0.05):
    n = r.shape[0]
    pkd = PKD(fPeriod=period)
    pkd.pkdLoadResident(r[:, 0], r[:, 1], r[:, 2], v[:, 0], v[:, 1], v[:, 2], np.full(n, 1.0 / n), np.full(n, 0.01),
                        active, dt0=dt0)
    return pkd


def test_kick_drift_bit_exact_vs_reference_golden(gpu_lib):
    r, v, a, active, dtGrav, dt = inputs()
    P = PARAMS
    pkd = _resident(r, v, active)
    pkd.pkdKick(P["dvFacOne"], P["dvFacTwo"], a)
    s = pkd.pkdFetchResident()
    assert np.array_equal(s["v"], GOLD["kick_v"]) and np.array_equal(s["r"], r)
    assert np.array_equal(s["iOrder"], np.arange(r.shape[0]))
    pkd.close()
    pkd = _resident(r, v, active)
    pkd.pkdDrift(P["dDelta"], P["fCenter"], 1)
    s = pkd.pkdFetchResident()
    assert np.array_equal(s["r"], GOLD["drift_r"]) and np.array_equal(s["v"], v)
    assert np.any(np.abs(GOLD["drift_r"] - (r + P["dDelta"] * v)) > 0.5), "the fixture must exercise the wrap"
    pkd.close()
    # kick then drift, open boundaries, every particle active
    pkd = _resident(r * 3.0, v, None)
    pkd.pkdKick(P["dvFacOne"], P["dvFacTwo"], a)
    pkd.pkdDrift(P["dDelta"], bPeriodic=0)
    s = pkd.pkdFetchResident()
    assert np.array_equal(s["r"], GOLD["open_r"]) and np.array_equal(s["v"], GOLD["open_v"])
    # the oracle restatement gives the same bits (it is pinned to the reference in the CPU suite)
    r2, v2, _, _ = oracle_step_ops(r * 3.0, v, a, None, dtGrav, dt, what=KICK | DRIFT, **dict(P, bPeriodic=0))
    assert np.array_equal(s["r"], r2) and np.array_equal(s["v"], v2)
    pkd.close()


def test_drift_rejects_runaway_particle(gpu_lib):
    r, v, a, active, _, _ = inputs(n=256)
    v[3, 0] = 100.0  # leaves the box by more than one period: the reference asserts (pkd.c:3754)
    pkd = _resident(r, v, None)
    with pytest.raises(GasolineB200Error):
        pkd.pkdDrift(0.05, (0.0, 0.0, 0.0), 1)
    pkd.close()


def test_sequence_guards(gpu_lib):
    p = ics.plummer(3000, seed=6)
    v = np.zeros((p.n, 3))
    pkd = PKD()
    pkd.pkdLoadResident(p.x, p.y, p.z, v[:, 0], v[:, 1], v[:, 2], p.m, p.h)
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    with pytest.raises(GasolineB200Error):
        pkd.pkdGravAll(g, download=False)  # no tree yet
    pkd.pkdBuildBinaryResident(8, 0.7)
    with pytest.raises(GasolineB200Error):
        pkd.pkdKick(1.0, 0.1)  # no forces for this order yet
    pkd.pkdGravAll(g, download=False)
    pkd.pkdKick(1.0, 0.1)
    pkd.pkdDrift(0.01)
    with pytest.raises(GasolineB200Error):
        pkd.pkdGravAll(g, download=False)  # moved since the build
    pkd.close()


def _cpu_kdk(p, v, g, theta, nSteps, dDelta, active=None):
    """kick-drift-kick with the oracle's forces and the oracle's kick/drift, particle order = input order"""
    r = np.stack([p.x, p.y, p.z], axis=1).copy()
    v = v.copy()
    dtg = np.zeros(p.n)

    def forces(r):
        q = ics.Particles(r[:, 0].copy(), r[:, 1].copy(), r[:, 2].copy(), p.m, p.h, p.period, "step")
        o = oracle.OracleGravity(q, active=active)
        o.build_tree(8, theta, 4)
        out = o.gravity(g.nReps, g.bPeriodic, g.iOrder, g.bEwald, g.iEwOrder, g.fEwCut, g.fEwhCut)
        t = o.tree()
        o.close()
        a = np.zeros((p.n, 3))
        a[t["iOrder"]] = out["acc"]
        d = np.zeros(p.n)
        d[t["iOrder"]] = out["dtGrav"]
        return a, d

    a, dtg = forces(r)
    for _ in range(nSteps):
        r, v, _, _ = oracle_step_ops(r, v, a, active, dtg, dtg, dvFacOne=1.0, dvFacTwo=0.5 * dDelta, dDelta=dDelta,
                                     fCenter=(0, 0, 0), bPeriodic=g.bPeriodic, fPeriod=p.period if g.bPeriodic else (1, 1, 1),
                                     what=KICK | DRIFT)
        a, dtg = forces(r)
        r, v, _, _ = oracle_step_ops(r, v, a, active, dtg, dtg, dvFacOne=1.0, dvFacTwo=0.5 * dDelta, what=KICK)
    return r, v, dtg, a


@pytest.mark.parametrize("name", ["plummer4000", "periodic12_ewald"])
def test_resident_kdk_matches_cpu_integration(name, gpu_lib):
    if name == "plummer4000":
        p, g, dDelta = ics.plummer(4000, seed=21), GravityParams(nReps=0, bPeriodic=0, bEwald=0), 0.01
        vs = 0.3
    else:
        p, g, dDelta = ics.periodic_box(12), GravityParams(nReps=1, bPeriodic=1, bEwald=1), 0.02
        vs = 0.05
    rng = np.random.default_rng(5)
    v0 = rng.normal(0, vs, size=(p.n, 3))
    nSteps = 4
    r_ref, v_ref, dtg_ref, a_ref = _cpu_kdk(p, v0, g, 0.7, nSteps, dDelta)

    pkd = PKD(fPeriod=p.period)
    pkd.pkdLoadResident(p.x, p.y, p.z, v0[:, 0], v0[:, 1], v0[:, 2], p.m, p.h)
    pkd.pkdBuildBinaryResident(8, 0.7)
    pkd.pkdGravAll(g, download=False)
    for _ in range(nSteps):
        pkd.pkdKick(1.0, 0.5 * dDelta)
        pkd.pkdDrift(dDelta, (0.0, 0.0, 0.0), g.bPeriodic)
        pkd.pkdBuildBinaryResident(8, 0.7)
        out = pkd.pkdGravAll(g, download=False)
        pkd.pkdKick(1.0, 0.5 * dDelta)
    dmin = pkd.pkdGravStep(0.2)
    s = pkd.pkdFetchResident()
    pkd.close()
    assert sorted(s["iOrder"].tolist()) == list(range(p.n))
    r_gpu = np.zeros_like(r_ref); v_gpu = np.zeros_like(v_ref); dt_gpu = np.zeros(p.n)
    r_gpu[s["iOrder"]], v_gpu[s["iOrder"]], dt_gpu[s["iOrder"]] = s["r"], s["v"], s["dt"]
    dr = r_gpu - r_ref
    if g.bPeriodic:
        dr -= np.round(dr)  # a particle within rounding of the face may wrap in one run and not the other
    # the two runs differ only through the forces (1e-5 rms / 1e-4 max relative, the north-star tolerance): measure the
    # differences against what the forces contributed -- the accumulated kicks and the displacement they caused
    T = nSteps * dDelta
    a_rms = np.sqrt(np.mean(np.sum(a_ref ** 2, axis=1)))
    scale_v = T * a_rms
    scale_r = 0.5 * T * T * a_rms
    er, ev = np.abs(dr).max() / scale_r, np.abs(v_gpu - v_ref).max() / scale_v
    print(f"{name}: {nSteps} KDK steps: max |dr| / (a T^2/2) {er:.2e}, max |dv| / (a T) {ev:.2e}, dt_min {dmin:.4g}")
    assert er < 1e-4 and ev < 1e-4
    # pkdGravStep: dt = min(dt0, dEta/sqrt(dtGrav)) -- against the CPU integration's dtGrav within the force tolerance
    dt_ref = np.minimum(1e30, 0.2 / np.sqrt(dtg_ref))
    assert np.allclose(dt_gpu, dt_ref, rtol=2e-4)
    assert dmin == dt_gpu.min()


def test_gravstep_bit_exact_on_device_dtgrav(gpu_lib):
    p = ics.plummer(5000, seed=13)
    rng = np.random.default_rng(1)
    active = (rng.random(p.n) < 0.5).astype(np.int32)
    v = np.zeros((p.n, 3))
    pkd = PKD()
    pkd.pkdLoadResident(p.x, p.y, p.z, v[:, 0], v[:, 1], v[:, 2], p.m, p.h, active, dt0=0.03)
    pkd.pkdBuildBinaryResident(8, 0.7)
    out = pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0))
    dmin = pkd.pkdGravStep(0.2)
    s = pkd.pkdFetchResident()
    act_tree = active[s["iOrder"]]
    dtg = np.where(act_tree != 0, out["dtGrav"], 1.0)
    _, _, dt_ref, _ = oracle_step_ops(s["r"], s["v"], np.zeros((p.n, 3)), act_tree, dtg, np.full(p.n, 0.03), dEta=0.2,
                                      what=GRAVSTEP)
    assert np.array_equal(s["dt"], dt_ref)
    assert dmin == dt_ref.min() and np.all(s["dt"][act_tree == 0] == 0.03)
    pkd.close()


def test_rungs_and_timestep_selection_bit_exact(gpu_lib):
    """pkdActiveRung -> tree build -> gravity on the active set -> pkdInitDt, pkdAccelStep, pkdGravStep, pkdDtToRung,
    pkdActiveRung, all on the resident store; against the oracle restatement (pinned to the compiled reference in the CPU
    suite) fed with the very a / fPot / dtGrav the GPU produced: every dt, rung, flag and counter identical."""
    from oracle.oracle import ACCELSTEP, ACTIVERUNG, DTTORUNG, GRAVSTEP_R, INITDT, oracle_rung_ops
    p = ics.plummer(6000, seed=31)
    rng = np.random.default_rng(8)
    v0 = rng.normal(0, 0.3, size=(p.n, 3))
    rung0 = rng.integers(0, 3, p.n).astype(np.int32)
    dDelta, dEta = 0.02, 0.1
    pkd = PKD()
    pkd.pkdLoadResident(p.x, p.y, p.z, v0[:, 0], v0[:, 1], v0[:, 2], p.m, p.h, dt0=0.5)
    pkd.pkdSetRungs(rung0)
    nAct = pkd.pkdActiveRung(1, 1)  # before the build: the flags travel with the particles through the partition
    assert nAct == int((rung0 >= 1).sum())
    pkd.pkdBuildBinaryResident(8, 0.7)
    out = pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0))
    assert out["nActive"] == nAct
    s0 = pkd.pkdFetchResident()
    rung_t, act_t = pkd.pkdFetchRungs()
    assert np.array_equal(rung_t, rung0[s0["iOrder"]]) and np.array_equal(act_t, (rung_t >= 1).astype(np.int32))
    pkd.pkdInitDt(dDelta)
    pkd.pkdAccelStep(dEta, bSqrtPhi=1)
    pkd.pkdGravStep(dEta)
    imax, nmax, ideal = pkd.pkdDtToRung(1, dDelta, 8)
    nAct2 = pkd.pkdActiveRung(2, 1)
    s = pkd.pkdFetchResident()
    rung_g, act_g = pkd.pkdFetchRungs()
    dtg = np.where(act_t != 0, out["dtGrav"], 1.0)
    act_o, dt_o, rung_o, o = oracle_rung_ops(s0["v"], out["acc"], out["pot"], p.h[s0["iOrder"]], dtg, act_t,
                                             np.full(p.n, 0.5), rung_t, dDelta=dDelta, dEta=dEta, bSqrtPhi=1, iRung=1,
                                             iMaxRung=8, bAll=1, iRungActive=2, bGreater=1,
                                             what=INITDT | ACCELSTEP | GRAVSTEP_R | DTTORUNG | ACTIVERUNG)
    assert np.array_equal(s["dt"], dt_o)
    assert np.array_equal(rung_g, rung_o) and np.array_equal(act_g, act_o)
    assert (imax, nmax, ideal, nAct2) == tuple(int(x) for x in o)
    assert len(np.unique(rung_g)) >= 4, "the case must spread the particles over several rungs"
    # the new active set is what the loaded domain now evaluates (gg_set_active inside pkdActiveRung)
    out2 = pkd.pkdGravAll(GravityParams(nReps=0, bPeriodic=0, bEwald=0))
    assert out2["nActive"] == nAct2
    assert np.all(out2["acc"][act_g == 0] == 0)
    pkd.close()
