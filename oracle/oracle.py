"""ctypes binding of oracle/liboracle.so (gravity_oracle.c, our CPU restatement of the reference algorithm).
TEST INFRASTRUCTURE ONLY -- see the header of gravity_oracle.c for who may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gravity_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_void_p, _dp]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_build_tree.restype = C.c_double
        L.orc_build_tree.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.orc_build_tree_open.restype = C.c_double
        L.orc_build_tree_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_export_bnumbers.argtypes = [C.c_void_p, _dp]
        L.orc_num_nodes.argtypes = [C.c_void_p]
        L.orc_root.argtypes = [C.c_void_p]
        L.orc_export_nodes.argtypes = [C.c_void_p] + [_dp] * 7 + [_ip] * 5
        L.orc_export_particles.argtypes = [C.c_void_p, _ip, _dp, _dp, _dp, _dp, _dp, _ip]
        L.orc_export_root.argtypes = [C.c_void_p, _dp]
        L.orc_set_active_tree.argtypes = [C.c_void_p, _ip]
        L.orc_import_tree.argtypes = [C.c_void_p, C.c_int, C.c_int] + [_dp] * 6 + [_ip] * 4 + [_dp]
        L.orc_ewald_table.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_int]
        L.orc_gravity.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                  _dp, _dp, _dp, _dp, _ip, _dp, C.c_int, C.c_int]
        L.orc_bucket_lists.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_step_ops.restype = C.c_int
        L.oracle_step_ops.argtypes = [C.c_int, _dp, _dp, _dp, C.c_void_p, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                      _dp, C.c_int, _dp, C.c_double, C.c_int]
        L.oracle_rung_ops.restype = None
        L.oracle_rung_ops.argtypes = RUNG_ARGTYPES
        _lib = L
    return _lib


RUNG_ARGTYPES = [C.c_int, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _ip, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip]
INITDT, ACCELSTEP, GRAVSTEP_R, DTTORUNG, ACTIVERUNG = 1, 2, 4, 8, 16


def rung_ops(fn, v, a, fPot, fSoft, dtGrav, active, dt, rung, dDelta=0.01, dEta=0.2, dVelFac=1.0, dAccFac=1.0, bEpsAcc=1,
             bSqrtPhi=0, iRung=0, iMaxRung=16, bAll=1, iRungActive=0, bGreater=1, what=INITDT | ACCELSTEP | DTTORUNG):
    """Common driver of oracle_rung_ops / ref_rung_ops (same C signature): returns updated copies (active, dt, rung)
    and out = [iMaxRungOut, nMaxRung, iMaxRungIdeal, nActive]."""
    c = lambda x, t: np.array(x, dtype=t, copy=True)
    active, dt, rung = c(active, np.int32), c(dt, np.float64), c(rung, np.int32)
    out = np.zeros(4, np.int32)
    fn(dt.shape[0], np.ascontiguousarray(v, np.float64).reshape(-1, 3), np.ascontiguousarray(a, np.float64).reshape(-1, 3),
       np.ascontiguousarray(fPot, np.float64), np.ascontiguousarray(fSoft, np.float64),
       np.ascontiguousarray(dtGrav, np.float64), active, dt, rung, float(dDelta), float(dEta), float(dVelFac),
       float(dAccFac), int(bEpsAcc), int(bSqrtPhi), int(iRung), int(iMaxRung), int(bAll), int(iRungActive),
       int(bGreater), int(what), out)
    return active, dt, rung, out


def oracle_rung_ops(*args, **kw):
    """pkdInitDt / pkdAccelStep / pkdGravStep / pkdDtToRung / pkdActiveRung restated (gravity_oracle.c)."""
    return rung_ops(lib().oracle_rung_ops, *args, **kw)


KICK, DRIFT, GRAVSTEP = 1, 2, 4


def step_ops(fn, r, v, a, active, dtGrav, dt, dvFacOne=1.0, dvFacTwo=0.0, dDelta=0.0, fCenter=(0, 0, 0), bPeriodic=0,
             fPeriod=(1, 1, 1), dEta=0.2, what=KICK | DRIFT):
    """Common driver of oracle_step_ops / ref_step_ops (same C signature): returns updated copies (r, v, dt) and the
    function's return value."""
    r = np.array(r, dtype=np.float64, copy=True).reshape(-1, 3)
    v = np.array(v, dtype=np.float64, copy=True).reshape(-1, 3)
    dt = np.array(dt, dtype=np.float64, copy=True)
    act = None
    if active is not None:
        act_arr = np.ascontiguousarray(active, dtype=np.int32)
        act = act_arr.ctypes.data_as(C.c_void_p)
    rc = fn(r.shape[0], r, v, np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3), act,
            np.ascontiguousarray(dtGrav, dtype=np.float64), dt, float(dvFacOne), float(dvFacTwo), float(dDelta),
            np.array(fCenter, dtype=np.float64), int(bPeriodic), np.array(fPeriod, dtype=np.float64), float(dEta), int(what))
    return r, v, dt, rc


def oracle_step_ops(*args, **kw):
    """pkdKick / pkdDrift / pkdGravStep restated (gravity_oracle.c: oracle_step_ops)."""
    return step_ops(lib().oracle_step_ops, *args, **kw)


class OracleGravity:
    """Same surface as oracle.reflib.RefGravity, backed by our restatement."""

    def __init__(self, p, active=None, tree=None):
        """p: Particles (input order) -- or, with `tree` (dict as returned by .tree()), the particles are taken from
        the tree (already in tree order) and the tree is imported instead of built."""
        L = lib()
        self.period = np.array(p.period if tree is None else tree["period"], dtype=np.float64)
        act = None
        if tree is not None:
            self.n = len(tree["x"])
            self._act = np.ascontiguousarray(tree["active"], dtype=np.int32)
            act = self._act.ctypes.data_as(C.c_void_p)
            cols = [np.ascontiguousarray(tree[k], dtype=np.float64) for k in ("x", "y", "z", "m", "h")]
        else:
            self.n = p.n
            if active is not None:
                self._act = np.ascontiguousarray(active, dtype=np.int32)
                act = self._act.ctypes.data_as(C.c_void_p)
            cols = [np.ascontiguousarray(a, dtype=np.float64) for a in (p.x, p.y, p.z, p.m, p.h)]
        self.h = L.orc_create(self.n, *cols, act, self.period)
        if tree is not None:
            c = lambda k, dt: np.ascontiguousarray(tree[k], dtype=dt)
            L.orc_import_tree(self.h, int(tree["nNodes"]), int(tree["iRoot"]), c("bnd", np.float64),
                              c("r", np.float64), c("fMass", np.float64), c("fSoft", np.float64),
                              c("fOpen2", np.float64), c("mom", np.float64), c("pLower", np.int32),
                              c("pUpper", np.int32), c("iLower", np.int32), c("iUpper", np.int32),
                              c("root", np.float64))

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def build_tree(self, nBucket=8, theta=0.7, iOrder=4, iOpenType=1):
        """iOpenType as in opentype.h:5-9 (1 = OPEN_JOSH with dCrit = theta; 2 = OPEN_ABSPAR with dCrit = the error bound)."""
        self.t_build = lib().orc_build_tree_open(self.h, nBucket, iOpenType, theta, iOrder)
        return self.t_build

    def bnumbers(self):
        """B2..B6 of every cell of the tree built last (pkdCalcCell, pkd.c:2083-2087): [nNodes][5]."""
        out = np.zeros((lib().orc_num_nodes(self.h), 5))
        lib().orc_export_bnumbers(self.h, out)
        return out

    def tree(self):
        L = lib()
        nn = L.orc_num_nodes(self.h)
        t = dict(nNodes=nn, iRoot=L.orc_root(self.h), period=self.period.copy(),
                 bnd=np.zeros((nn, 6)), r=np.zeros((nn, 3)), fMass=np.zeros(nn), fSoft=np.zeros(nn),
                 fOpen2=np.zeros(nn), mom=np.zeros((nn, 31)), bmax=np.zeros(nn),
                 pLower=np.zeros(nn, np.int32), pUpper=np.zeros(nn, np.int32),
                 iLower=np.zeros(nn, np.int32), iUpper=np.zeros(nn, np.int32), iDim=np.zeros(nn, np.int32))
        L.orc_export_nodes(self.h, t["bnd"], t["r"], t["fMass"], t["fSoft"], t["fOpen2"], t["mom"], t["bmax"],
                           t["pLower"], t["pUpper"], t["iLower"], t["iUpper"], t["iDim"])
        n = self.n
        t.update(iOrder=np.zeros(n, np.int32), x=np.zeros(n), y=np.zeros(n), z=np.zeros(n), m=np.zeros(n),
                 h=np.zeros(n), active=np.zeros(n, np.int32))
        L.orc_export_particles(self.h, t["iOrder"], t["x"], t["y"], t["z"], t["m"], t["h"], t["active"])
        t["root"] = np.zeros(35)
        L.orc_export_root(self.h, t["root"])
        return t

    def set_active_tree(self, active):
        lib().orc_set_active_tree(self.h, np.ascontiguousarray(active, dtype=np.int32))

    def ewald_table(self, fhCut=2.8, iOrder=4):
        buf = np.zeros((4096, 5))
        n = lib().orc_ewald_table(self.h, fhCut, iOrder, buf.ctypes.data, 4096)
        return buf[:n].copy()

    def gravity(self, nReps, bPeriodic, iOrder=4, bEwald=1, iEwOrder=4, dEwCut=2.6, dEwhCut=2.8, walk_only=False,
                threads=0):
        L = lib()
        n, nn = self.n, L.orc_num_nodes(self.h)
        acc = np.zeros((n, 3)); pot = np.zeros(n); dt = np.zeros(n); w = np.zeros(n)
        counts = np.zeros((nn, 3), np.int32); stats = np.zeros(8)
        L.orc_gravity(self.h, nReps, bPeriodic, iOrder, bEwald, iEwOrder, dEwCut, dEwhCut, acc, pot, dt, w, counts,
                      stats, int(walk_only), threads)
        return dict(acc=acc, pot=pot, dtGrav=dt, fWeight=w, counts=counts, nActive=int(stats[0]),
                    dPartSum=float(stats[1]), dCellSum=float(stats[2]), dSoftSum=float(stats[3]),
                    dFlop=float(stats[4]), seconds=float(stats[5]))

    def bucket_lists(self, iBucket, nReps, iOrder=4, nmax=20000):
        n3 = np.zeros(3, np.int32)
        ilp = np.zeros((nmax, 5)); ilcs = np.zeros((nmax, 11)); ilcn = np.zeros((nmax, 35))
        lib().orc_bucket_lists(self.h, iBucket, nReps, iOrder, n3, ilp.ctypes.data, nmax, ilcs.ctypes.data, nmax,
                               ilcn.ctypes.data, nmax)
        return ilp[:n3[0]].copy(), ilcs[:n3[1]].copy(), ilcn[:n3[2]].copy()
