"""ctypes binding of oracle/_ref/libgasref.so -- the reference's OWN compiled objects (see oracle/Makefile,
oracle/ref_api.c).  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgasref.so")
GPU_LIB_PATH = os.path.join(_HERE, "_ref", "libgasref_gpu.so")  # same objects, pkdGravAll = the product's shim
BIN_PATH = os.path.join(_HERE, "_ref", "gasoline_ref")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available() -> bool:
    return os.path.exists(LIB_PATH)


def gpu_host_available() -> bool:
    return os.path.exists(GPU_LIB_PATH)


_libs = {}


def lib(gpu_host: bool = False):
    """gpu_host=False: the pure reference.  gpu_host=True: the reference host with the product's pkdGravAll."""
    if gpu_host not in _libs:
        L = C.CDLL(GPU_LIB_PATH if gpu_host else LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_void_p, _dp]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_build_tree.restype = C.c_double
        L.ref_build_tree.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.ref_num_nodes.argtypes = [C.c_void_p]
        L.ref_root.argtypes = [C.c_void_p]
        L.ref_export_nodes.argtypes = [C.c_void_p] + [_dp] * 7 + [_ip] * 5
        L.ref_export_particles.argtypes = [C.c_void_p, _ip, _dp, _dp, _dp, _dp, _dp, _ip]
        L.ref_export_root.argtypes = [C.c_void_p, _dp]
        L.ref_set_active_tree.argtypes = [C.c_void_p, _ip]
        L.ref_ewald_table.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp, C.c_int]
        L.ref_gravity.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                  C.c_double, _dp, _dp, _dp, _dp, _ip, _dp]
        L.ref_bucket_lists.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_step_ops.restype = None
        L.ref_step_ops.argtypes = [C.c_int, _dp, _dp, _dp, C.c_void_p, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                   _dp, C.c_int, _dp, C.c_double, C.c_int]
        L.ref_gravity_sun.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, _ip]
        L.ref_gravity_comove.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, _dp]
        from oracle.oracle import RUNG_ARGTYPES
        L.ref_rung_ops.restype = None
        L.ref_rung_ops.argtypes = RUNG_ARGTYPES
        _libs[gpu_host] = L
    return _libs[gpu_host]


def ref_rung_ops(*args, **kw):
    """The reference's own pkdInitDt / pkdAccelStep / pkdGravStep / pkdDtToRung / pkdActiveRung (ref_api.c)."""
    from oracle.oracle import rung_ops
    return rung_ops(lib().ref_rung_ops, *args, **kw)


def ref_step_ops(*args, **kw):
    """The reference's own pkdKick / pkdDrift / pkdGravStep (ref_api.c: ref_step_ops); same driver as the oracle's."""
    from oracle.oracle import step_ops
    return step_ops(lib().ref_step_ops, *args, **kw)


class RefGravity:
    """One-rank reference run: tree build (msrBuildTree path) + gravity (pstGravity -> pkdGravAll)."""

    def __init__(self, p, active=None, gpu_host=False):
        self._L = L = lib(gpu_host)
        self.n = p.n
        per = np.array(p.period, dtype=np.float64)
        act = None
        if active is not None:
            self._act = np.ascontiguousarray(active, dtype=np.int32)
            act = self._act.ctypes.data_as(C.c_void_p)
        self.h = L.ref_create(p.n, *(np.ascontiguousarray(a, dtype=np.float64) for a in
                                     (p.x, p.y, p.z, p.m, p.h)), act, per)
        self.period = per

    def close(self):
        if self.h:
            self._L.ref_destroy(self.h)
            self.h = None

    def build_tree(self, nBucket=8, theta=0.7, iOrder=4):
        self.t_build = self._L.ref_build_tree(self.h, nBucket, theta, iOrder)
        return self.t_build

    def tree(self):
        """SoA copy of the reference's kdNodes + particles in tree order + ilcnRoot."""
        L = self._L
        nn = L.ref_num_nodes(self.h)
        t = dict(nNodes=nn, iRoot=L.ref_root(self.h), period=self.period.copy(),
                 bnd=np.zeros((nn, 6)), r=np.zeros((nn, 3)), fMass=np.zeros(nn), fSoft=np.zeros(nn),
                 fOpen2=np.zeros(nn), mom=np.zeros((nn, 31)), bmom=np.zeros((nn, 6)),
                 pLower=np.zeros(nn, np.int32), pUpper=np.zeros(nn, np.int32),
                 iLower=np.zeros(nn, np.int32), iUpper=np.zeros(nn, np.int32), iDim=np.zeros(nn, np.int32))
        L.ref_export_nodes(self.h, t["bnd"], t["r"], t["fMass"], t["fSoft"], t["fOpen2"], t["mom"], t["bmom"],
                           t["pLower"], t["pUpper"], t["iLower"], t["iUpper"], t["iDim"])
        n = self.n
        t.update(iOrder=np.zeros(n, np.int32), x=np.zeros(n), y=np.zeros(n), z=np.zeros(n), m=np.zeros(n),
                 h=np.zeros(n), active=np.zeros(n, np.int32))
        L.ref_export_particles(self.h, t["iOrder"], t["x"], t["y"], t["z"], t["m"], t["h"], t["active"])
        t["root"] = np.zeros(35)
        L.ref_export_root(self.h, t["root"])
        return t

    def set_active_tree(self, active):
        """ACTIVE flags in tree order, after build_tree (ref_set_active_tree)."""
        self._L.ref_set_active_tree(self.h, np.ascontiguousarray(active, dtype=np.int32))

    def ewald_table(self, fhCut=2.8, iOrder=4):
        buf = np.zeros((4096, 5))
        n = self._L.ref_ewald_table(self.h, fhCut, iOrder, buf, 4096)
        return buf[:n].copy()

    def gravity(self, nReps, bPeriodic, iOrder=4, bEwald=1, iEwOrder=4, dEwCut=2.6, dEwhCut=2.8):
        L = self._L
        n, nn = self.n, L.ref_num_nodes(self.h)
        acc = np.zeros((n, 3)); pot = np.zeros(n); dt = np.zeros(n); w = np.zeros(n)
        counts = np.zeros((nn, 3), np.int32); stats = np.zeros(8)
        L.ref_gravity(self.h, nReps, bPeriodic, iOrder, bEwald, iEwOrder, dEwCut, dEwhCut, acc, pot, dt, w,
                      counts, stats)
        return dict(acc=acc, pot=pot, dtGrav=dt, fWeight=w, counts=counts, nActive=int(stats[0]),
                    dPartSum=float(stats[1]), dCellSum=float(stats[2]), dSoftSum=float(stats[3]),
                    dFlop=float(stats[4]), seconds=float(stats[5]))

    def gravity_comove(self, dRhoFac, iOrder=4):
        """pstGravity with bComove=1 on open boundaries (pkd.c:2967-2991): (acc, pot) in tree order."""
        acc, pot = np.zeros((self.n, 3)), np.zeros(self.n)
        self._L.ref_gravity_comove(self.h, iOrder, float(dRhoFac), acc, pot)
        return acc, pot

    def gravity_sun(self, dSunSoft, iOrder=4):
        """pstGravity with bDoSun=1 (pkd.c:3003-3041): (aSun[3], (nPart, nCellSoft, nCellNewt) of the dummy bucket)."""
        a, c3 = np.zeros(3), np.zeros(3, np.int32)
        self._L.ref_gravity_sun(self.h, iOrder, float(dSunSoft), a, c3)
        return a, c3

    def bucket_lists(self, iBucket, nReps, iOrder=4, nmax=20000):
        n3 = np.zeros(3, np.int32)
        ilp = np.zeros((nmax, 5)); ilcs = np.zeros((nmax, 11)); ilcn = np.zeros((nmax, 35))
        self._L.ref_bucket_lists(self.h, iBucket, nReps, iOrder, n3, ilp.ctypes.data, nmax, ilcs.ctypes.data,
                               nmax, ilcn.ctypes.data, nmax)
        return ilp[:n3[0]].copy(), ilcs[:n3[1]].copy(), ilcn[:n3[2]].copy()
