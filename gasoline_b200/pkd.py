"""Host-side mirror of the reference's `pkd` interface for the gravity path, over the C ABI.

The reference's per-rank object is `PKD` (pkd.h:597-660): it owns pStore (particles), kdNodes (tree), kdTop and
ilcnRoot, and the path is driven as  pkdBuildBinary -> pkdGravAll(-> pkdBucketWalk/Interact/Ewald).  This class
keeps those names, argument meanings and result conventions; every force is computed by the CUDA library
(gasoline_b200/lib/libgasoline_b200.so).  There is no CPU fallback: if the library or a GPU is missing, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from .ics import FLOAT_MAXVAL

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgasoline_b200.so")
GG_NMOM, GG_NROOT = 31, 35
GG_FLAG_WALK_ONLY, GG_FLAG_NO_DOWNLOAD = 1, 2
OPEN_JOSH, OPEN_ABSPAR, OPEN_RELPAR, OPEN_ABSTOT, OPEN_RELTOT = 1, 2, 3, 4, 5  # opentype.h:5-9 = GG_OPEN_*


class GasolineB200Error(RuntimeError):
    pass


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class gg_tree(C.Structure):
    _fields_ = [("nNodes", C.c_int), ("iRoot", C.c_int), ("bnd", _dp), ("r", _dp), ("fMass", _dp), ("fSoft", _dp),
                ("fOpen2", _dp), ("mom", _dp), ("pLower", _ip), ("pUpper", _ip), ("iLower", _ip), ("iUpper", _ip)]


class gg_particles(C.Structure):
    _fields_ = [("n", C.c_int), ("x", _dp), ("y", _dp), ("z", _dp), ("fMass", _dp), ("fSoft", _dp), ("active", _ip)]


class gg_params(C.Structure):
    _fields_ = [("nReps", C.c_int), ("bPeriodic", C.c_int), ("iOrder", C.c_int), ("bEwald", C.c_int),
                ("iEwOrder", C.c_int), ("fEwCut", C.c_double), ("fEwhCut", C.c_double), ("bComove", C.c_int),
                ("dRhoFac", C.c_double), ("fPeriod", C.c_double * 3), ("accumulate", C.c_int), ("flags", C.c_int),
                ("bDoSun", C.c_int), ("dSunSoft", C.c_double)]


class gg_stats(C.Structure):
    _fields_ = [("nActive", C.c_int), ("dPartSum", C.c_double), ("dCellSum", C.c_double), ("dSoftSum", C.c_double),
                ("dFlop", C.c_double), ("dFlopEwald", C.c_double), ("msTree", C.c_double), ("msEwald", C.c_double),
                ("msTotal", C.c_double), ("nKernelLaunches", C.c_int), ("nMaxPart", C.c_int),
                ("nMaxCellSoft", C.c_int), ("nMaxCellNewt", C.c_int), ("msWalk", C.c_double), ("msEval", C.c_double),
                ("nListEntries", C.c_double), ("aSun", C.c_double * 3), ("nSunPart", C.c_int), ("nSunCellSoft", C.c_int),
                ("nSunCellNewt", C.c_int)]


class gg_exchange_stats(C.Structure):
    _fields_ = [("msExport", C.c_double), ("msTransfer", C.c_double), ("msIngest", C.c_double), ("msTotal", C.c_double),
                ("bytesSent", C.c_double), ("bytesReceived", C.c_double), ("bytesWholeDomain", C.c_double),
                ("nKernelLaunches", C.c_int)]


GG_UNIQUE_ID_BYTES = 128
_lib = None


def load_library(path: str | None = None):
    """dlopen the CUDA library. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("GASOLINE_B200_LIB") or LIB_PATH  # the env override is for kernel experiments
    if not os.path.exists(path):
        raise GasolineB200Error(f"{path} is missing: run `python -m gasoline_b200.build` (nvcc, sm_100a). "
                                "There is no CPU fallback for this path.")
    L = C.CDLL(path)
    L.gg_last_error.restype = C.c_char_p
    L.gg_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.gg_destroy.argtypes = [C.c_void_p]
    L.gg_set_local.argtypes = [C.c_void_p, C.c_int, C.POINTER(gg_tree), C.POINTER(gg_particles)]
    L.gg_set_remote.argtypes = [C.c_void_p, C.c_int, C.POINTER(gg_tree), C.POINTER(gg_particles), C.c_int]
    L.gg_clear_remote.argtypes = [C.c_void_p]
    L.gg_let_export.argtypes = [C.c_void_p, C.c_int, _dp, C.POINTER(gg_params), C.POINTER(C.c_void_p),
                                C.POINTER(C.c_size_t), _ip]
    L.gg_export_size.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), _ip]
    L.gg_export_local.argtypes = [C.c_void_p, C.c_void_p]
    L.gg_set_remote_packed.argtypes = [C.c_void_p, C.c_int, _ip, C.c_void_p]
    L.gg_set_top.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp]
    L.gg_set_root_moments.argtypes = [C.c_void_p, _dp]
    L.gg_announce.argtypes = [C.c_void_p, C.c_void_p]
    L.gg_local_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]
    L.gg_local_particles.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
    L.gg_local_nodes.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8
    L.gg_local_end.argtypes = [C.c_void_p]
    L.gg_gravity.argtypes = [C.c_void_p, C.POINTER(gg_params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.POINTER(gg_stats)]
    L.gg_bucket_counts.argtypes = [C.c_void_p, C.c_void_p]
    L.gg_bucket_walk.argtypes = [C.c_void_p, C.POINTER(gg_params), C.c_int, _ip]
    L.gg_ewald_table.argtypes = [C.c_void_p, C.POINTER(gg_params), C.c_void_p, C.c_int, _ip]
    L.gg_bucket_interact.argtypes = [C.c_void_p, C.POINTER(gg_params), C.c_int, C.c_int, _dp, _dp, _dp, _ip]
    L.gg_bucket_ewald.argtypes = [C.c_void_p, C.POINTER(gg_params), C.c_int, C.c_int, _dp, _dp, _ip]
    L.gg_device_results.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 4
    L.gg_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.gg_host_free.argtypes = [C.c_void_p]
    L.gg_tree_build.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _ip, _ip, C.c_int, C.c_double, C.c_int, C.c_int,
                                C.POINTER(C.c_void_p)]
    L.gg_tree_build_open.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _ip, _ip, C.c_int, C.c_int, C.c_double, C.c_int,
                                     C.c_int, C.POINTER(C.c_void_p)]
    L.gg_tree_bnumbers.argtypes = [C.c_void_p, _dp]
    L.gg_tree_view.argtypes = [C.c_void_p, C.POINTER(gg_tree), _dp]
    L.gg_tree_free.argtypes = [C.c_void_p]
    L.gg_cell_moments.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp]
    L.gg_tree_moments_m2m.argtypes = [C.POINTER(gg_tree), C.POINTER(gg_particles), _dp]
    L.gg_build_local.argtypes = [C.c_void_p, C.c_int, C.POINTER(gg_particles), C.c_int, C.c_double, _ip, _ip, _dp]
    L.gg_build_local_open.argtypes = [C.c_void_p, C.c_int, C.POINTER(gg_particles), C.c_int, C.c_int, C.c_double, _ip, _ip,
                                      _dp]
    L.gg_set_active.argtypes = [C.c_void_p, _ip]
    L.gg_orb_bisect.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp, _ip, _dp, _dp, C.c_int, _dp, _ip, _ip]
    L.gg_orb_bisect_all.argtypes = L.gg_orb_bisect.argtypes
    L.gg_orb_split_wrap.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp]
    L.gg_build_info.argtypes = [C.c_void_p, _ip, _ip, _dp]
    L.gg_domain_summary.argtypes = [C.c_void_p] + [_dp] * 7
    L.gg_domain_moments_about.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.gg_tree_fetch.argtypes = [C.c_void_p] + [_dp] * 6 + [_ip] * 4 + [_dp] * 5 + [_ip]
    L.gg_state_load.argtypes = [C.c_void_p, C.c_int] + [_dp] * 8 + [_ip, C.c_double]
    L.gg_state_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, _ip]
    L.gg_state_build_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, _ip]
    L.gg_state_kick.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    L.gg_state_drift.argtypes = [C.c_void_p, C.c_double, _dp, C.c_int, _dp]
    L.gg_state_gravstep.argtypes = [C.c_void_p, C.c_double, _dp]
    L.gg_state_fetch.argtypes = [C.c_void_p] + [_dp] * 6 + [_ip, _dp]
    L.gg_state_init_dt.argtypes = [C.c_void_p, C.c_double]
    L.gg_state_accelstep.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    L.gg_state_dt_to_rung.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, _ip, _ip, _ip]
    L.gg_state_active_rung.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip]
    L.gg_state_set_rungs.argtypes = [C.c_void_p, _ip]
    L.gg_state_fetch_rungs.argtypes = [C.c_void_p, _ip, _ip]
    L.gg_orb_load.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gg_orb_bounds.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _ip]
    L.gg_orb_weight.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp, _dp]
    L.gg_orb_split.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp]
    L.gg_orb_fetch.argtypes = [C.c_void_p, C.c_void_p]
    L.gg_comm_unique_id.argtypes = [C.c_void_p]
    L.gg_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.gg_group_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.gg_group_destroy.argtypes = [C.c_void_p]
    L.gg_group_destroy.restype = None
    L.gg_comm_init_local.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.gg_comm_free.argtypes = [C.c_void_p]
    L.gg_comm_info.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip]
    L.gg_comm_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.gg_exchange.argtypes = [C.c_void_p, C.POINTER(gg_params), _dp, C.POINTER(gg_exchange_stats)]
    L.gg_measure_fp32_peak.argtypes = [C.c_void_p, _dp, _dp]
    L.gg_flush_l2.argtypes = [C.c_void_p]
    L.gg_timer_start.argtypes = [C.c_void_p]
    L.gg_timer_stop.argtypes = [C.c_void_p, _dp]
    _lib = L
    return L


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().gg_last_error().decode(errors="replace")
        raise GasolineB200Error(f"{what} failed ({rc}): {msg}")


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory (gg_host_alloc).  The allocation lives for the life of the
    process (a handful of bench/host staging buffers), so the array can be passed around freely."""
    L = load_library()
    count = int(np.prod(shape))
    n = max(count * np.dtype(dtype).itemsize, 1)
    p = C.c_void_p()
    _check(L.gg_host_alloc(C.byref(p), n), "gg_host_alloc")
    buf = (C.c_char * n).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def comm_unique_id() -> bytes:
    """gg_comm_unique_id (ncclGetUniqueId): created on ONE rank, handed to every rank's commInitNccl by the host."""
    buf = C.create_string_buffer(GG_UNIQUE_ID_BYTES)
    _check(load_library().gg_comm_unique_id(buf), "gg_comm_unique_id")
    return buf.raw


class Group:
    """gg_group: the meeting point of ranks that live in one process (threads)."""

    def __init__(self, n: int):
        self._L = load_library()
        self.handle = C.c_void_p()
        self.n = int(n)
        _check(self._L.gg_group_create(C.byref(self.handle), self.n), "gg_group_create")

    def __del__(self):
        try:
            if self.handle:
                self._L.gg_group_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


@dataclass
class GravityParams:
    """The scalar arguments of pkdGravAll (pkd.h:797-801) with the reference's defaults (master.c:399-675)."""
    nReps: int = 0            # nReplicas; 1 when periodic (master.c:1763-1765)
    bPeriodic: int = 0
    iOrder: int = 4
    bEwald: int = 1
    iEwOrder: int = 4
    fEwCut: float = 2.6
    fEwhCut: float = 2.8
    bComove: int = 0
    dRhoFac: float = 0.0
    bDoSun: int = 0
    dSunSoft: float = 0.0


class Tree:
    """SoA copy of a k-d tree with the reference's KDN semantics (pkd.h:454-469), arrays owned by numpy."""
    FIELDS = ("bnd", "r", "fMass", "fSoft", "fOpen2", "mom", "pLower", "pUpper", "iLower", "iUpper")

    def __init__(self, nNodes, iRoot, **arrays):
        self.nNodes, self.iRoot = int(nNodes), int(iRoot)
        for k in self.FIELDS:
            dt = np.int32 if k[0] in "pi" else np.float64
            setattr(self, k, np.ascontiguousarray(arrays[k], dtype=dt))

    def view(self, with_mom: bool = True) -> gg_tree:
        """with_mom=False leaves gg_tree.mom NULL: the device forms the cells' moments itself (gg_moments.cu)."""
        return gg_tree(self.nNodes, self.iRoot, _d(self.bnd), _d(self.r), _d(self.fMass), _d(self.fSoft),
                       _d(self.fOpen2), _d(self.mom) if with_mom else None, _i(self.pLower), _i(self.pUpper),
                       _i(self.iLower), _i(self.iUpper))

    def as_dict(self):
        d = {k: getattr(self, k) for k in self.FIELDS}
        d.update(nNodes=self.nNodes, iRoot=self.iRoot)
        return d


class PKD:
    """One rank's particle store + tree on one B200 (mirrors struct pkdContext, pkd.h:597-660)."""

    def __init__(self, device: int = -1, idSelf: int = 0, fPeriod=(FLOAT_MAXVAL,) * 3, pinned: bool = False,
                 device_moments: bool = False):
        """pinned: keep pStore / kdNodes copies in page-locked host memory (what a host does with gg_host_alloc so
        that the per-step upload runs at full PCIe/C2C speed).  device_moments: do not transfer kdNodes[].mom (58 % of
        the upload); the device forms the multipole moments from the particles (gg_tree.mom = NULL)."""
        self._L = load_library()
        self.pinned = pinned
        self.device_moments = bool(device_moments)
        self._ctx = C.c_void_p()
        _check(self._L.gg_create(C.byref(self._ctx), device), "gg_create")
        self.idSelf = idSelf
        self.fPeriod = tuple(float(v) for v in fPeriod)
        self.nLocal = 0
        self.tree: Tree | None = None
        self.ilcnRoot = None
        self.stats = None
        self.iOrderMap = None  # tree position -> input index

    # -- lifetime -------------------------------------------------------------------------------------------
    def pkdFinish(self):
        if self._ctx:
            self._L.gg_destroy(self._ctx)
            self._ctx = C.c_void_p()

    close = pkdFinish

    def __del__(self):
        try:
            self.pkdFinish()
        except Exception:
            pass

    # -- particles + tree -----------------------------------------------------------------------------------
    def pkdLoadParticles(self, x, y, z, fMass, fSoft, active=None):
        """Fill pStore (what pkdReadTipsy pkd.c:297 does on the host). Arrays are copied; order = iOrder."""
        self.x, self.y, self.z, self.fMass, self.fSoft = (self._own(a, np.float64) for a in (x, y, z, fMass, fSoft))
        self.nLocal = int(self.x.shape[0])
        self.active = None if active is None else self._own(active, np.int32)
        self.iOrderMap = np.arange(self.nLocal, dtype=np.int32)
        self.tree = None
        self._resident = False

    def _own(self, a, dt):
        """Private copy of a host array, in pinned memory when the PKD was created with pinned=True."""
        a = np.asarray(a)
        if not self.pinned:
            return np.array(a, dtype=dt, copy=True)
        out = pinned_empty(a.shape, dt)
        out[...] = a
        return out

    def pkdBuildBinary(self, nBucket: int = 8, dCrit: float = 0.7, iOrder: int = 4, nThreads: int = 0,
                       iOpenType: int = OPEN_JOSH):
        """pkdBuildBinary (pkd.c:2627) + pkdCalcRoot (pkd.c:4395): spatial-bisection tree, opening radius by pkdCalcOpen
        (pkd.c:2228-2264; iOpenType as in opentype.h:5-9, default OPEN_JOSH with theta=dCrit); permutes pStore into tree
        order.  Host-side C++ (csrc/gg_tree_build.cpp)."""
        if self.nLocal == 0:
            raise GasolineB200Error("pkdBuildBinary: no particles")
        bt = C.c_void_p()
        order = np.zeros(self.nLocal, dtype=np.int32)
        act = _i(self.active) if self.active is not None else None
        _check(self._L.gg_tree_build_open(self.nLocal, _d(self.x), _d(self.y), _d(self.z), _d(self.fMass), _d(self.fSoft),
                                          act, _i(order), nBucket, int(iOpenType), dCrit, iOrder, nThreads, C.byref(bt)),
               "gg_tree_build_open")
        try:
            v = gg_tree()
            root = np.zeros(GG_NROOT)
            _check(self._L.gg_tree_view(bt, C.byref(v), _d(root)), "gg_tree_view")
            nn = v.nNodes
            cp = lambda p, shape, dt: self._own(np.ctypeslib.as_array(p, shape=shape), dt)
            self.tree = Tree(nn, v.iRoot, bnd=cp(v.bnd, (nn, 6), np.float64), r=cp(v.r, (nn, 3), np.float64),
                             fMass=cp(v.fMass, (nn,), np.float64), fSoft=cp(v.fSoft, (nn,), np.float64),
                             fOpen2=cp(v.fOpen2, (nn,), np.float64), mom=cp(v.mom, (nn, GG_NMOM), np.float64),
                             pLower=cp(v.pLower, (nn,), np.int32), pUpper=cp(v.pUpper, (nn,), np.int32),
                             iLower=cp(v.iLower, (nn,), np.int32), iUpper=cp(v.iUpper, (nn,), np.int32))
            self.ilcnRoot = root
        finally:
            self._L.gg_tree_free(bt)
        self.iOrderMap = self.iOrderMap[order]
        self._uploaded = False
        return self.tree

    def pkdBuildBinaryDevice(self, nBucket: int = 8, dCrit: float = 0.7, want_root: bool = False,
                             iOpenType: int = OPEN_JOSH):
        """pkdBuildBinary (pkd.c:2627) ON THE DEVICE (gg_build_local, csrc/gg_tree_gpu.cu): the particles go up in their
        current order, the tree -- bit-identical to pkdBuildBinary's -- is built and left loaded on the GPU, moments
        formed there; pkdGravAll follows without any tree transfer.  Results come back in tree order; iOrderMap maps
        tree position -> original index.  self.tree stays None (pkdFetchTree downloads it when a host wants it).
        iOpenType: OPEN_JOSH, or one of the criteria whose radius is Bmax (OPEN_RELPAR / ABSTOT / RELTOT, pkd.c:2261-2264);
        OPEN_ABSPAR is the host builder's (pkdBuildBinary)."""
        if self.nLocal == 0:
            raise GasolineB200Error("pkdBuildBinaryDevice: no particles")
        pv = gg_particles(self.nLocal, _d(self.x), _d(self.y), _d(self.z), _d(self.fMass), _d(self.fSoft),
                          _i(self.active) if self.active is not None else None)
        if getattr(self, "_devOrder", None) is None or self._devOrder.shape[0] != self.nLocal:
            self._devOrder = pinned_empty((self.nLocal,), np.int32) if self.pinned else np.zeros(self.nLocal, np.int32)
        nn = C.c_int()
        root = np.zeros(GG_NROOT) if want_root else None
        _check(self._L.gg_build_local_open(self._ctx, self.idSelf, C.byref(pv), int(nBucket), int(iOpenType), float(dCrit),
                                           _i(self._devOrder), C.byref(nn), _d(root) if want_root else None),
               "gg_build_local_open")
        self.nNodesDevice = int(nn.value)
        self.treeOrder = self._devOrder  # tree position -> index into the arrays given to pkdLoadParticles
        self.tree = None
        if want_root:
            self.ilcnRoot = root
        self._uploaded = True
        return self.nNodesDevice

    def pkdBuildInfo(self):
        """(nNodes, tree levels, device milliseconds) of the last pkdBuildBinaryDevice."""
        nn, nl, ms = C.c_int(), C.c_int(), C.c_double()
        _check(self._L.gg_build_info(self._ctx, C.byref(nn), C.byref(nl), C.byref(ms)), "gg_build_info")
        return nn.value, nl.value, ms.value

    def pkdSetActive(self, active):
        """New ACTIVE flags (tree order; None = all) for the loaded domain, without re-uploading tree or particles
        (gg_set_active) -- msrActiveRung between two force evaluations on one tree."""
        if not getattr(self, "_uploaded", False):
            raise GasolineB200Error("pkdSetActive: no domain loaded")
        a = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
        _check(self._L.gg_set_active(self._ctx, _i(a) if a is not None else None), "gg_set_active")
        self.active = a

    def pkdDomainSummary(self):
        """The root cell of the device-built local tree (gg_domain_summary): dict bnd, r, fMass, fSoft, fOpen2, mom,
        root (pkdCalcRoot's expansion of this rank) -- what pstColCells / pstCalcRoot collect from a rank."""
        bnd, r, sc, mom, root = np.zeros(6), np.zeros(3), np.zeros(3), np.zeros(GG_NMOM), np.zeros(GG_NROOT)
        _check(self._L.gg_domain_summary(self._ctx, _d(bnd), _d(r), _d(sc[0:1]), _d(sc[1:2]), _d(sc[2:3]), _d(mom),
                                         _d(root)), "gg_domain_summary")
        return dict(bnd=bnd, r=r, fMass=float(sc[0]), fSoft=float(sc[1]), fOpen2=float(sc[2]), mom=mom, root=root)

    def pkdDomainMomentsAbout(self, rcm):
        """pkdCalcCell over the whole local domain about rcm (gg_domain_moments_about): (mom[31], Bmax)."""
        c = np.ascontiguousarray(rcm, dtype=np.float64)
        mom, bmax = np.zeros(GG_NMOM), np.zeros(1)
        _check(self._L.gg_domain_moments_about(self._ctx, _d(c), _d(mom), _d(bmax)), "gg_domain_moments_about")
        return mom, float(bmax[0])

    def pkdFetchTree(self, with_mom: bool = True):
        """Download the device-built tree (gg_tree_fetch): returns (Tree, dict of the particles in tree order)."""
        nn, n = self.nNodesDevice, self.nLocal
        a = dict(bnd=np.zeros((nn, 6)), r=np.zeros((nn, 3)), fMass=np.zeros(nn), fSoft=np.zeros(nn), fOpen2=np.zeros(nn),
                 mom=np.zeros((nn, GG_NMOM)), pLower=np.zeros(nn, np.int32), pUpper=np.zeros(nn, np.int32),
                 iLower=np.zeros(nn, np.int32), iUpper=np.zeros(nn, np.int32))
        p = dict(x=np.zeros(n), y=np.zeros(n), z=np.zeros(n), fMass=np.zeros(n), fSoft=np.zeros(n),
                 active=np.zeros(n, np.int32) if self.active is not None else None)
        _check(self._L.gg_tree_fetch(self._ctx, _d(a["bnd"]), _d(a["r"]), _d(a["fMass"]), _d(a["fSoft"]), _d(a["fOpen2"]),
                                     _d(a["mom"]) if with_mom else None, _i(a["pLower"]), _i(a["pUpper"]),
                                     _i(a["iLower"]), _i(a["iUpper"]), _d(p["x"]), _d(p["y"]), _d(p["z"]),
                                     _d(p["fMass"]), _d(p["fSoft"]),
                                     _i(p["active"]) if p["active"] is not None else None), "gg_tree_fetch")
        return Tree(nn, 0, **a), p

    # -- device-resident particle store (gg_state_*): pStore lives in HBM, kick / drift / tree build / gravity on it --
    def pkdLoadResident(self, x, y, z, vx, vy, vz, fMass, fSoft, active=None, dt0: float = 1e30):
        """Upload pStore once (gg_state_load); afterwards pkdDrift / pkdKick / pkdBuildBinaryResident / pkdGravAll /
        pkdGravStep run on the device copy with no per-step particle traffic."""
        cols = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, vx, vy, vz, fMass, fSoft)]
        self.nLocal = int(cols[0].shape[0])
        act = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
        self.active = act
        _check(self._L.gg_state_load(self._ctx, self.nLocal, *[_d(a) for a in cols], _i(act) if act is not None else None,
                                     float(dt0)), "gg_state_load")
        self.tree = None
        self._uploaded = False
        self._resident = True

    def pkdBuildBinaryResident(self, nBucket: int = 8, dCrit: float = 0.7, iOpenType: int = OPEN_JOSH):
        """pkdBuildBinary on the resident store (gg_state_build): the store is permuted into tree order on the device."""
        nn = C.c_int()
        _check(self._L.gg_state_build_open(self._ctx, self.idSelf, int(nBucket), int(iOpenType), float(dCrit), C.byref(nn)),
               "gg_state_build_open")
        self.nNodesDevice = int(nn.value)
        self._uploaded = True
        return self.nNodesDevice

    # -- ORB domain decomposition: the rank's services of pstDomainDecomp (driver: domain.pst_domain_decomp)
    def pkdOrbLoad(self, x=None, y=None, z=None, fWeight=None):
        """gg_orb_load: this rank's particles enter the decomposition in ROOT.  x is None: the resident store's positions
        (pkdLoadResident).  fWeight None: 1 for every particle (what reading a file leaves, pkd.c:686)."""
        if x is None:
            n, ptr = self.nLocal, [None, None, None]
            self._orb_keep = []
        else:
            cols = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z)]
            n, ptr = int(cols[0].shape[0]), [a.ctypes.data for a in cols]
            self._orb_keep = cols
        w = None if fWeight is None else np.ascontiguousarray(fWeight, dtype=np.float64)
        _check(self._L.gg_orb_load(self._ctx, n, ptr[0], ptr[1], ptr[2], None if w is None else w.ctypes.data), "gg_orb_load")
        self._orb_n = n

    def pkdCalcBound(self, iCell):
        """pstCalcBound's leaf for the PST cells iCell: (bnd [k][6] = fMin, fMax; nIn [k]) of this rank's particles."""
        ic = np.ascontiguousarray(iCell, dtype=np.int32)
        bnd, nIn = np.zeros((len(ic), 6)), np.zeros(len(ic), np.int32)
        _check(self._L.gg_orb_bounds(self._ctx, len(ic), _i(ic), _d(bnd), _i(nIn)), "gg_orb_bounds")
        return bnd, nIn

    def pkdWeight(self, iCell, iDim, fSplit):
        """pstWeight's leaf (pkdWeight, pkd.c:945) for one trial split per PST cell: nLow, nHigh, fLow, fHigh."""
        ic, idim = np.ascontiguousarray(iCell, dtype=np.int32), np.ascontiguousarray(iDim, dtype=np.int32)
        fs = np.ascontiguousarray(fSplit, dtype=np.float64)
        nLow, nHigh = np.zeros(len(ic), np.int32), np.zeros(len(ic), np.int32)
        fLow, fHigh = np.zeros(len(ic)), np.zeros(len(ic))
        _check(self._L.gg_orb_weight(self._ctx, len(ic), _i(ic), _i(idim), _d(fs), _i(nLow), _i(nHigh), _d(fLow), _d(fHigh)),
               "gg_orb_weight")
        return nLow, nHigh, fLow, fHigh

    def pkdOrbBisect(self, iCell, iDim, fLow, fUp, live, nLower, nUpper, split_work: bool = True, collective: bool = False):
        """_pstRootSplit's root finder for all cells of a level with its state on the device (gg_orb_bisect): returns
        (fSplit, hasSplit, ittr).  collective=False: for a context that holds ALL particles of the decomposition.
        collective=True (gg_orb_bisect_all): every rank of this context's communicator calls it with the same arguments and
        holds part of the particles; the ranks' answers to a trial are all-gathered between the devices."""
        ic, idim = np.ascontiguousarray(iCell, dtype=np.int32), np.ascontiguousarray(iDim, dtype=np.int32)
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (fLow, fUp, nLower, nUpper)]
        lv = np.ascontiguousarray(live, dtype=np.int32)
        k = len(ic)
        fs, has, ittr = np.zeros(k), np.zeros(k, np.int32), np.zeros(k, np.int32)
        fn = self._L.gg_orb_bisect_all if collective else self._L.gg_orb_bisect
        _check(fn(self._ctx, k, _i(ic), _i(idim), _d(a[0]), _d(a[1]), _i(lv), _d(a[2]), _d(a[3]),
                  1 if split_work else 0, _d(fs), _i(has), _i(ittr)), "gg_orb_bisect_all" if collective else "gg_orb_bisect")
        return fs, has.astype(bool), ittr

    def pkdOrbSplit(self, iCell, iDim, fSplit):
        ic, idim = np.ascontiguousarray(iCell, dtype=np.int32), np.ascontiguousarray(iDim, dtype=np.int32)
        fs = np.ascontiguousarray(fSplit, dtype=np.float64)
        _check(self._L.gg_orb_split(self._ctx, len(ic), _i(ic), _i(idim), _d(fs)), "gg_orb_split")

    def pkdOrbSplitWrap(self, iCell, iDim, fSplit, fSplitInactive):
        """gg_orb_split_wrap: the outcome of pkdColRejects with a second boundary (pkd.c:1463-1485) -- the lower child takes
        the wrapped interval between fSplitInactive and fSplit (the store-overflow split, pst.c:1049-1270)."""
        ic, idim = np.ascontiguousarray(iCell, dtype=np.int32), np.ascontiguousarray(iDim, dtype=np.int32)
        fs, fi = np.ascontiguousarray(fSplit, dtype=np.float64), np.ascontiguousarray(fSplitInactive, dtype=np.float64)
        _check(self._L.gg_orb_split_wrap(self._ctx, len(ic), _i(ic), _i(idim), _d(fs), _d(fi)), "gg_orb_split_wrap")

    def pkdOrbCells(self) -> np.ndarray:
        """The PST cell of every particle, in pkdOrbLoad order."""
        out = np.zeros(max(self._orb_n, 1), np.int32)
        _check(self._L.gg_orb_fetch(self._ctx, out.ctypes.data), "gg_orb_fetch")
        return out[:self._orb_n]

    def pkdKick(self, dvFacOne: float, dvFacTwo: float, a=None):
        """pkdKick (pkd.c:3780) on the resident store; a: optional [n][3] accelerations in the store's order (default:
        the device results of the last pkdGravAll)."""
        ap = None
        if a is not None:
            a = np.ascontiguousarray(a, dtype=np.float64)
            ap = a.ctypes.data_as(C.c_void_p)
        _check(self._L.gg_state_kick(self._ctx, float(dvFacOne), float(dvFacTwo), ap), "gg_state_kick")

    def pkdDrift(self, dDelta: float, fCenter=(0.0, 0.0, 0.0), bPeriodic: int = 0):
        """pkdDrift (pkd.c:3686) on the resident store, periodic wrap with this PKD's fPeriod."""
        c = np.array(fCenter, dtype=np.float64)
        L = np.array(self.fPeriod, dtype=np.float64)
        _check(self._L.gg_state_drift(self._ctx, float(dDelta), _d(c), int(bPeriodic), _d(L)), "gg_state_drift")
        self._uploaded = False

    def pkdGravStep(self, dEta: float) -> float:
        """pkdGravStep (pkd.c:4609) on the resident store; returns the smallest time step of all particles."""
        dmin = np.zeros(1)
        _check(self._L.gg_state_gravstep(self._ctx, float(dEta), _d(dmin)), "gg_state_gravstep")
        return float(dmin[0])

    def pkdInitDt(self, dDelta: float):
        """pkdInitDt (pkd.c:4818) on the resident store."""
        _check(self._L.gg_state_init_dt(self._ctx, float(dDelta)), "gg_state_init_dt")

    def pkdAccelStep(self, dEta: float, dVelFac: float = 1.0, dAccFac: float = 1.0, bEpsAcc: int = 1, bSqrtPhi: int = 0):
        """pkdAccelStep (pkd.c:4625) on the resident store, with the last pkdGravAll's accelerations / potentials."""
        _check(self._L.gg_state_accelstep(self._ctx, float(dEta), float(dVelFac), float(dAccFac), int(bEpsAcc),
                                          int(bSqrtPhi)), "gg_state_accelstep")

    def pkdDtToRung(self, iRung: int, dDelta: float, iMaxRung: int, bAll: int = 1):
        """pkdDtToRung (pkd.c:4715) on the resident store: returns (iMaxRungOut, nMaxRung, iMaxRungIdeal)."""
        o = np.zeros(3, np.int32)
        _check(self._L.gg_state_dt_to_rung(self._ctx, int(iRung), float(dDelta), int(iMaxRung), int(bAll), _i(o[0:1]),
                                           _i(o[1:2]), _i(o[2:3])), "gg_state_dt_to_rung")
        return int(o[2]), int(o[0]), int(o[1])

    def pkdActiveRung(self, iRung: int, bGreater: int = 1) -> int:
        """pkdActiveRung (pkd.c:4569) on the resident store; the flags also become the sink set of the loaded domain."""
        n = np.zeros(1, np.int32)
        _check(self._L.gg_state_active_rung(self._ctx, int(iRung), int(bGreater), _i(n)), "gg_state_active_rung")
        return int(n[0])

    def pkdSetRungs(self, rung):
        _check(self._L.gg_state_set_rungs(self._ctx, _i(np.ascontiguousarray(rung, dtype=np.int32))), "gg_state_set_rungs")

    def pkdFetchRungs(self):
        """(iRung, ACTIVE) of the resident store in its current order."""
        r, a = np.zeros(self.nLocal, np.int32), np.zeros(self.nLocal, np.int32)
        _check(self._L.gg_state_fetch_rungs(self._ctx, _i(r), _i(a)), "gg_state_fetch_rungs")
        return r, a

    def pkdFetchResident(self):
        """Download the resident store (current order): dict with r [n][3], v [n][3], iOrder, dt."""
        n = self.nLocal
        c = [np.zeros(n) for _ in range(6)]
        ids, dt = np.zeros(n, np.int32), np.zeros(n)
        _check(self._L.gg_state_fetch(self._ctx, *[_d(a) for a in c], _i(ids), _d(dt)), "gg_state_fetch")
        return dict(r=np.stack(c[:3], axis=1), v=np.stack(c[3:], axis=1), iOrder=ids, dt=dt)

    def pkdSetTree(self, tree: Tree, x, y, z, fMass, fSoft, active=None, ilcnRoot=None, iOrderMap=None):
        """Adopt a tree built elsewhere (e.g. the host's own kdNodes); particles must already be in its order."""
        self.x, self.y, self.z, self.fMass, self.fSoft = (self._own(a, np.float64) for a in (x, y, z, fMass, fSoft))
        self.nLocal = int(self.x.shape[0])
        self.active = None if active is None else self._own(active, np.int32)
        self.tree = tree
        self.ilcnRoot = None if ilcnRoot is None else np.array(ilcnRoot, dtype=np.float64, copy=True)
        self.iOrderMap = np.arange(self.nLocal, dtype=np.int32) if iOrderMap is None else np.asarray(iOrderMap, np.int32)
        self._uploaded = False

    def upload_sliced(self, nSlices: int = 4, announce: "GravityParams | None" = None):
        """The same ingest in slices (gg_local_begin / gg_local_particles / gg_local_nodes / gg_local_end): what a host
        that has to flatten AoS records first does (the pkdGravAll shim).  Moments are formed on the device."""
        if self.tree is None:
            raise GasolineB200Error("upload_sliced: build or set a tree first")
        t, n, nn = self.tree, self.nLocal, self.tree.nNodes
        if self.ilcnRoot is not None:
            _check(self._L.gg_set_root_moments(self._ctx, _d(self.ilcnRoot)), "gg_set_root_moments")
        if announce is not None:
            prm = self._params(announce, 0, 0)
            _check(self._L.gg_announce(self._ctx, C.byref(prm)), "gg_announce")
        try:
            bnd = np.ascontiguousarray(t.bnd.reshape(-1, 6)[t.iRoot]) if t.bnd is not None and t.bnd.size else None
            _check(self._L.gg_local_begin(self._ctx, self.idSelf, nn, t.iRoot, n, _d(bnd) if bnd is not None else None,
                                          1 if self.active is not None else 0), "gg_local_begin")
            off = lambda a, lo, k=1: C.c_void_p(int(a.ctypes.data) + int(a.itemsize) * int(lo) * k)
            cuts = np.linspace(0, n, nSlices + 1).astype(int)
            for lo, hi in reversed(list(zip(cuts[:-1], cuts[1:]))):  # (any order)
                _check(self._L.gg_local_particles(self._ctx, int(lo), int(hi - lo), off(self.x, lo), off(self.y, lo),
                                                  off(self.z, lo), off(self.fMass, lo), off(self.fSoft, lo),
                                                  off(self.active, lo) if self.active is not None else None),
                       "gg_local_particles")
            cuts = np.linspace(0, nn, nSlices + 1).astype(int)
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                _check(self._L.gg_local_nodes(self._ctx, int(lo), int(hi - lo), off(t.r, lo, 3), off(t.fMass, lo),
                                              off(t.fSoft, lo), off(t.fOpen2, lo), off(t.pLower, lo), off(t.pUpper, lo),
                                              off(t.iLower, lo), off(t.iUpper, lo)), "gg_local_nodes")
            _check(self._L.gg_local_end(self._ctx), "gg_local_end")
        finally:
            if announce is not None:
                _check(self._L.gg_announce(self._ctx, None), "gg_announce")
        self._uploaded = True

    def upload(self, announce: "GravityParams | None" = None):
        """Ingest pStore + kdNodes into HBM (gg_set_local); also pkd->ilcnRoot when present.  announce = the parameters
        of the pkdGravAll that follows (gg_announce): with Ewald on, the correction then starts while the domain is
        still being copied (the root expansion goes up first for that)."""
        if self.tree is None:
            raise GasolineB200Error("upload: build or set a tree first")
        tv = self.tree.view(with_mom=not self.device_moments)
        pv = gg_particles(self.nLocal, _d(self.x), _d(self.y), _d(self.z), _d(self.fMass), _d(self.fSoft),
                          _i(self.active) if self.active is not None else None)
        if self.ilcnRoot is not None:
            _check(self._L.gg_set_root_moments(self._ctx, _d(self.ilcnRoot)), "gg_set_root_moments")
        if announce is not None:
            prm = self._params(announce, 0, 0)
            _check(self._L.gg_announce(self._ctx, C.byref(prm)), "gg_announce")
        try:
            _check(self._L.gg_set_local(self._ctx, self.idSelf, C.byref(tv), C.byref(pv)), "gg_set_local")
        finally:
            if announce is not None:
                _check(self._L.gg_announce(self._ctx, None), "gg_announce")
        self._uploaded = True

    def pkdSetRemote(self, id_: int, tree: Tree, x, y, z, fMass, fSoft):
        """A remote domain's tree + particles (what pkdRemoteWalk reads via mdlAquire, walk.c:181)."""
        cols = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, fMass, fSoft)]
        tv = tree.view()
        pv = gg_particles(int(cols[0].shape[0]), *[_d(a) for a in cols], None)
        _check(self._L.gg_set_remote(self._ctx, id_, C.byref(tv), C.byref(pv), 0), "gg_set_remote")

    def let_export(self, bnds, g: "GravityParams"):
        """Locally essential (pruned) copies of this domain for the remote domains whose root bounds are bnds[r]
        (fMin[3], fMax[3]): gg_let_export.  Returns (device pointer, offsets[nRemote+1], hdr[nRemote][3]); the buffer
        belongs to the context and is valid until the next call."""
        b = np.ascontiguousarray(bnds, dtype=np.float64).reshape(-1, 6)
        nR = b.shape[0]
        prm = self._params(g, 0, 0)
        dev = C.c_void_p()
        offs = (C.c_size_t * (nR + 1))()
        hdr = np.zeros((nR, 3), dtype=np.int32)
        _check(self._L.gg_let_export(self._ctx, nR, _d(b), C.byref(prm), C.byref(dev), offs, _i(hdr)), "gg_let_export")
        return int(dev.value), np.array(list(offs), dtype=np.int64), hdr

    def export_size(self):
        """(bytes, [nNodes, nPart, iRoot]) of this domain in device record layout (gg_export_size)."""
        nbytes, hdr = C.c_size_t(), np.zeros(3, dtype=np.int32)
        _check(self._L.gg_export_size(self._ctx, C.byref(nbytes), _i(hdr)), "gg_export_size")
        return int(nbytes.value), hdr

    def export_local(self, device_ptr: int):
        """Write the local domain's device records to a device buffer (the send side of an NCCL all-gather)."""
        _check(self._L.gg_export_local(self._ctx, C.c_void_p(device_ptr)), "gg_export_local")

    def pkdSetRemotePacked(self, id_: int, hdr, device_ptr: int):
        """A remote domain from device records at device_ptr (a slice of an NCCL receive buffer)."""
        h = np.ascontiguousarray(hdr, dtype=np.int32)
        _check(self._L.gg_set_remote_packed(self._ctx, int(id_), _i(h), C.c_void_p(device_ptr)), "gg_set_remote_packed")

    def pkdDistribCells(self, pLower, bUsed, r, fMass, fSoft, fOpen2, mom):
        """pkdDistribCells (pkd.c:4376): the gathered top tree kdTop[0..nCell), heap indexed from ROOT=1."""
        a = [np.ascontiguousarray(pLower, np.int32), np.ascontiguousarray(bUsed, np.int32)] + \
            [np.ascontiguousarray(v, np.float64) for v in (r, fMass, fSoft, fOpen2, mom)]
        _check(self._L.gg_set_top(self._ctx, int(a[0].shape[0]), _i(a[0]), _i(a[1]), *[_d(v) for v in a[2:]]),
               "gg_set_top")

    def pkdDistribRoot(self, ilcnRoot):
        """pkdDistribRoot (pkd.c:4472): the Ewald root expansion for every rank."""
        self.ilcnRoot = np.array(ilcnRoot, dtype=np.float64, copy=True)
        _check(self._L.gg_set_root_moments(self._ctx, _d(self.ilcnRoot)), "gg_set_root_moments")

    # -- multi-rank exchange below the ABI (csrc/gg_comm.cu) ---------------------------------------------------
    def commInitNccl(self, unique_id: bytes, rank: int, nRanks: int):
        """gg_comm_init: this rank's NCCL communicator (collective over all ranks; id from comm_unique_id() on one rank)."""
        buf = C.create_string_buffer(bytes(unique_id), GG_UNIQUE_ID_BYTES)
        _check(self._L.gg_comm_init(self._ctx, buf, int(rank), int(nRanks)), "gg_comm_init")
        self.commRank, self.commSize = int(rank), int(nRanks)

    def commInitLocal(self, group: "Group", rank: int):
        """gg_comm_init_local: ranks that are threads of this process (several may share one GPU)."""
        _check(self._L.gg_comm_init_local(self._ctx, group.handle, int(rank)), "gg_comm_init_local")
        self._group = group  # keep it alive
        self.commRank, self.commSize = int(rank), group.n

    def commInfo(self):
        o = np.zeros(4, np.int32)
        _check(self._L.gg_comm_info(self._ctx, _i(o[0:1]), _i(o[1:2]), _i(o[2:3]), _i(o[3:4])), "gg_comm_info")
        return dict(rank=int(o[0]), nRanks=int(o[1]), transport="group" if o[2] else "nccl", nccl_version=int(o[3]))

    def commAllgather(self, a: np.ndarray) -> np.ndarray:
        """gg_comm_allgather of a small host array: returns [nRanks] + a.shape."""
        a = np.ascontiguousarray(a)
        out = np.zeros((self.commSize,) + a.shape, dtype=a.dtype)
        _check(self._L.gg_comm_allgather(self._ctx, a.ctypes.data_as(C.c_void_p), a.nbytes, out.ctypes.data_as(C.c_void_p)),
               "gg_comm_allgather")
        return out

    def pkdExchange(self, g: "GravityParams", bndAll=None, want_stats: bool = True):
        """gg_exchange: the collective that replaces pkdRemoteWalk's pulls (walk.c:181-304) -- pruned locally-essential trees
        of every other rank become this rank's remote domains.  bndAll [nRanks][6]: the ranks' root bounds (None: gathered
        by the library).  Returns the phase timings / byte counts."""
        prm = self._params(g, 0, 0)
        b = None if bndAll is None else np.ascontiguousarray(bndAll, dtype=np.float64)
        st = gg_exchange_stats()
        _check(self._L.gg_exchange(self._ctx, C.byref(prm), _d(b) if b is not None else None,
                                   C.byref(st) if want_stats else None), "gg_exchange")
        return {k: getattr(st, k) for k, _ in gg_exchange_stats._fields_}

    # -- the hot path ---------------------------------------------------------------------------------------
    def _params(self, g: GravityParams, accumulate: int, flags: int) -> gg_params:
        return gg_params(g.nReps, g.bPeriodic, g.iOrder, g.bEwald, g.iEwOrder, g.fEwCut, g.fEwhCut, g.bComove,
                         g.dRhoFac, (C.c_double * 3)(*self.fPeriod), accumulate, flags, int(g.bDoSun), float(g.dSunSoft))

    def pkdGravAll(self, g: GravityParams, a=None, fPot=None, dtGrav=None, fWeight=None, walk_only=False,
                   download=True, accumulate=None):
        """pkdGravAll (pkd.c:2868).  With arrays given: a, fPot accumulate (+=), dtGrav is a running max, fWeight is
        overwritten for active particles -- the reference's in-place semantics on pStore.  Without: fresh arrays.
        Returns a dict with the arrays (tree order) and the scalars the reference returns through pointers
        (nActive, dPartSum, dCellSum, dSoftSum, dFlop) plus device timings."""
        if not getattr(self, "_uploaded", False) and not getattr(self, "_resident", False):
            self.upload()  # (a resident store that moved since its last build makes gg_gravity fail loudly instead)
        n = self.nLocal
        accumulate = (1 if a is not None else 0) if accumulate is None else int(bool(accumulate))
        flags = (GG_FLAG_WALK_ONLY if walk_only else 0) | (0 if download else GG_FLAG_NO_DOWNLOAD)
        if a is None and download and not walk_only:
            a = np.zeros((n, 3)); fPot = np.zeros(n); dtGrav = np.zeros(n); fWeight = np.zeros(n)
        prm = self._params(g, accumulate, flags)
        st = gg_stats()
        ptr = lambda v: v.ctypes.data_as(C.c_void_p) if v is not None else None
        _check(self._L.gg_gravity(self._ctx, C.byref(prm), ptr(a), ptr(fPot), ptr(dtGrav), ptr(fWeight), C.byref(st)),
               "gg_gravity")
        self.stats = {k: (np.array(st.aSun[:]) if k == "aSun" else getattr(st, k)) for k, _ in gg_stats._fields_}
        out = dict(self.stats)
        out.update(acc=a, pot=fPot, dtGrav=dtGrav, fWeight=fWeight)
        return out

    def pkdGravAllChunked(self, g: GravityParams, a, fPot, dtGrav, fWeight, nChunks: int, on_chunk=None):
        """gg_gravity_chunked: pkdGravAll in overwrite mode into mapped pinned arrays, the list evaluation in nChunks
        launches; on_chunk(first, count) is called as each particle range becomes final.  Returns the stats dict plus
        `chunks` = [(first, count), ...] in call order."""
        if not getattr(self, "_uploaded", False) and not getattr(self, "_resident", False):
            self.upload()
        prm = self._params(g, 0, 0)
        st = gg_stats()
        seen = []

        def cb(user, first, count):
            seen.append((int(first), int(count)))
            if on_chunk is not None:
                on_chunk(int(first), int(count))

        fn = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int)(cb)
        ptr = lambda v: v.ctypes.data_as(C.c_void_p)
        self._L.gg_gravity_chunked.argtypes = [C.c_void_p, C.POINTER(gg_params)] + [C.c_void_p] * 4 + \
                                             [C.POINTER(gg_stats), C.c_int, C.c_void_p, C.c_void_p]
        _check(self._L.gg_gravity_chunked(self._ctx, C.byref(prm), ptr(a), ptr(fPot), ptr(dtGrav), ptr(fWeight), C.byref(st),
                                          int(nChunks), C.cast(fn, C.c_void_p), None), "gg_gravity_chunked")
        out = {k: (np.array(st.aSun[:]) if k == "aSun" else getattr(st, k)) for k, _ in gg_stats._fields_}
        out.update(acc=a, pot=fPot, dtGrav=dtGrav, fWeight=fWeight, chunks=seen)
        return out

    def upload_bytes(self) -> int:
        """Bytes gg_set_local copies host->device for the current tree + particles."""
        t = self.tree
        b = sum(getattr(t, k).nbytes for k in Tree.FIELDS if k != "bnd" and not (k == "mom" and self.device_moments))
        b += sum(a.nbytes for a in (self.x, self.y, self.z, self.fMass, self.fSoft))
        if self.active is not None:
            b += self.active.nbytes
        return int(b)

    def tree_moments_m2m(self) -> np.ndarray:
        """The cells' reduced multipoles by the device's bottom-up algorithm, executed on the host
        (gg_tree_moments_m2m): [nNodes][GG_NMOM]."""
        out = np.zeros((self.tree.nNodes, GG_NMOM))
        tv = self.tree.view()
        pv = gg_particles(self.nLocal, _d(self.x), _d(self.y), _d(self.z), _d(self.fMass), _d(self.fSoft), None)
        _check(self._L.gg_tree_moments_m2m(C.byref(tv), C.byref(pv), _d(out)), "gg_tree_moments_m2m")
        return out

    def measure_fp32_peak(self):
        """(TFLOP/s, ms) of the dependent-FFMA microbenchmark on this GPU (gg_measure_fp32_peak)."""
        tf, ms = C.c_double(), C.c_double()
        _check(self._L.gg_measure_fp32_peak(self._ctx, C.byref(tf), C.byref(ms)), "gg_measure_fp32_peak")
        return tf.value, ms.value

    def flush_l2(self):
        _check(self._L.gg_flush_l2(self._ctx), "gg_flush_l2")

    def timer_start(self):
        """CUDA event on the library's stream (gg_timer_start)."""
        _check(self._L.gg_timer_start(self._ctx), "gg_timer_start")

    def timer_stop(self) -> float:
        """Milliseconds of device time since timer_start, host-induced gaps included (gg_timer_stop)."""
        ms = C.c_double()
        _check(self._L.gg_timer_stop(self._ctx, C.byref(ms)), "gg_timer_stop")
        return ms.value

    def pkdBucketCounts(self) -> np.ndarray:
        """(nPart, nCellSoft, nCellNewt) per tree node after pkdGravAll -- what pkdBucketWalk leaves in
        pkd->nPart/nCellSoft/nCellNewt (walk.c:175-177); -1 where no active sink bucket."""
        nn = self.tree.nNodes if self.tree is not None else self.nNodesDevice
        counts = np.zeros((nn, 3), dtype=np.int32)
        _check(self._L.gg_bucket_counts(self._ctx, counts.ctypes.data_as(C.c_void_p)), "gg_bucket_counts")
        return counts

    def pkdBucketWalk(self, iBucket: int, g: GravityParams):
        """pkdBucketWalk (walk.c:306) for one bucket: returns (nPart, nCellSoft, nCellNewt)."""
        if not getattr(self, "_uploaded", False):
            self.upload()
        prm = self._params(g, 0, 0)
        n3 = np.zeros(3, dtype=np.int32)
        _check(self._L.gg_bucket_walk(self._ctx, C.byref(prm), int(iBucket), _i(n3)), "gg_bucket_walk")
        return tuple(int(v) for v in n3)

    def pkdBucketInteract(self, iBucket: int, g: GravityParams, nMax: int = 64):
        """pkdBucketWalk + pkdBucketInteract (walk.h:32, grav.h:100) for one bucket: (acc [nP][3], pot, dtGrav, (nPart,
        nCellSoft, nCellNewt)) of the bucket's particles."""
        if not getattr(self, "_uploaded", False):
            self.upload()
        prm = self._params(g, 0, 0)
        a, p, d, n3 = np.zeros((nMax, 3)), np.zeros(nMax), np.zeros(nMax), np.zeros(3, np.int32)
        _check(self._L.gg_bucket_interact(self._ctx, C.byref(prm), int(iBucket), nMax, _d(a), _d(p), _d(d), _i(n3)),
               "gg_bucket_interact")
        return a, p, d, tuple(int(v) for v in n3)

    def pkdBucketEwald(self, iBucket: int, g: GravityParams, nMax: int = 64):
        """pkdBucketEwald (ewald.h:8) for one bucket: (acc [nP][3], pot, flops as the reference returns them)."""
        if not getattr(self, "_uploaded", False):
            self.upload()
        prm = self._params(g, 0, 0)
        a, p, nf = np.zeros((nMax, 3)), np.zeros(nMax), np.zeros(1, np.int32)
        _check(self._L.gg_bucket_ewald(self._ctx, C.byref(prm), int(iBucket), nMax, _d(a), _d(p), _i(nf)), "gg_bucket_ewald")
        return a, p, int(nf[0])

    def pkdEwaldInit(self, fhCut: float = 2.8, iOrder: int = 4) -> np.ndarray:
        """pkdEwaldInit (ewald.c:182): rows (hx,hy,hz,hCfac,hSfac) of the k-space table."""
        g = GravityParams(bPeriodic=1, iEwOrder=iOrder, fEwhCut=fhCut)
        prm = self._params(g, 0, 0)
        buf = np.zeros((4096, 5))
        n = C.c_int()
        _check(self._L.gg_ewald_table(self._ctx, C.byref(prm), buf.ctypes.data_as(C.c_void_p), 4096, C.byref(n)),
               "gg_ewald_table")
        return buf[: n.value].copy()
