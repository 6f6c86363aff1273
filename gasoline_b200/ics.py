"""Synthetic initial conditions and Tipsy I/O for the gravity hot path.

The workloads are the ones BASELINE.json / SURVEY.md section 8(d) freeze:

* periodic boxes (C1, C3, C4, C5): L=1, n^3 particles on a cell-centred grid in [-0.5, 0.5) displaced by a
  Gaussian random field (Zel'dovich displacements, P(k) ~ k^-2 cut at the grid Nyquist, rms displacement one
  grid spacing) or, for the small C1 case, i.i.d. jitter; positions are rounded to float32 exactly as a native
  Tipsy file stores them (tipsydefs.h:17-23, widened to double on read, pkd.c:686-695);
* C2: Plummer sphere a=1, M=1, truncated at r <= 20.

Units G=1.  Everything is numpy on the host; this is input generation, not the timed path.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

FLOAT_MAXVAL = 1.7976931348623157e308  # floattype.h:19 -- "not periodic" marker (walk.c:326)


@dataclass
class Particles:
    """Host particle set in input (iOrder) order. All arrays float64; positions hold float32 values."""

    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    m: np.ndarray
    h: np.ndarray
    period: tuple  # (Lx, Ly, Lz); FLOAT_MAXVAL on an axis = open
    name: str = ""

    @property
    def n(self) -> int:
        return int(self.x.shape[0])

    @property
    def periodic(self) -> bool:
        return self.period[0] < FLOAT_MAXVAL


def _f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def periodic_box(n: int, seed: int = 12345, mode: str = "zeldovich", rms_disp: float = 1.0,
                 jitter: float = 0.3) -> Particles:
    """n^3 particles in the unit box centred on the origin (master.c:1728 fCenter=0)."""
    if n >= 384 and mode == "zeldovich":
        return _periodic_box_large(n, seed, rms_disp)
    rng = np.random.default_rng(seed)
    g = (np.arange(n) + 0.5) / n - 0.5
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    if mode == "jitter":
        d = rng.normal(0.0, jitter / n, size=(3, n, n, n))
    else:
        k1 = np.fft.fftfreq(n, d=1.0 / n) * 2 * np.pi
        kx, ky, kz = np.meshgrid(k1, k1, np.fft.rfftfreq(n, d=1.0 / n) * 2 * np.pi, indexing="ij")
        k2 = kx * kx + ky * ky + kz * kz
        k2[0, 0, 0] = 1.0
        knyq = np.pi * n
        amp = np.where(k2 <= knyq * knyq, k2 ** -0.5, 0.0)  # sqrt(P), P ~ k^-2
        amp[0, 0, 0] = 0.0
        dk = amp * (rng.normal(size=k2.shape) + 1j * rng.normal(size=k2.shape))
        # displacement = grad(phi): psi_k = i k delta_k / k^2
        d = np.stack([np.fft.irfftn(1j * kk * dk / k2, s=(n, n, n), axes=(0, 1, 2)) for kk in (kx, ky, kz)])
        d *= rms_disp / n / np.sqrt(np.mean(np.sum(d * d, axis=0)) / 3.0 + 1e-300)
    pos = []
    for A, dd in zip((X, Y, Z), d):
        p = A + dd
        p = p - np.floor(p + 0.5)  # wrap into [-0.5, 0.5)
        p = _f32(p.ravel())
        p[p >= 0.5] -= 1.0
        pos.append(p)
    N = n ** 3
    return Particles(pos[0], pos[1], pos[2], _f32(np.full(N, 1.0 / N)), _f32(np.full(N, 1.0 / (20.0 * n))),
                     (1.0, 1.0, 1.0), f"periodic_{mode}_{n}^3_seed{seed}")


def _periodic_box_large(n: int, seed: int, rms_disp: float) -> Particles:
    """The same Zel'dovich realisation recipe for boxes of 384^3 and more (BASELINE.json configs[4]: 512^3): threaded
    FFTs (scipy.fft, all cores), one displacement component in memory at a time, broadcasting instead of meshgrids.
    (Not bit-identical to the small-box path -- another FFT library -- so the boxes the fixtures use keep theirs.)"""
    import os
    import scipy.fft as sfft
    rng = np.random.default_rng(seed)
    workers = os.cpu_count() or 1
    k1 = (np.fft.fftfreq(n, d=1.0 / n) * 2 * np.pi)
    k3 = (np.fft.rfftfreq(n, d=1.0 / n) * 2 * np.pi)
    kx, ky, kz = k1[:, None, None], k1[None, :, None], k3[None, None, :]
    k2 = kx * kx + ky * ky + kz * kz
    k2[0, 0, 0] = 1.0
    knyq = np.pi * n
    dk = rng.normal(size=k2.shape) + 1j * rng.normal(size=k2.shape)
    dk *= np.where(k2 <= knyq * knyq, k2 ** -1.5, 0.0)  # sqrt(P) / k^2, P ~ k^-2
    dk[0, 0, 0] = 0.0
    del k2
    disp = [sfft.irfftn(1j * kk * dk, s=(n, n, n), axes=(0, 1, 2), workers=workers) for kk in (kx, ky, kz)]
    del dk
    norm = rms_disp / n / np.sqrt(sum(float(np.mean(d * d)) for d in disp) / 3.0 + 1e-300)
    g = (np.arange(n) + 0.5) / n - 0.5
    pos = []
    for ax, d in enumerate(disp):
        d *= norm
        d += g.reshape([n if a == ax else 1 for a in range(3)])
        d -= np.floor(d + 0.5)  # wrap into [-0.5, 0.5)
        q = _f32(d.ravel())
        q[q >= 0.5] -= 1.0
        pos.append(q)
        disp[ax] = None
    N = n ** 3
    return Particles(pos[0], pos[1], pos[2], _f32(np.full(N, 1.0 / N)), _f32(np.full(N, 1.0 / (20.0 * n))),
                     (1.0, 1.0, 1.0), f"periodic_zeldovich_{n}^3_seed{seed}")


def plummer(N: int, seed: int = 12345, eps: float = 0.005, rmax: float = 20.0) -> Particles:
    """Plummer sphere a=1, M=1 by inverse CDF, truncated at rmax, isotropic; open boundaries."""
    rng = np.random.default_rng(seed)
    mmax = rmax ** 3 / (1.0 + rmax * rmax) ** 1.5
    u = rng.uniform(0.0, mmax, N)
    r = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    ct = rng.uniform(-1.0, 1.0, N)
    ph = rng.uniform(0.0, 2 * np.pi, N)
    st = np.sqrt(1.0 - ct * ct)
    return Particles(_f32(r * st * np.cos(ph)), _f32(r * st * np.sin(ph)), _f32(r * ct),
                     _f32(np.full(N, 1.0 / N)), _f32(np.full(N, eps)),
                     (FLOAT_MAXVAL,) * 3, f"plummer_{N}_seed{seed}")


def write_tipsy_native(path: str, p: Particles, time: float = 0.0) -> None:
    """Native (host-endian) Tipsy: 32-byte header, then dark records of 9 float32
    (mass,pos[3],vel[3],eps,phi) -- tipsydefs.h:17-23,38-45; header padding pkd.c:262-263."""
    N = p.n
    with open(path, "wb") as f:
        f.write(struct.pack("=diiiii", time, N, 3, 0, N, 0) + b"\0\0\0\0")
        rec = np.zeros((N, 9), dtype=np.float32)
        rec[:, 0] = p.m
        rec[:, 1] = p.x
        rec[:, 2] = p.y
        rec[:, 3] = p.z
        rec[:, 7] = p.h
        rec.tofile(f)


def write_tipsy_standard(path: str, p: Particles, time: float = 0.0) -> None:
    """Standard Tipsy (bStandard = 1): the same field sequence as the native file written through XDR, i.e. big-endian
    -- header double time, ints nbodies ndim nsph ndark nstar and one pad int (xdrHeader, master.c:3563-3575: 32 bytes),
    then per dark particle 9 big-endian float32 (pkdReadTipsy's xdr_float sequence, pkd.c:456-560; sizes pkdSeek
    pkd.c:253-283)."""
    N = p.n
    with open(path, "wb") as f:
        f.write(struct.pack(">diiiiii", time, N, 3, 0, N, 0, 0))
        rec = np.zeros((N, 9), dtype=">f4")
        rec[:, 0] = p.m
        rec[:, 1] = p.x
        rec[:, 2] = p.y
        rec[:, 3] = p.z
        rec[:, 7] = p.h
        rec.tofile(f)


def read_tipsy_standard(path: str) -> Particles:
    """Dark particles of a standard (XDR, big-endian) Tipsy file; gas records (12 floats) are skipped."""
    with open(path, "rb") as f:
        time, n, ndim, nsph, ndark, nstar, _pad = struct.unpack(">diiiiii", f.read(32))
        f.seek(48 * nsph, 1)
        rec = np.fromfile(f, dtype=">f4", count=9 * ndark).reshape(ndark, 9).astype(np.float64)
    return Particles(rec[:, 1].copy(), rec[:, 2].copy(), rec[:, 3].copy(), rec[:, 0].copy(), rec[:, 7].copy(),
                     (FLOAT_MAXVAL,) * 3, os.path.basename(path))


def read_array_binary(path: str, ndim: int) -> np.ndarray:
    """Reference binary array output (iBinaryOutput=2): int N, then N doubles per dimension
    (master.c:5527, outtype.c:1001-1012). Returns (N, ndim)."""
    with open(path, "rb") as f:
        n = struct.unpack("=i", f.read(4))[0]
        a = np.fromfile(f, dtype=np.float64, count=n * ndim)
    return a.reshape(ndim, n).T.copy()


def read_tipsy_native(path: str) -> Particles:
    """Read the dark particles of a native Tipsy file (what pkdReadTipsy pkd.c:297 does: float32 on disk, widened to
    double, pkd.c:686-695).  Gas / star records are skipped; open boundaries are assumed (the period is a run parameter)."""
    with open(path, "rb") as f:
        time, n, ndim, nsph, ndark, nstar = struct.unpack("=diiiii", f.read(28))
        f.read(4)  # header padding to 32 bytes (pkd.c:262-263)
        f.seek(48 * nsph, 1)  # gas records: 12 float32 (tipsydefs.h:9-15)
        rec = np.fromfile(f, dtype=np.float32, count=9 * ndark).reshape(ndark, 9).astype(np.float64)
    return Particles(rec[:, 1].copy(), rec[:, 2].copy(), rec[:, 3].copy(), rec[:, 0].copy(), rec[:, 7].copy(),
                     (FLOAT_MAXVAL,) * 3, os.path.basename(path))


def write_array_binary(path: str, a) -> None:
    """The reference's binary array output, iBinaryOutput=2 (master.c:5527, outtype.c:1001-1012): int N, then all first
    components, all second, ... as native doubles.  a: (N,) for a scalar (.pot, .dt) or (N, 3) for a vector (.accg)."""
    a = np.asarray(a, dtype=np.float64)
    a = a.reshape(a.shape[0], -1)
    with open(path, "wb") as f:
        f.write(struct.pack("=i", a.shape[0]))
        np.ascontiguousarray(a.T).tofile(f)


def write_array_ascii(path: str, a, interleaved: bool = True) -> None:
    """The reference's ASCII array output, iBinaryOutput=0: "N\n", then one value per line in %.14g.  A vector written by
    one node (msrOneNodeOutVector, master.c:5420: iDim = -3) comes particle by particle -- x, y, z of particle 0, then of
    particle 1, ... (outtype.c:949-957); interleaved=False gives the per-component order of the parallel writer
    (outtype.c:958-963: all first components, then all second, ...).  Scalars (.pot, .dt) are the same either way."""
    a = np.asarray(a, dtype=np.float64)
    a = a.reshape(a.shape[0], -1)
    v = a.ravel() if interleaved else a.T.ravel()
    with open(path, "w") as f:
        f.write(f"{a.shape[0]}\n")
        f.write("".join("%.14g\n" % x for x in v))


def read_array_ascii(path: str, ndim: int, interleaved: bool = True) -> np.ndarray:
    """Inverse of write_array_ascii: returns (N, ndim)."""
    with open(path) as f:
        n = int(f.readline())
        v = np.array([float(x) for x in f.read().split()], dtype=np.float64)
    return v.reshape(n, ndim).copy() if interleaved else v.reshape(ndim, n).T.copy()
