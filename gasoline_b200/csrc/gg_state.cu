// gg_state.cu -- the particle store resident on the device across force evaluations and the two steps either side of
// the force path: pkdKick (pkd.c:3780; -DNBODY branch pkd.c:3956-3962), pkdDrift
// (pkd.c:3686-3777) and pkdGravStep (pkd.c:4609-4623).  Element-wise FP64, HBM-bound (kick: 24 B acceleration + 48 B
// velocity traffic per particle; drift: 72 B), written with __dmul_rn/__dadd_rn in the reference's operation order so
// the results are bit-identical to the reference's (which is compiled without FMA).
#include "gg_internal.h"

namespace {

// v = v*dvFacOne + a*dvFacTwo on ACTIVE particles; v is SoA [3][n], a is the force kernels' AoS [n][3]
__global__ void __launch_bounds__(256) k_kick(int n, double *v, const double *a, const int *active, double f1, double f2) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (active && !active[i]) return;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const size_t k = (size_t)j * n + i;
        v[k] = __dadd_rn(__dmul_rn(v[k], f1), __dmul_rn(a[3 * (size_t)i + j], f2));
    }
}

// r += dDelta*v for ALL particles, then the reference's two independent wrap tests per axis (pkd.c:3731-3757)
__global__ void __launch_bounds__(256) k_drift(int n, double *x, double *y, double *z, const double *v, double dDelta,
                                               double cx, double cy, double cz, int bPeriodic, double Lx, double Ly,
                                               double Lz, int *nOutside) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double *r[3] = {x + i, y + i, z + i};
    const double c[3] = {cx, cy, cz}, L[3] = {Lx, Ly, Lz};
    int bad = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double q = __dadd_rn(*r[j], __dmul_rn(dDelta, v[(size_t)j * n + i]));
        if (bPeriodic) {
            const double hi = __dadd_rn(c[j], __dmul_rn(0.5, L[j])), lo = __dsub_rn(c[j], __dmul_rn(0.5, L[j]));
            if (q >= hi) q = __dsub_rn(q, L[j]);
            if (q < lo) q = __dadd_rn(q, L[j]);
            bad |= !(q >= lo && q < hi);
        }
        *r[j] = q;
    }
    if (bad) atomicAdd(nOutside, 1);
}

// dt = min(dt, dEta/sqrt(dtGrav)) on ACTIVE particles; the smallest dt of all particles -> dtMinBits (ordered integer)
__global__ void __launch_bounds__(256) k_gravstep(int n, double *dt, const double *dtGrav, const int *active, double dEta,
                                                  unsigned long long *dtMinBits, int *nBad) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    double d = __longlong_as_double(0x7ff0000000000000ll);
    if (i < n) {
        d = dt[i];
        if (!active || active[i]) {
            if (!(dtGrav[i] > 0.0)) atomicAdd(nBad, 1); // the reference asserts dtGrav > 0 (pkd.c:4616)
            const double g = __ddiv_rn(dEta, __dsqrt_rn(dtGrav[i]));
            if (g < d) { d = g; dt[i] = d; }
        }
    }
    // positive doubles order like their bit patterns
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    const unsigned h = (unsigned)(b >> 32), mh = __reduce_min_sync(0xffffffffu, h);
    const unsigned ml = __reduce_min_sync(0xffffffffu, h == mh ? (unsigned)b : 0xffffffffu);
    if ((threadIdx.x & 31) == 0) atomicMin(dtMinBits, ((unsigned long long)mh << 32) | ml);
}

// the state's per-particle payload follows the tree build's permutation: out[i] = in[iorder[i]]
// (id and rung travel interleaved: idr[2*i] = persistent id, idr[2*i+1] = rung)
__global__ void __launch_bounds__(256) k_permute(int n, const int *iorder, const double *vIn, double *vOut, const int *idIn,
                                                 int *idOut, const double *dtIn, double *dtOut) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int s = iorder[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) vOut[(size_t)j * n + i] = vIn[(size_t)j * n + s];
    reinterpret_cast<int2 *>(idOut)[i] = reinterpret_cast<const int2 *>(idIn)[s];
    dtOut[i] = dtIn[s];
}

__global__ void __launch_bounds__(256) k_state_init(int n, int *id, double *dt, double dt0) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    reinterpret_cast<int2 *>(id)[i] = make_int2(i, 0);
    dt[i] = dt0;
}

// pkdInitDt (pkd.c:4818): ACTIVE particles dt = dDelta
__global__ void __launch_bounds__(256) k_init_dt(int n, double *dt, const int *active, double dDelta) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n && (!active || active[i])) dt[i] = dDelta;
}

// pkdAccelStep (pkd.c:4625-4672, the -DNBODY build): ACTIVE particles, acc = sqrt(a.a) dAccFac;
// bEpsAcc: dT = dEta sqrt(fSoft / acc); bSqrtPhi: dT = min(dT, dEta 3.5 sqrt(dAccFac |fPot|) / acc); dt = min(dt, dT)
__global__ void __launch_bounds__(256) k_accelstep(int n, double *dt, const double *a, const double *pot, const double *fSoft,
                                                   const int *active, double dEta, double dAccFac, int bEpsAcc,
                                                   int bSqrtPhi) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n || (active && !active[i])) return;
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) acc = __dadd_rn(acc, __dmul_rn(a[3 * (size_t)i + j], a[3 * (size_t)i + j]));
    acc = __dmul_rn(__dsqrt_rn(acc), dAccFac);
    double dT = 1.7976931348623157e308;
    if (bEpsAcc && acc > 0) dT = __dmul_rn(dEta, __dsqrt_rn(__ddiv_rn(fSoft[i], acc)));
    if (bSqrtPhi && acc > 0) {
        const double t = __ddiv_rn(__dmul_rn(__dmul_rn(dEta, 3.5), __dsqrt_rn(__dmul_rn(dAccFac, fabs(pot[i])))), acc);
        if (t < dT) dT = t;
    }
    if (dT < dt[i]) dt[i] = dT;
}

// pkdDtToRung (pkd.c:4715-4810) with pkdOneParticleDtToRung (pkd.c:4689-4712).  hist[r] counts the particles left on
// rung r (ALL particles: the reference's iMaxRungOut / nMaxRung scan them all), ideal = max(rung + 1) before clamping.
__global__ void __launch_bounds__(256) k_dt_to_rung(int n, int *idr, const double *dt, int iRung, double dDelta, int iMaxRung,
                                                    int bAll, int *hist, int *ideal, int *nBad) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int r = idr[2 * (size_t)i + 1];
    if (r >= iRung) {
        if (bAll) {
            const double d = dt[i];
            // the reference asserts dt > 0 (pkd.c:4749) and dDelta/dt < 2.1e9 (pkd.c:4694: integer overflow)
            if (!(d > 0.0) || !(__ddiv_rn(dDelta, d) < 2.1e9)) { atomicAdd(nBad, 1); return; }
            int iSteps = (int)floor(__ddiv_rn(dDelta, d)), t = iRung;
            if (fmod(dDelta, d) == 0.0) iSteps--;
            if (iSteps < 0) iSteps = 0;
            t += 32 - __clz(iSteps); // one rung per bit of iSteps
            atomicMax(ideal, t + 1);
            if (t >= iMaxRung) t = iMaxRung - 1;
            r = t;
        } else r = dDelta <= dt[i] ? iRung : iRung + 1;
        idr[2 * (size_t)i + 1] = r;
    }
    atomicAdd(&hist[min(max(r, 0), 127)], 1);
}

// pkdActiveRung (pkd.c:4569-4590)
__global__ void __launch_bounds__(256) k_active_rung(int n, const int *idr, int *active, int iRung, int bGreater, int *count) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    int a = 0;
    if (i < n) {
        const int r = idr[2 * (size_t)i + 1];
        a = r == iRung || (bGreater && r > iRung);
        active[i] = a;
    }
    const int c = __popc(__ballot_sync(0xffffffffu, a));
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// max |r_p - rcm| over the particles (Bmax of pkdCalcCell, pkd.c:2018-2135, for one domain about a given centre):
// distances in the reference's expression, the maximum as an ordered-integer atomicMax
__global__ void __launch_bounds__(256) k_bmax_about(int n, const PartS *parts, double cx, double cy, double cz,
                                                    unsigned long long *out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    unsigned long long b = 0ull;
    if (i < n) {
        const double dx = __dsub_rn(parts[i].x, cx), dy = __dsub_rn(parts[i].y, cy), dz = __dsub_rn(parts[i].z, cz);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        b = (unsigned long long)__double_as_longlong(__dsqrt_rn(d2));
    }
    const unsigned h = (unsigned)(b >> 32), mh = __reduce_max_sync(0xffffffffu, h);
    const unsigned ml = __reduce_max_sync(0xffffffffu, h == mh ? (unsigned)b : 0u);
    if ((threadIdx.x & 31) == 0) atomicMax(out, ((unsigned long long)mh << 32) | ml);
}

} // namespace

cudaError_t gg_launch_bmax_about(int n, const PartS *parts, const double c[3], unsigned long long *out, cudaStream_t st) {
    if (n > 0) k_bmax_about<<<(n + 255) / 256, 256, 0, st>>>(n, parts, c[0], c[1], c[2], out);
    return cudaGetLastError();
}

cudaError_t gg_launch_kick(int n, double *v, const double *a, const int *active, double f1, double f2, cudaStream_t st) {
    if (n > 0) k_kick<<<(n + 255) / 256, 256, 0, st>>>(n, v, a, active, f1, f2);
    return cudaGetLastError();
}
cudaError_t gg_launch_drift(int n, double *x, double *y, double *z, const double *v, double dDelta, const double c[3],
                            int bPeriodic, const double L[3], int *nOutside, cudaStream_t st) {
    if (n > 0)
        k_drift<<<(n + 255) / 256, 256, 0, st>>>(n, x, y, z, v, dDelta, c[0], c[1], c[2], bPeriodic, L[0], L[1], L[2],
                                                 nOutside);
    return cudaGetLastError();
}
cudaError_t gg_launch_gravstep(int n, double *dt, const double *dtGrav, const int *active, double dEta,
                               unsigned long long *dtMinBits, int *nBad, cudaStream_t st) {
    if (n > 0) k_gravstep<<<(n + 255) / 256, 256, 0, st>>>(n, dt, dtGrav, active, dEta, dtMinBits, nBad);
    return cudaGetLastError();
}
cudaError_t gg_launch_permute(int n, const int *iorder, const double *vIn, double *vOut, const int *idIn, int *idOut,
                              const double *dtIn, double *dtOut, cudaStream_t st) {
    if (n > 0) k_permute<<<(n + 255) / 256, 256, 0, st>>>(n, iorder, vIn, vOut, idIn, idOut, dtIn, dtOut);
    return cudaGetLastError();
}
cudaError_t gg_launch_state_init(int n, int *id, double *dt, double dt0, cudaStream_t st) {
    if (n > 0) k_state_init<<<(n + 255) / 256, 256, 0, st>>>(n, id, dt, dt0);
    return cudaGetLastError();
}

cudaError_t gg_launch_init_dt(int n, double *dt, const int *active, double dDelta, cudaStream_t st) {
    if (n > 0) k_init_dt<<<(n + 255) / 256, 256, 0, st>>>(n, dt, active, dDelta);
    return cudaGetLastError();
}
cudaError_t gg_launch_accelstep(int n, double *dt, const double *a, const double *pot, const double *fSoft, const int *active,
                                double dEta, double dAccFac, int bEpsAcc, int bSqrtPhi, cudaStream_t st) {
    if (n > 0) k_accelstep<<<(n + 255) / 256, 256, 0, st>>>(n, dt, a, pot, fSoft, active, dEta, dAccFac, bEpsAcc, bSqrtPhi);
    return cudaGetLastError();
}
cudaError_t gg_launch_dt_to_rung(int n, int *idr, const double *dt, int iRung, double dDelta, int iMaxRung, int bAll, int *hist,
                                 int *ideal, int *nBad, cudaStream_t st) {
    if (n > 0) k_dt_to_rung<<<(n + 255) / 256, 256, 0, st>>>(n, idr, dt, iRung, dDelta, iMaxRung, bAll, hist, ideal, nBad);
    return cudaGetLastError();
}
cudaError_t gg_launch_active_rung(int n, const int *idr, int *active, int iRung, int bGreater, int *count, cudaStream_t st) {
    if (n > 0) k_active_rung<<<(n + 255) / 256, 256, 0, st>>>(n, idr, active, iRung, bGreater, count);
    return cudaGetLastError();
}
