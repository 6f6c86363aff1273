#!/usr/bin/env python
"""bench.py -- tree-gravity force evaluation (Gasoline's pkdGravAll path) on B200: interactions/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload periodic:256:0.5]

A "step" is ONE force evaluation = one pkdGravAll (pkd.c:2868) per rank over every sink bucket of the workload: tree
walks, list evaluation and the Ewald correction; with N > 1 ranks also what every rank's pkdGravAll pays per call in the
reference for its remote walks (pkdRemoteWalk, walk.c:181): here the upload of the top tree, the export of the pruned
locally-essential trees, their exchange over NCCL and their ingestion (gg_set_top + gg_exchange).  The metric is the
reference's own interaction count (pkd.c:2945-2949: particle-list + intra-bucket pairs + Newtonian-cell + softened-cell
entries summed over active sinks; Ewald terms are not interactions) divided by time.

Workload (every N): BASELINE.json configs[3] -- periodic box, 256^3 = 16.8 M particles, theta = 0.5, nReplicas = 1,
Ewald on -- the configuration north_star's targets are quoted on; it fits one GPU, so N > 1 is STRONG scaling: the same
box split into N ORB domains (the reference's pstDomainDecomp run on the devices), one per GPU.  At N = 1 the line also
carries configs[1] (isolated Plummer sphere, 1 M particles, theta = 0.7) under "configs1".

  value     device time per step (CUDA events on the stream the library launches on, recorded around the whole step:
            [gg_set_top + gg_exchange +] gg_gravity, host-induced gaps included) with this rank's tree and particles
            already resident in HBM; L2 is flushed between timed iterations; max over ranks.
  e2e       the same through the public host API with HOST buffers: every step re-ingests the host's tree + particles
            from pinned memory (gg_set_local: the reference frees and rebuilds kdNodes before every gravity call,
            pkd.c:2636-2642; the cells' moments are formed on the device instead of transferred), exchanges, evaluates
            and delivers a, fPot, dtGrav, fWeight into pinned host arrays; wall clock between barriers, max over ranks.
  e2e_pkdGravAll   the reference's own C host (pstBuildTree -> pstGravity -> pkdGravAll, compiled where it lies) with
            the product's pkdGravAll / pkdBuildBinary linked in: N = 1 in-process (oracle/_ref/libgasref_gpu.so), N > 1
            the reference BINARY on N pthread-MDL ranks, rank r on GPU r (oracle/_ref/gasoline_ref_gpu).
  parity    out of the timed region, at every N: a sample of complete sink buckets (of rank 0) re-evaluated by the CPU
            oracle on the same trees (N > 1: the ranks' trees hung under the top tree, oracle/multidomain.py, pinned to
            multi-rank runs of the reference): list counts bit-exact, accelerations / potentials within 1e-5 / 1e-4.
  roofline  the dominant kernel, k_eval (list evaluation), against the FP32 FMA pipe (FP32 CUDA-core + SFU work, not
            HBM- or tensor-bound -- DESIGN.md 5): achieved = the reference's own flop score of the lists it evaluated
            (grav.c:246-247) / the kernel's duration (CUDA events around the launch); peak = dependent-FFMA
            microbenchmark measured in this run on this GPU.  The HBM view is reported beside it.
  cpu_baseline / --impl reference
            the reference's own compiled pkdGravAll on the host cores (oracle/cpu_baseline.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "particle interactions/s (tree gravity, pkdGravAll)"
UNIT = "interactions/s"
DEFAULT_SPEC = "periodic:256:0.5"   # BASELINE.json configs[3]
CONFIGS1_SPEC = "plummer:1000000:0.7"  # BASELINE.json configs[1]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="plummer:N:theta | periodic:n:theta (default: configs[3])")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the legs beside the contract's (configs[1], the C-host leg, the from-particles legs)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--parity-buckets", type=int, default=48)
    ap.add_argument("--host-orb", action="store_true",
                    help="N > 1: domains from the host-side median split instead of the reference's ORB run on the devices")
    return ap.parse_args()


def workload_spec(a) -> str:
    return a.workload or DEFAULT_SPEC


def config_of(spec: str, n_gpus: int) -> dict:
    kind, n, theta = spec.split(":")
    if kind == "plummer":
        name = f"isolated Plummer sphere, {int(n)} particles, theta={theta}, no Ewald, nBucket=8, iOrder=4 (hexadecapole)"
        total = int(n)
    else:
        name = f"periodic box {n}^3 particles, theta={theta}, nReplicas=1, Ewald on, nBucket=8, iOrder=4"
        total = int(n) ** 3
    return {"workload": name, "spec": spec, "particles": total, "particles_per_gpu": total // n_gpus,
            "parallelism": "1 GPU" if n_gpus == 1 else f"{n_gpus} ORB domains, one per GPU; top tree + pruned LET "
                                                       "exchange over NCCL inside every timed step",
            "l2": "flushed between timed iterations (384 MB memset)"}


def make_ic(spec: str):
    from gasoline_b200 import ics
    from gasoline_b200.pkd import GravityParams
    kind, n_s, theta_s = spec.split(":")
    if kind == "plummer":
        return ics.plummer(int(n_s)), GravityParams(nReps=0, bPeriodic=0, bEwald=0), float(theta_s)
    return ics.periodic_box(int(n_s)), GravityParams(nReps=1, bPeriodic=1, bEwald=1), float(theta_s)


def make_ic_shared(spec: str, rank: int, world: int, barrier):
    """N > 1: the initial conditions are generated ONCE (rank 0) and shared with the other ranks of the node through
    /dev/shm as float32 columns (what a Tipsy file holds); every rank maps them and only ever touches the particles it
    needs (its chunk for the decomposition, then its domain).  Masses and softenings are uniform: zero-stride views."""
    from gasoline_b200 import ics
    from gasoline_b200.pkd import GravityParams
    kind, n_s, theta_s = spec.split(":")
    d = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(),
                     f"gg_ic_{os.environ.get('MASTER_PORT', '0')}_{spec.replace(':', '_')}")
    if rank == 0:
        p, g, theta = make_ic(spec)
        os.makedirs(d, exist_ok=True)
        for k in ("x", "y", "z"):
            np.save(os.path.join(d, k + ".npy"), getattr(p, k).astype(np.float32))
        np.save(os.path.join(d, "meta.npy"), np.array([p.m[0], p.h[0], p.period[0], p.period[1], p.period[2]]))
        del p
    barrier()
    meta = np.load(os.path.join(d, "meta.npy"))
    cols = [np.load(os.path.join(d, k + ".npy"), mmap_mode="r") for k in ("x", "y", "z")]
    n = cols[0].shape[0]
    p = ics.Particles(cols[0], cols[1], cols[2], np.broadcast_to(np.float64(meta[0]), (n,)),
                      np.broadcast_to(np.float64(meta[1]), (n,)), tuple(float(v) for v in meta[2:5]), f"shared:{spec}")
    g = GravityParams(nReps=0, bPeriodic=0, bEwald=0) if kind == "plummer" else GravityParams(nReps=1, bPeriodic=1, bEwald=1)
    return p, g, float(theta_s), d


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms: int = 50):
        """index: one GPU, or a comma-separated list (ONE poller for all ranks' GPUs: a poller per rank makes eight
        processes take the driver's lock twenty times a second, which showed up as milliseconds of skew between the
        ranks of a step that contains a collective)."""
        self.index, self.rows, self.proc, self.period = index, [], None, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period), "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            time.sleep(0.3)  # the first sample takes nvidia-smi a moment
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(a):
    """The reference's own CPU pkdGravAll on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import cpu_baseline
    spec = workload_spec(a)
    per_step = max(3.0, min(a.cpu_seconds, 100.0 / max(a.steps + a.warmup, 1)))
    t0 = time.time()
    r = cpu_baseline.run(spec, per_step, 0, "auto", repeats=max(1, a.steps + a.warmup))
    out = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config_of(spec, a.gpus),
           "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                            "sample": r["sample"],
                            "how": "the stock pkdGravAll on one shared tree: P forked workers with disjoint sink sets, best "
                                   "of the repeats -- no remote walks, no domain imbalance, so it flatters the reference"},
           "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def algorithmic_bytes(n_particles: int, n_nodes: int, n_list_entries: float) -> float:
    """Compulsory HBM traffic of one k_eval launch: every source particle record (32 B) and every tree
    node (64 B walk record + 128 B moment record) read once, the per-bucket interaction lists streamed once (4 B per
    entry), 40 B of results written per particle."""
    return 32.0 * n_particles + 192.0 * n_nodes + 4.0 * n_list_entries + 40.0 * n_particles


def ncu_traffic(spec: str):
    """DRAM bytes of one k_eval launch from the committed ncu capture of this workload (None if absent)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "k_eval_summary.json")))
        if d.get("workload") == spec:
            return d["dram_bytes_per_launch"]
        return d.get("by_workload", {}).get(spec, {}).get("dram_bytes_per_launch")
    except (OSError, KeyError, ValueError):
        return None


def bind_to_gpu_cpus(index: int):
    """Pin this rank to the CPU cores NVML reports as local to its GPU (what an MPI launcher's binding does for the
    reference).  Returns the number of cores in the mask, or None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def roofline_of(g, peak_tf, eval_ms, flop_tree, n, nodes, entries, spec, peaks):
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ach_tf = flop_tree / (eval_ms * 1e-3) * 1e-12 if eval_ms > 0 else 0.0
    alg = algorithmic_bytes(n, nodes, entries)
    return {"bound": "fp32", "kernel": f"k_eval<{g.iOrder},{'true' if g.bPeriodic else 'false'}>", "achieved": ach_tf,
            "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf if peak_tf else None,
            "traffic": ncu_traffic(spec),
            "peak_source": "dependent-FFMA microbenchmark measured in this run (nominal 74.4 TFLOP/s at 1965 MHz); "
                           "MEASURED_PEAKS.json holds HBM GB/s and bf16 TF/s only",
            "flops": "the reference's own score of the evaluated lists (grav.c:246-247: 38/particle, 82/soft cell, "
                     "312/hexadecapole cell), per launch = per GPU",
            "hbm": {"achieved": alg / (eval_ms * 1e-3) * 1e-9 if eval_ms > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg / (eval_ms * 1e-3) * 1e-9 / hbm_peak if eval_ms > 0 else 0.0, "algorithmic_bytes": alg,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback (B200_PROFILING.md)"},
            "ms_per_launch": eval_ms}


def parity_check(p, g, theta, idx_all, rank0_out, rank0_counts, n_buckets, seed=2026):
    """Sampled sink buckets of rank 0 against the CPU oracle on the same trees (N > 1: the combined tree of all ranks).
    rank0_out / rank0_counts: the GPU's results for rank 0's domain (tree order) and its per-node list counts."""
    from gasoline_b200 import domain
    from oracle import multidomain, oracle
    t0 = time.time()
    world = len(idx_all)
    doms = [domain.Domain(r, world, p.x[ix], p.y[ix], p.z[ix], p.m[ix], p.h[ix], p.period, theta) for r, ix in enumerate(idx_all)]
    if world > 1:
        domain.run_in_process(doms, exchange_trees=False)
        T, nodeBase, partBase = multidomain.from_domains(doms)
    else:
        h = doms[0].host
        T = h.tree.as_dict()
        T.update(x=h.x, y=h.y, z=h.z, m=h.fMass, h=h.fSoft, active=np.zeros(h.nLocal, np.int32), root=h.ilcnRoot,
                 period=np.array(p.period), iOrder=h.iOrderMap)
        nodeBase, partBase = {0: 0}, {0: 0}
    t = doms[0].host.tree
    bk = np.where(t.iLower == -1)[0]
    pick = np.random.default_rng(seed).choice(bk, size=min(n_buckets, len(bk)), replace=False)
    act0 = np.zeros(doms[0].host.nLocal, bool)
    for b in pick:
        act0[t.pLower[b]:t.pUpper[b] + 1] = True
    T["active"] = np.asarray(T["active"]).copy()
    T["active"][partBase[0] + np.nonzero(act0)[0]] = 1
    o = oracle.OracleGravity(None, tree=T)
    ref = o.gravity(g.nReps, g.bPeriodic, g.iOrder, g.bEwald, g.iEwOrder, g.fEwCut, g.fEwhCut)
    o.close()
    counts_ok = bool(np.array_equal(rank0_counts[pick], ref["counts"][nodeBase[0] + pick]))
    sl = partBase[0] + np.nonzero(act0)[0]
    ra, rp = ref["acc"][sl], ref["pot"][sl]
    ga, gp = rank0_out["acc"][act0], rank0_out["pot"][act0]
    d = np.linalg.norm(ga - ra, axis=1) / np.linalg.norm(ra, axis=1)
    floor = np.sqrt(np.mean(rp ** 2))
    dp = np.abs(gp - rp) / np.maximum(np.abs(rp), floor)
    dps = np.abs(gp - rp) / np.abs(rp)
    res = {"checker": "CPU oracle (oracle/gravity_oracle.c" + (", ranks' trees under the top tree: oracle/multidomain.py)" if world > 1 else ")"),
           "buckets": int(len(pick)), "sinks": int(act0.sum()), "of_rank": 0, "counts_bit_exact": counts_ok,
           "acc_rel_rms": float(np.sqrt(np.mean(d * d))), "acc_rel_max": float(d.max()),
           "pot_rel_rms": float(np.sqrt(np.mean(dp * dp))), "pot_rel_max": float(dp.max()),
           "pot_rel_max_strict": float(dps.max()),
           "pot_metric": "|dphi| / max(|phi_i|, rms phi); strict = |dphi| / |phi_i| per particle",
           "fweight_bit_exact": bool(np.array_equal(rank0_out["fWeight"][act0], ref["fWeight"][sl])),
           "tolerance": {"rms": 1e-5, "max": 1e-4}, "seconds": time.time() - t0}
    res["ok"] = bool(counts_ok and res["acc_rel_rms"] <= 1e-5 and res["acc_rel_max"] <= 1e-4 and
                     res["pot_rel_rms"] <= 1e-5 and res["pot_rel_max"] <= 1e-4)
    return res


def c_host_leg(spec, world, steps, warmup, timeout=900):
    """The reference's C host with the product's pkdGravAll + pkdBuildBinary: see tools/c_host_leg.py."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "c_host_leg.py"), "--workload", spec, "--ranks",
                            str(world), "--steps", str(steps), "--warmup", str(min(warmup, 2))], cwd=ROOT,
                           capture_output=True, text=True, timeout=timeout)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # the GPU line must still be printed
        return {"value": None, "unit": UNIT, "error": f"{type(e).__name__}: {e}"[:300]}


def measure_single(spec, a, local, full=True):
    """One GPU, one workload: resident timing, e2e, optional extra legs.  Returns a dict of measurements."""
    import torch
    from gasoline_b200.pkd import PKD, pinned_empty
    p, g, theta = make_ic(spec)
    t0 = time.time()
    pkd = PKD(device=local, fPeriod=p.period, pinned=True)
    pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    pkd.pkdBuildBinary(8, theta, 4)
    t_tree = time.time() - t0
    n = pkd.nLocal
    # gg_tree.mom = NULL: the cells' multipole moments (58 % of the tree bytes) are not transferred; the device forms
    # them from the particles while the walk runs (gg_moments.cu; forces identical, tests/test_gpu_device_moments.py)
    pkd.device_moments = True
    pkd.upload()
    peak_tf, _ = pkd.measure_fp32_peak()
    for _ in range(a.warmup):
        pkd.pkdGravAll(g, download=False)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    acc = dict(ms=0.0, tree=0.0, ewald=0.0, eval=0.0, walk=0.0, launches=0)
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        pkd.flush_l2()
        pkd.timer_start()
        st = pkd.pkdGravAll(g, download=False)
        acc["ms"] += pkd.timer_stop()
        acc["tree"] += st["msTree"]; acc["ewald"] += st["msEwald"]; acc["eval"] += st["msEval"]; acc["walk"] += st["msWalk"]
        acc["launches"] += st["nKernelLaunches"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    inter = st["dPartSum"] + st["dCellSum"] + st["dSoftSum"]
    # ---- end to end through the host API: host buffers in, host buffers out, every step
    outs = (pinned_empty((n, 3)), pinned_empty(n), pinned_empty(n), pinned_empty(n))
    # (upload(announce=g): the host says what evaluation follows, like pkdGravAll's one call does -- with Ewald on, the
    #  correction then runs beside the copies, see gg_announce)
    for _ in range(min(a.warmup, 2)):
        pkd.upload(announce=g)
        pkd.pkdGravAll(g, *outs, accumulate=False)
    torch.cuda.synchronize()
    e0 = time.perf_counter()
    for _ in range(a.steps):
        pkd.upload(announce=g)
        last = pkd.pkdGravAll(g, *outs, accumulate=False)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - e0) / a.steps
    m = dict(p=p, g=g, theta=theta, pkd=pkd, n=n, t_tree=t_tree, peak_tf=peak_tf, acc=acc, wall=wall, clocks=clocks, st=st,
             inter=inter, e2e_s=e2e_s, h2d=pkd.upload_bytes(), d2h=6 * 8 * n, outs=outs, last=last)
    if not full:
        return m
    # ---- the same step with the tree built on the device (gg_build_local): particles in input order
    #      in pinned host memory -> device tree build (bit-identical tree) -> pkdGravAll -> results in host memory
    pkd2 = PKD(device=local, fPeriod=p.period, pinned=True)
    pkd2.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
    ms_build = 0.0
    w2 = min(a.warmup, 2)
    for it in range(w2 + a.steps):
        if it == w2:
            torch.cuda.synchronize()
            f0 = time.perf_counter()
        pkd2.pkdBuildBinaryDevice(8, theta)
        if it >= w2:
            ms_build += pkd2.pkdBuildInfo()[2]
        st2 = pkd2.pkdGravAll(g, *outs, accumulate=False)
    torch.cuda.synchronize()
    fp_s = (time.perf_counter() - f0) / a.steps
    assert st2["dPartSum"] + st2["dCellSum"] + st2["dSoftSum"] == inter  # same tree -> same lists
    fromp = {"value": inter / fp_s, "unit": UNIT, "ms_per_step": fp_s * 1e3, "tree_build_device_ms": ms_build / a.steps,
             "h2d_bytes_per_step": 5 * 8 * n, "d2h_bytes_per_step": 6 * 8 * n + 4 * n,
             "what": "host particles (any order) -> gg_build_local (pkdBuildBinary on the device) -> gg_gravity -> host "
                     "arrays; the host tree build of the e2e leg is not needed"}
    pkd2.close()
    # ---- SURVEY 8f ranks 2+3: the particle store resident in HBM; one kick-drift-kick step = kick, drift, tree build,
    #      gravity, kick, grav-step, all on the device, no per-step particle traffic
    pkd3 = PKD(device=local, fPeriod=p.period)
    zero = np.zeros(n)
    pkd3.pkdLoadResident(p.x, p.y, p.z, zero, zero, zero, p.m, p.h)
    pkd3.pkdBuildBinaryResident(8, theta)
    pkd3.pkdGravAll(g, download=False)
    dstep = 1e-4  # small: the workload (list lengths) stays that of the configuration
    inter_res = 0.0
    for it in range(w2 + a.steps):
        if it == w2:
            torch.cuda.synchronize()
            r0 = time.perf_counter()
        pkd3.pkdKick(1.0, 0.5 * dstep)
        pkd3.pkdDrift(dstep, (0.0, 0.0, 0.0), g.bPeriodic)
        pkd3.pkdBuildBinaryResident(8, theta)
        st3 = pkd3.pkdGravAll(g, download=False)
        pkd3.pkdKick(1.0, 0.5 * dstep)
        dt_min = pkd3.pkdGravStep(0.2)
        if it >= w2:
            inter_res += st3["dPartSum"] + st3["dCellSum"] + st3["dSoftSum"]
    torch.cuda.synchronize()
    rs_s = (time.perf_counter() - r0) / a.steps
    fromp["resident_kdk_step"] = {
        "value": inter_res / a.steps / rs_s, "unit": UNIT, "ms_per_step": rs_s * 1e3, "dt_min": dt_min,
        "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16,
        "what": "gg_state_kick, gg_state_drift, gg_state_build, gg_gravity(NO_DOWNLOAD), gg_state_kick, "
                "gg_state_gravstep on the device-resident store"}
    pkd3.close()
    m["from_particles"] = fromp
    return m


def single_summary(m, a, spec, peaks):
    """The JSON fields of a one-GPU measurement (used for the main line at N = 1 and for the configs[1] extra)."""
    acc, st, g = m["acc"], m["st"], m["g"]
    k = a.steps
    ms_step, eval_ms = acc["ms"] / k, acc["eval"] / k
    roof = roofline_of(g, m["peak_tf"], eval_ms, st["dFlop"] - st["dFlopEwald"], m["n"], m["pkd"].tree.nNodes,
                       st["nListEntries"], spec, peaks)
    roof["step_breakdown_ms"] = {"k_walk": acc["walk"] / k, "scan+k_scatter": (acc["tree"] - acc["eval"] - acc["walk"]) / k,
                                 "k_eval": eval_ms, "k_ewald": acc["ewald"] / k,
                                 "other (memsets, k_stats, gaps)": ms_step - (acc["tree"] + acc["ewald"]) / k}
    return {"value": m["inter"] / (ms_step * 1e-3), "ms_per_step": ms_step, "roofline": roof,
            "e2e": {"value": m["inter"] / m["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": int(m["h2d"]),
                    "d2h_bytes_per_step": int(m["d2h"]), "ms_per_step": m["e2e_s"] * 1e3},
            "gpu_launches": int(acc["launches"]), "interactions_per_step": m["inter"],
            "host_tree_build_s": m["t_tree"], "wall_ms_per_step_resident": m["wall"] / k * 1e3}


def run_single(a):
    import torch
    from gasoline_b200 import build as gbuild
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(0)
    gbuild.build()
    spec = workload_spec(a)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    m = measure_single(spec, a, 0, full=not a.no_extra)
    s = single_summary(m, a, spec, peaks)
    out = {"metric": METRIC, "value": s["value"], "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": s["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32 (FP64 opening tests, FP64 accumulation across lanes, FP64 Ewald)", "data": "synthetic",
           "config": config_of(spec, 1), "clocks": m["clocks"], "e2e": s["e2e"], "gpu_launches": s["gpu_launches"],
           "roofline": s["roofline"], "interactions_per_step": s["interactions_per_step"],
           "host_tree_build_s": s["host_tree_build_s"], "wall_ms_per_step_resident": s["wall_ms_per_step_resident"]}
    if "from_particles" in m:
        out["e2e_from_particles"] = m["from_particles"]
    if not a.no_parity:
        try:
            counts = m["pkd"].pkdBucketCounts()
            res = dict(acc=m["outs"][0], pot=m["outs"][1], fWeight=m["outs"][3])
            # the GPU results are in tree order of the host-built tree; the checker rebuilds the same tree from the
            # same particles (tests pin the builder to the reference's tree)
            out["parity"] = parity_check(m["p"], m["g"], m["theta"], [np.arange(m["p"].n)], res, counts, a.parity_buckets)
        except Exception as e:
            out["parity"] = {"ok": False, "error": f"{type(e).__name__}: {e}"[:300]}
    m["pkd"].close()
    del m
    if not a.no_extra:
        if spec != CONFIGS1_SPEC:
            try:  # BASELINE.json configs[1] beside the headline configuration
                a1 = argparse.Namespace(**vars(a))
                m1 = measure_single(CONFIGS1_SPEC, a1, 0, full=False)
                s1 = single_summary(m1, a1, CONFIGS1_SPEC, peaks)
                s1.update(config=config_of(CONFIGS1_SPEC, 1), clocks=m1["clocks"], unit=UNIT)
                m1["pkd"].close()
                out["configs1"] = s1
            except Exception as e:
                out["configs1"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        out["e2e_pkdGravAll"] = c_host_leg(spec, 1, a.steps, a.warmup)
    if not a.no_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline", "--workload", spec, "--seconds",
                                str(a.cpu_seconds)], cwd=ROOT, capture_output=True, text=True, timeout=1200)
            cb = json.loads(r.stdout.strip().splitlines()[-1])
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            out["cpu_baseline"]["how"] = ("the stock pkdGravAll on one shared tree: P forked workers with disjoint sink sets "
                                          "-- no remote walks, no domain imbalance, so it flatters the reference")
        except Exception as e:  # the GPU line must still be printed
            out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(out), flush=True)


def run_multi(a):
    import torch
    import torch.distributed as dist
    from gasoline_b200 import build as gbuild, domain
    from gasoline_b200.pkd import pinned_empty

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_cpus(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        gbuild.build()
    dist.barrier()
    spec = workload_spec(a)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    p, g, theta, ic_dir = make_ic_shared(spec, rank, world, barrier)
    big = p.n > 40_000_000  # (C5: the parity checker and the C-host leg would need the whole box on one host process)

    # ---- the domains: the reference's ORB decomposition (pstDomainDecomp) with the per-rank services on the devices
    t0 = time.time()
    orb = {}
    if a.host_orb:
        idx = domain.orb_decompose(p.x, p.y, p.z, world)[rank]
    else:
        # the root finder of every level as ONE collective below the C ABI (gg_orb_bisect_all: the trial answers travel
        # device to device over NCCL); GG_ORB_COLLECTIVE=0: the trial answers through torch.distributed and the host
        orb_coll = os.environ.get("GG_ORB_COLLECTIVE", "1") != "0"
        idx = domain.device_orb_share(p, rank, world, local, "cuda", timing=orb, collective=orb_coll)
        orb["bisection"] = ("gg_orb_bisect_all (in-stream NCCL all-gather per trial)" if orb.get("collective")
                            else "host loop, torch.distributed all-gathers per trial")
    orb["seconds"] = time.time() - t0
    pkd, exchange = domain.setup_rank(p, theta, rank, world, local, idx=idx)
    t_tree = time.time() - t0 - orb["seconds"]
    n = pkd.nLocal
    pkd.device_moments = True
    exchange(top=True, let=g)  # upload, top-tree assembly (part of the tree build, like pstBuildTree's interior branch), first exchange
    peak_tf, _ = pkd.measure_fp32_peak()

    # ---- device-resident timing: per step gg_set_top + gg_exchange + gg_gravity
    for _ in range(a.warmup):
        exchange(top=False, let=g, upload=False)
        pkd.pkdGravAll(g, download=False)
    # clocks of all the job's GPUs from ONE poller on rank 0 (see ClockSampler)
    sampler = ClockSampler(",".join(str(i) for i in range(world)), period_ms=100) if rank == 0 else None
    barrier()
    if sampler is not None:
        sampler.start()
    barrier()
    acc = dict(ms=0.0, tree=0.0, ewald=0.0, eval=0.0, walk=0.0, launches=0, export=0.0, transfer=0.0, ingest=0.0)
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        pkd.flush_l2()
        pkd.timer_start()
        exchange(top=False, let=g, upload=False)
        st = pkd.pkdGravAll(g, download=False)
        acc["ms"] += pkd.timer_stop()
        xs = exchange.driver.stats
        acc["export"] += xs["msExport"]; acc["transfer"] += xs["msTransfer"]; acc["ingest"] += xs["msIngest"]
        acc["tree"] += st["msTree"]; acc["ewald"] += st["msEwald"]; acc["eval"] += st["msEval"]; acc["walk"] += st["msWalk"]
        acc["launches"] += st["nKernelLaunches"] + xs["nKernelLaunches"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler is not None else None
    inter = st["dPartSum"] + st["dCellSum"] + st["dSoftSum"]

    # ---- end to end: host tree + particles in (gg_set_local), exchange, evaluation, results in pinned host arrays
    outs = (pinned_empty((n, 3)), pinned_empty(n), pinned_empty(n), pinned_empty(n))
    for _ in range(min(a.warmup, 2)):
        exchange(top=False, let=g)
        pkd.pkdGravAll(g, *outs, accumulate=False)
    barrier()
    e0 = time.perf_counter()
    for _ in range(a.steps):
        exchange(top=False, let=g)
        pkd.pkdGravAll(g, *outs, accumulate=False)
    barrier()
    e2e_s = (time.perf_counter() - e0) / a.steps
    e2e_phases = {k: v * 1e3 for k, v in exchange.driver.timing.items()}
    h2d, d2h = pkd.upload_bytes(), 6 * 8 * n
    counts0 = pkd.pkdBucketCounts() if rank == 0 else None

    # ---- parity self-check (rank 0 evaluates a sample of its buckets with the CPU oracle on all ranks' trees)
    parity = None
    if big and not a.no_parity and rank == 0:
        parity = {"ok": None, "skipped": f"{p.n} particles: the checker would need every rank's tree in one host process; "
                                         "parity at this size is covered by the property tests (tests/test_gpu_fullsize.py)"}
    if not a.no_parity and not big:
        lens = torch.zeros(world, dtype=torch.int64, device="cuda")
        lens[rank] = len(idx)
        dist.all_reduce(lens)
        cap = int(lens.max().item())
        mine = torch.full((cap,), -1, dtype=torch.int64, device="cuda")
        mine[:len(idx)] = torch.from_numpy(np.asarray(idx, np.int64)).cuda()
        allidx = torch.empty(world * cap, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allidx, mine)
        if rank == 0:
            ai = allidx.cpu().numpy().reshape(world, cap)
            idx_all = [ai[r][: int(lens[r].item())] for r in range(world)]
            try:
                parity = parity_check(p, g, theta, idx_all, dict(acc=outs[0], pot=outs[1], fWeight=outs[3]), counts0,
                                      a.parity_buckets)
            except Exception as e:
                parity = {"ok": False, "error": f"{type(e).__name__}: {e}"[:300]}
        del allidx, mine
        barrier()

    # ---- from particles: every rank's tree built on its GPU (gg_build_local), top tree assembled from device summaries
    from_particles = None
    if not a.no_extra:
        pkd4, exchange4 = domain.setup_rank(p, theta, rank, world, local, device_build=True, idx=idx)
        n4 = pkd4.nLocal
        w2 = min(a.warmup, 2)
        for it in range(w2 + a.steps):
            if it == w2:
                barrier()
                f0 = time.perf_counter()
            exchange4(let=g, rebuild=True)
            st4 = pkd4.pkdGravAll(g, *outs, accumulate=False)
        barrier()
        fp_s = (time.perf_counter() - f0) / a.steps
        fp_vals = torch.tensor([fp_s], dtype=torch.float64, device="cuda")
        fp_sums = torch.tensor([st4["dPartSum"] + st4["dCellSum"] + st4["dSoftSum"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(fp_vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(fp_sums, op=dist.ReduceOp.SUM)
        from_particles = {"value": fp_sums.item() / fp_vals.item(), "unit": UNIT, "ms_per_step": fp_vals.item() * 1e3,
                          "tree_build_device_ms_rank0": pkd4.pkdBuildInfo()[2],
                          "phases_ms_rank0": {k: v * 1e3 for k, v in exchange4.driver.timing.items()},
                          "h2d_bytes_per_step_rank0": 5 * 8 * n4, "d2h_bytes_per_step_rank0": 6 * 8 * n4 + 4 * n4,
                          "what": "per rank: host particles -> gg_build_local -> root summaries / top tree (gg_comm_allgather) "
                                  "-> gg_exchange (pruned LET over NCCL) -> gg_gravity -> host arrays"}
        pkd4.close()

    # ---- reduce over ranks: time = max, work = sum
    vals = torch.tensor([acc["ms"], acc["tree"], e2e_s, wall, acc["eval"], acc["walk"], acc["ewald"], acc["export"],
                         acc["transfer"], acc["ingest"]], dtype=torch.float64, device="cuda")
    sums = torch.tensor([inter, st["dFlop"] - st["dFlopEwald"], float(acc["launches"]), float(n), float(pkd.tree.nNodes),
                         float(h2d), float(d2h), st["nListEntries"]], dtype=torch.float64, device="cuda")
    mins = vals.clone()
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_max, tree_max, e2e_max, wall_max, eval_max, walk_max, ewald_max, exp_max, xfer_max, ing_max = vals.tolist()
    inter_all, flop_all, launches_all, n_all, nodes_all, h2d_all, d2h_all, entries_all = sums.tolist()
    c_host = None
    if not a.no_extra and not big:
        # the reference BINARY on `world` pthread-MDL ranks, rank r on GPU r: rank 0 launches it, the others keep off the GPUs
        flag = os.path.join(tempfile.gettempdir(), f"gg_bench_{os.environ.get('MASTER_PORT', '0')}.done")
        if rank == 0 and os.path.exists(flag):
            os.remove(flag)
        barrier()
        if rank == 0:
            c_host = c_host_leg(spec, world, a.steps, a.warmup)
            open(flag, "w").write("done")
        else:
            t_wait = time.time()
            while not os.path.exists(flag) and time.time() - t_wait < 1200:
                time.sleep(0.05)
        barrier()
        if rank == 0:
            os.remove(flag)

    if rank == 0:
        k = a.steps
        ms_step = ms_max / k
        eval_ms = eval_max / k
        roof = roofline_of(g, peak_tf, eval_ms, flop_all / world, int(n_all / world), int(nodes_all / world),
                           entries_all / world, spec, peaks)
        roof["ms_per_launch_note"] = "slowest rank's launch; flops = the per-GPU mean"
        roof["step_breakdown_ms"] = {"gg_exchange: LET export": exp_max / k, "gg_exchange: NCCL send/recv": xfer_max / k,
                                     "gg_exchange: ingest": ing_max / k, "k_walk": walk_max / k,
                                     "scan+k_scatter": (tree_max - eval_max - walk_max) / k, "k_eval": eval_ms,
                                     "k_ewald": ewald_max / k, "max over ranks of each phase": True,
                                     "note": "k_ewald runs on the side stream beside gg_exchange's phases (gg_early_ewald): it is "
                                             "not on the critical path, and the exchange phases' times include the contention"}
        out = {"metric": METRIC, "value": inter_all / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f32 (FP64 opening tests, FP64 accumulation across lanes, FP64 Ewald)",
               "data": "synthetic", "config": config_of(spec, world), "clocks": clocks,
               "e2e": {"value": inter_all / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
                       "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_max * 1e3, "phases_ms_rank0": e2e_phases},
               "gpu_launches": int(launches_all), "roofline": roof, "interactions_per_step": inter_all,
               "host_tree_build_s": t_tree, "wall_ms_per_step_resident": wall_max / k * 1e3,
               "rank_balance": {"ms_per_step_min": mins[0].item() / k, "ms_per_step_max": ms_step,
                                "k_eval_ms_min": mins[4].item() / k, "k_eval_ms_max": eval_ms},
               "domain_decomposition": dict(orb, what="host median split" if a.host_orb else
                                            "pstDomainDecomp with the per-rank services on the devices (gg_orb_*)"),
               "let_bytes_rank0": {"sent": exchange.driver.let_bytes[0], "received": exchange.driver.let_bytes[1],
                                   "whole_domain": int(exchange.driver.stats["bytesWholeDomain"])},
               "comm": pkd.commInfo()}
        if numa is not None:
            out["config"]["cpu_binding"] = f"rank bound to the {numa} cores local to its GPU (NVML affinity)"
        if parity is not None:
            out["parity"] = parity
        if from_particles is not None:
            out["e2e_from_particles"] = from_particles
        if c_host is not None:
            out["e2e_pkdGravAll"] = c_host
        print(json.dumps(out), flush=True)
    pkd.close()
    barrier()
    if rank == 0:
        import shutil
        shutil.rmtree(ic_dir, ignore_errors=True)
    dist.destroy_process_group()


def _json_only_stdout():
    """Libraries below us (NCCL's version banner, ...) write to file descriptor 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the whole run and give Python's sys.stdout a private copy of the real one."""
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    args = parse_args()
    _json_only_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_multi(args)
    else:
        run_single(args)
