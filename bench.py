#!/usr/bin/env python
"""bench.py -- tree-gravity force evaluation (Gasoline's pkdGravAll path) on B200: interactions/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload plummer:1000000:0.7]

A "step" is ONE force evaluation = one pkdGravAll (pkd.c:2868) over every sink bucket of the workload: tree walks,
list evaluation and (periodic workloads) the Ewald correction.  The metric is the reference's own interaction
count (pkd.c:2945-2949: particle-list + intra-bucket pairs + Newtonian-cell + softened-cell entries summed over
active sinks; Ewald terms are not interactions) divided by time.

  value     device time (CUDA events recorded by the library on the stream it launches on) with the tree and the
            particles already resident in HBM; L2 is flushed between timed iterations.
  e2e       the same through the public host API with HOST buffers: every step re-ingests the host's tree +
            particles from pinned memory (gg_set_local: the reference frees and rebuilds kdNodes before every
            gravity call, pkd.c:2636-2642; the cells' moments are formed on the device instead of transferred),
            runs the kernels and delivers a, fPot, dtGrav, fWeight into pinned host arrays (stored by the kernels
            as each sink bucket finishes; d2h_bytes_per_step counts them).
  roofline  the dominant kernel, k_eval<4> (list evaluation), against the FP32 FMA pipe (this is FP32 CUDA-core + SFU
            work, not HBM- or tensor-bound -- DESIGN.md 5): achieved = the reference's own flop score of the lists it
            evaluated (grav.c:246-247) / the kernel's duration (CUDA events recorded around the launch on the
            library's stream); peak = dependent-FFMA microbenchmark measured in this run on this GPU.  The HBM view
            (algorithmic bytes / time vs MEASURED_PEAKS.json hbm_gbs) is reported beside it; traffic = DRAM bytes of
            one k_eval launch from the committed ncu capture (profiles/).
  cpu_baseline / --impl reference
            the reference's own compiled pkdGravAll on the host cores (oracle/cpu_baseline.py).

Workload at N=1 = BASELINE.json configs[1]: isolated Plummer sphere, 1M particles, theta=0.7, no Ewald.
N>1: weak scaling -- N x 1M-particle Plummer sphere split into N ORB domains, one per GPU (gasoline_b200/domain.py).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "particle interactions/s (tree gravity, pkdGravAll)"
UNIT = "interactions/s"
PER_GPU_PARTICLES = 1_000_000
THETA = 0.7


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="plummer:N:theta | periodic:n:theta (default: configs[1])")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--device-orb", action="store_true",
                    help="N > 1: the from-particles leg takes its domains from the reference's ORB decomposition run on "
                         "the devices (gg_orb_*) instead of the host-side median split")
    return ap.parse_args()


def workload_spec(a) -> str:
    if a.workload:
        return a.workload
    return f"plummer:{PER_GPU_PARTICLES * max(a.gpus, 1)}:{THETA}"


def config_of(spec: str, n_gpus: int) -> dict:
    kind, n, theta = spec.split(":")
    if kind == "plummer":
        name = f"isolated Plummer sphere, {int(n)} particles, theta={theta}, no Ewald, nBucket=8, iOrder=4 (hexadecapole)"
    else:
        name = f"periodic box {n}^3 particles, theta={theta}, nReplicas=1, Ewald on, nBucket=8, iOrder=4"
    return {"workload": name, "spec": spec, "particles_per_gpu": int(n) ** (1 if kind == "plummer" else 3) // n_gpus,
            "parallelism": "1 GPU" if n_gpus == 1 else f"{n_gpus} ORB domains, one per GPU, LET exchange over NCCL",
            "l2": "flushed between timed iterations (384 MB memset)"}


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
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
           "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config_of(spec, a.gpus),
           "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                            "sample": r["sample"]},
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
        return d["dram_bytes_per_launch"] if d.get("workload") == spec else None
    except (OSError, KeyError, ValueError):
        return None


def bind_to_gpu_cpus(index: int):
    """Pin this rank to the CPU cores NVML reports as local to its GPU (what an MPI launcher's binding does for the
    reference): the rank's pinned host buffers are then first-touched on the GPU's NUMA node and its host<->device
    copies do not cross sockets.  Returns the number of cores in the mask, or None when NVML cannot tell."""
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


def run_ours(a):
    import torch
    import torch.distributed as dist
    from gasoline_b200 import build as gbuild, ics
    from gasoline_b200.pkd import PKD, GravityParams, pinned_empty
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_cpus(local) if world > 1 else None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gbuild.build()
    spec = workload_spec(a)
    kind, n_s, theta_s = spec.split(":")
    theta = float(theta_s)
    if kind == "plummer":
        p = ics.plummer(int(n_s))
        g = GravityParams(nReps=0, bPeriodic=0, bEwald=0)
    else:
        p = ics.periodic_box(int(n_s))
        g = GravityParams(nReps=1, bPeriodic=1, bEwald=1)

    t0 = time.time()
    if world == 1:
        pkd = PKD(device=local, fPeriod=p.period, pinned=True)
        pkd.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
        pkd.pkdBuildBinary(8, theta, 4)
        exchange = None
    else:
        from gasoline_b200 import domain
        pkd, exchange = domain.setup_rank(p, theta, rank, world, local)
    t_tree = time.time() - t0
    n = pkd.nLocal
    # gg_tree.mom = NULL: the cells' multipole moments (58 % of the tree bytes) are not transferred; the device forms
    # them from the particles while the walk runs (gg_moments.cu; forces identical, tests/test_gpu_device_moments.py)
    pkd.device_moments = True
    pkd.upload()
    if exchange is not None:
        exchange(let=g)  # top tree + pruned locally-essential trees (NCCL all-to-all), once before the resident timing
    peak_tf, _ = pkd.measure_fp32_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(a.warmup):
        pkd.pkdGravAll(g, download=False)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ms_total = ms_tree = ms_ewald = ms_eval = ms_walk = 0.0
    launches = 0
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        pkd.flush_l2()
        st = pkd.pkdGravAll(g, download=False)
        ms_total += st["msTotal"]; ms_tree += st["msTree"]; ms_ewald += st["msEwald"]
        ms_eval += st["msEval"]; ms_walk += st["msWalk"]
        launches += st["nKernelLaunches"]
    barrier()
    wall_resident = time.perf_counter() - wall0
    clocks = sampler.stop()
    inter = st["dPartSum"] + st["dCellSum"] + st["dSoftSum"]

    # ---- end to end through the host API: host buffers in, host buffers out, every step
    out_a, out_p, out_d, out_w = (pinned_empty((n, 3)), pinned_empty(n), pinned_empty(n), pinned_empty(n))
    for _ in range(min(a.warmup, 2)):
        if exchange is not None:
            exchange(top=False, let=g)
        else:
            pkd.upload()
        pkd.pkdGravAll(g, out_a, out_p, out_d, out_w, accumulate=False)
    barrier()
    e0 = time.perf_counter()
    for _ in range(a.steps):
        if exchange is not None:
            # this rank's own upload (gg_set_local) + the NCCL tree exchange; the top tree belongs to the host's tree
            # build (pstBuildTree's interior branch), which is outside the timed region like pkdBuildBinary
            exchange(top=False, let=g)
        else:
            pkd.upload()
        pkd.pkdGravAll(g, out_a, out_p, out_d, out_w, accumulate=False)
    barrier()
    e2e_s = time.perf_counter() - e0
    h2d = pkd.upload_bytes()
    d2h = 6 * 8 * n

    # ---- the same step with the tree built on the device (gg_build_local): particles in input order
    #      in pinned host memory -> device tree build (bit-identical tree) -> pkdGravAll -> results in host memory
    from_particles = None
    if world == 1:
        pkd2 = PKD(device=local, fPeriod=p.period, pinned=True)
        pkd2.pkdLoadParticles(p.x, p.y, p.z, p.m, p.h)
        ms_build = 0.0
        for it in range(min(a.warmup, 2) + a.steps):
            if it == min(a.warmup, 2):
                barrier()
                f0 = time.perf_counter()
            pkd2.pkdBuildBinaryDevice(8, theta)
            if it >= min(a.warmup, 2):
                ms_build += pkd2.pkdBuildInfo()[2]
            st2 = pkd2.pkdGravAll(g, out_a, out_p, out_d, out_w, accumulate=False)
        barrier()
        fp_s = (time.perf_counter() - f0) / a.steps
        assert st2["dPartSum"] + st2["dCellSum"] + st2["dSoftSum"] == inter  # same tree -> same lists
        from_particles = {"value": inter / fp_s, "unit": UNIT, "ms_per_step": fp_s * 1e3,
                          "tree_build_device_ms": ms_build / a.steps, "h2d_bytes_per_step": 5 * 8 * n,
                          "d2h_bytes_per_step": 6 * 8 * n + 4 * n,
                          "what": "host particles (any order) -> gg_build_local (pkdBuildBinary on the device) -> "
                                  "gg_gravity -> host arrays; the host tree build of the e2e leg is not needed"}
        pkd2.close()
        # ---- SURVEY 8f ranks 2+3: the particle store resident in HBM; one kick-drift-kick step = kick, drift, tree
        #      build, gravity, kick, grav-step, all on the device, no per-step particle traffic
        pkd3 = PKD(device=local, fPeriod=p.period)
        zero = np.zeros(n)
        pkd3.pkdLoadResident(p.x, p.y, p.z, zero, zero, zero, p.m, p.h)
        pkd3.pkdBuildBinaryResident(8, theta)
        pkd3.pkdGravAll(g, download=False)
        dstep = 1e-4  # small: the workload (list lengths) stays that of the configuration
        inter_res = 0.0
        for it in range(min(a.warmup, 2) + a.steps):
            if it == min(a.warmup, 2):
                barrier()
                r0 = time.perf_counter()
            pkd3.pkdKick(1.0, 0.5 * dstep)
            pkd3.pkdDrift(dstep, (0.0, 0.0, 0.0), g.bPeriodic)
            pkd3.pkdBuildBinaryResident(8, theta)
            st3 = pkd3.pkdGravAll(g, download=False)
            pkd3.pkdKick(1.0, 0.5 * dstep)
            dt_min = pkd3.pkdGravStep(0.2)
            if it >= min(a.warmup, 2):
                inter_res += st3["dPartSum"] + st3["dCellSum"] + st3["dSoftSum"]
        barrier()
        rs_s = (time.perf_counter() - r0) / a.steps
        from_particles["resident_kdk_step"] = {
            "value": inter_res / a.steps / rs_s, "unit": UNIT, "ms_per_step": rs_s * 1e3, "dt_min": dt_min,
            "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16,
            "what": "gg_state_kick, gg_state_drift, gg_state_build, gg_gravity(NO_DOWNLOAD), gg_state_kick, "
                    "gg_state_gravstep on the device-resident store"}
        pkd3.close()

    if world > 1:
        # ---- the same per-step work with every rank's tree built on its GPU (gg_build_local) instead of handed over by
        #      the host: particles H2D, tree build, root summaries + top tree (two small all-gathers), pruned LET
        #      exchange, force evaluation, results to the host
        from gasoline_b200 import domain
        orb_info = None
        if a.device_orb:
            orb_t = {}
            domain.device_orb_share(p, rank, world, local, "cuda")  # warm-up (context, NCCL channels)
            barrier()
            o0 = time.perf_counter()
            domain.device_orb_share(p, rank, world, local, "cuda", timing=orb_t)
            barrier()
            orb_info = dict(orb_t, total_ms=(time.perf_counter() - o0) * 1e3,
                            what="pstDomainDecomp on the devices: per-rank chunk H2D, bisection trials (gg_orb_weight + one "
                                 "all-gather each), destinations, all-to-all of the particle indices (rank 0's clock)")
        pkd4, exchange4 = domain.setup_rank(p, theta, rank, world, local, device_build=True, device_orb=a.device_orb)
        n4 = pkd4.nLocal
        o4 = (pinned_empty((n4, 3)), pinned_empty(n4), pinned_empty(n4), pinned_empty(n4))
        for it in range(min(a.warmup, 2) + a.steps):
            if it == min(a.warmup, 2):
                barrier()
                f0 = time.perf_counter()
            exchange4(let=g, rebuild=True)
            st4 = pkd4.pkdGravAll(g, *o4, accumulate=False)
        barrier()
        fp_s = (time.perf_counter() - f0) / a.steps
        fp_vals = torch.tensor([fp_s], dtype=torch.float64, device="cuda")
        fp_sums = torch.tensor([st4["dPartSum"] + st4["dCellSum"] + st4["dSoftSum"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(fp_vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(fp_sums, op=dist.ReduceOp.SUM)
        from_particles = {"value": fp_sums.item() / fp_vals.item(), "unit": UNIT, "ms_per_step": fp_vals.item() * 1e3,
                          "tree_build_device_ms_rank0": pkd4.pkdBuildInfo()[2],
                          "phases_ms_rank0": {k: v * 1e3 for k, v in exchange4.driver.timing.items()},
                          "what": "per rank: host particles -> gg_build_local -> root summaries / top tree (all-gathers) -> "
                                  "pruned LET exchange (NCCL all-to-all) -> gg_gravity -> host arrays"}
        if orb_info is not None:
            from_particles["orb_device"] = orb_info
            from_particles["particles_rank0"] = int(n4)
        pkd4.close()

    # ---- reduce over ranks: time = max, work = sum
    vals = torch.tensor([ms_total, ms_tree, e2e_s, wall_resident, ms_eval, ms_walk, ms_ewald], dtype=torch.float64,
                        device="cuda")
    sums = torch.tensor([inter, st["dFlop"] - st["dFlopEwald"], float(launches), float(n), float(pkd.tree.nNodes),
                         float(h2d), float(d2h), st["nListEntries"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_total_max, ms_tree_max, e2e_max, wall_res_max, ms_eval_max, ms_walk_max, ms_ewald_max = vals.tolist()
    inter_all, flop_tree_all, launches_all, n_all, nodes_all, h2d_all, d2h_all, entries_all = sums.tolist()

    if rank == 0:
        ms_step = ms_total_max / a.steps
        tree_ms = ms_tree_max / a.steps
        value = inter_all / (ms_step * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # roofline of the dominant kernel (k_eval), per launch, per GPU
        eval_ms = ms_eval_max / a.steps
        flop_per_launch = flop_tree_all / world
        ach_tf = flop_per_launch / (eval_ms * 1e-3) * 1e-12
        alg_bytes = algorithmic_bytes(int(n_all / world), int(nodes_all / world), entries_all / world)
        roof = {"bound": "fp32", "kernel": f"k_eval<{g.iOrder}>", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": ach_tf / peak_tf if peak_tf else None, "traffic": ncu_traffic(spec),
                "peak_source": "dependent-FFMA microbenchmark measured in this run (nominal 74.4 TFLOP/s at 1965 MHz)",
                "flops": "the reference's own score of the evaluated lists (grav.c:246-247: 38/particle, 82/soft "
                         "cell, 312/hexadecapole cell)",
                "hbm": {"achieved": alg_bytes / (eval_ms * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (eval_ms * 1e-3) * 1e-9 / hbm_peak, "algorithmic_bytes": alg_bytes,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback (B200_PROFILING.md)"},
                "ms_per_launch": eval_ms,
                "step_breakdown_ms": {"k_walk": ms_walk_max / a.steps, "scan+k_scatter": tree_ms - eval_ms - ms_walk_max / a.steps,
                                      "k_eval": eval_ms, "k_ewald": ms_ewald_max / a.steps,
                                      "other (task list, memsets, k_stats)": ms_step - tree_ms - ms_ewald_max / a.steps}}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_step, "higher_is_better": True,
               "scaling": "strong" if a.workload else "weak",  # default: 1 M particles per GPU; --workload fixes the total
               "vs_baseline": None,
               "dtype": "f32 (FP64 opening tests, FP64 accumulation across lanes, FP64 Ewald)", "data": "synthetic",
               "config": config_of(spec, world), "clocks": clocks,
               "e2e": {"value": inter_all / (e2e_max / a.steps), "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
                       "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_max / a.steps * 1e3},
               "gpu_launches": int(launches_all), "roofline": roof,
               "interactions_per_step": inter_all, "host_tree_build_s": t_tree,
               "wall_ms_per_step_resident": wall_res_max / a.steps * 1e3}
        if from_particles is not None:
            out["e2e_from_particles"] = from_particles
        if numa is not None:
            out["config"]["cpu_binding"] = f"rank bound to the {numa} cores local to its GPU (NVML affinity)"
        if exchange is not None:
            out["exchange_phases_ms_rank0"] = {k: v * 1e3 for k, v in exchange.driver.timing.items()}
            out["let_bytes_rank0"] = {"sent": exchange.driver.let_bytes[0], "received": exchange.driver.let_bytes[1],
                                      "whole_domain": pkd.export_size()[0]}
        if world == 1 and not a.no_cpu_baseline:
            try:
                r = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline", "--workload", spec, "--seconds",
                                    str(a.cpu_seconds)], cwd=ROOT, capture_output=True, text=True, timeout=900)
                cb = json.loads(r.stdout.strip().splitlines()[-1])
                out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the GPU line must still be printed
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                       "sample": f"failed: {e}"}
        print(json.dumps(out), flush=True)
    pkd.close()
    if world > 1:
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
    else:
        run_ours(args)
