"""bench.py's e2e_pkdGravAll leg: the reference's own C HOST (compiled from /root/reference where it lies, oracle/Makefile)
with the product's pkdGravAll + pkdBuildBinary + pkdCalcRoot link-substituted (gasoline_b200/csrc/pkd_gravall_shim.c) --
what a Gasoline user gets per force evaluation after switching the link line.

  --ranks 1   in-process: oracle/_ref/libgasref_gpu.so driven like msrBuildTree + msrGravity drive it (oracle/ref_api.c):
              pstBuildTree (the GPU builds the tree, the shim permutes pStore and fills kdNodes), then pstGravity
              (a) on the tree the device already holds (only ACTIVE flags go up) and (b) with GG_SHIM_FORCE_UPLOAD=1,
              i.e. flattening PARTICLE / KDN into the SoA views and uploading them as for a host-built tree.
  --ranks N   the reference BINARY (main.c, master.c, pst.c ... + pthread MDL stand-in) on N thread ranks, rank r on
              GPU r, reading a Tipsy file, doing ITS OWN domain decomposition and top tree; per step it prints
              "Gravity Calculated, Wallclock: ..." (master.c:5935) around pstGravity, which is the number reported.
Prints one JSON line.  The host code above the ABI is the reference's; nothing of oracle/ is on the measured GPU path
except as that host."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

UNIT = "interactions/s"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="periodic:256:0.5")
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    from gasoline_b200 import ics
    from oracle import reflib
    kind, n_s, theta_s = a.workload.split(":")
    theta = float(theta_s)
    p = ics.plummer(int(n_s)) if kind == "plummer" else ics.periodic_box(int(n_s))
    per = 1 if p.periodic else 0
    out = {"unit": UNIT, "ranks": a.ranks, "workload": a.workload}
    if a.ranks == 1:
        if not reflib.gpu_host_available():
            print(json.dumps(dict(out, value=None, error="oracle/_ref/libgasref_gpu.so not built")))
            return
        os.environ["GG_SHIM_DEVICE_TREE"] = "1"
        r = reflib.RefGravity(p, gpu_host=True)
        tb = [r.build_tree(8, theta, 4) for _ in range(1 + min(a.warmup, 1))]
        res = None
        for _ in range(a.warmup):
            res = r.gravity(per, per, 4, per, 4)
        resident = [r.gravity(per, per, 4, per, 4)["seconds"] for _ in range(a.steps)]
        os.environ["GG_SHIM_FORCE_UPLOAD"] = "1"
        r.gravity(per, per, 4, per, 4)
        upload = []
        for _ in range(a.steps):
            res = r.gravity(per, per, 4, per, 4)
            upload.append(res["seconds"])
        r.close()
        inter = res["dPartSum"] + res["dCellSum"] + res["dSoftSum"]
        s_up, s_res = float(np.mean(upload)), float(np.mean(resident))
        out.update(value=inter / s_up, ms_per_step=s_up * 1e3, interactions_per_step=inter,
                   pstGravity_ms_device_tree_resident=s_res * 1e3, value_device_tree_resident=inter / s_res,
                   pstBuildTree_ms_device=min(tb) * 1e3,
                   what="reference host (pstBuildTree -> pstGravity, pst.c:2906 / 3248) with the shim's pkdBuildBinary + pkdGravAll; "
                        "value = pstGravity incl. flattening PARTICLE/KDN, H2D, kernels, D2H and the += write-back into pStore; "
                        "*_device_tree_resident = the same call when the tree the shim built is still on the device")
        print(json.dumps(out))
        return
    gpu_bin = os.path.join(os.path.dirname(reflib.BIN_PATH), "gasoline_ref_gpu")
    if not os.path.exists(gpu_bin):
        print(json.dumps(dict(out, value=None, error="oracle/_ref/gasoline_ref_gpu not built")))
        return
    with tempfile.TemporaryDirectory() as tmp:
        ics.write_tipsy_native(os.path.join(tmp, "ic.tipsy"), p)
        nsteps = a.warmup + a.steps
        open(os.path.join(tmp, "run.param"), "w").write(
            f"achInFile = {tmp}/ic.tipsy\nachOutName = {tmp}/out\nbPeriodic = {per}\ndPeriod = 1\nnReplicas = {per}\n"
            f"bEwald = {per}\ndTheta = {theta}\ndTheta2 = {theta}\nnSteps = {nsteps}\ndDelta = 1e-7\niOutInterval = {10 * nsteps + 10}\n"
            "iLogInterval = 1\nbVStep = 1\nbDoDensity = 0\niBinaryOutput = 0\nbParaRead = 0\nbParaWrite = 0\nbOverwrite = 1\n"
            "bDoGravity = 1\niCheckInterval = 0\n")
        env = dict(os.environ, MDL_NTHREADS=str(a.ranks), GG_SHIM_DEVICE_TREE="1")
        env.pop("REF_DUMP", None)
        t0 = time.time()
        r = subprocess.run([gpu_bin, "run.param"], cwd=tmp, env=env, capture_output=True, text=True, timeout=1500)
        wall = time.time() - t0
    grav = [float(x) for x in re.findall(r"Gravity Calculated, Wallclock: ([0-9.eE+-]+) secs", r.stdout)]
    avg = re.findall(r"dPartAvg:([0-9.eE+-]+) dCellAvg:([0-9.eE+-]+) dSoftAvg:([0-9.eE+-]+)", r.stdout)
    if len(grav) < a.warmup + 1 or not avg:
        print(json.dumps(dict(out, value=None, error="no timing lines from the host binary", rc=r.returncode,
                              tail=(r.stdout[-400:] + r.stderr[-400:]))))
        return
    timed = grav[-a.steps:] if len(grav) > a.steps else grav[1:]
    s = float(np.mean(timed))
    inter = sum(float(v) for v in avg[-1]) * p.n  # the host prints per-particle averages over all ranks (master.c:5950)
    out.update(value=inter / s, ms_per_step=s * 1e3, interactions_per_step=inter, first_call_ms=grav[0] * 1e3,
               steps_timed=len(timed), host_wall_s=wall, rc=r.returncode,
               what=f"the reference binary on {a.ranks} pthread-MDL ranks (its own pstDomainDecomp, pstBuildTree, top tree), "
                    "rank r on GPU r, pkdBuildBinary + pkdGravAll through the shim, remote trees by gg_exchange over NCCL; "
                    "value = interactions / the host's own wallclock around pstGravity (master.c:5886-5935), mean of the last steps")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
