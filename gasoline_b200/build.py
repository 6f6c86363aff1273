"""Build the CUDA extension in-tree: gasoline_b200/lib/libgasoline_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the built library travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libgasoline_b200.so")
SOURCES = ["gg_api.cu", "gg_tree_kernel.cu", "gg_ewald.cu", "gg_moments.cu", "gg_tree_gpu.cu", "gg_state.cu", "gg_orb.cu", "gg_comm.cu", "gg_tree_build.cpp"]
# Host code is compiled without FP contraction: the bit-exact tree build (gg_tree_build.cpp) and the top-tree arithmetic
# must round every product and sum like the reference's x86-64 build, also on hosts whose compiler fuses by default
# (aarch64).  Device code: the translation units whose results are claimed bit-identical to the reference get
# -fmad=false on top of their explicit __d*_rn intrinsics; the force kernels (k_eval, k_ewald, moments) keep FMA.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-pthread,-ffp-contract=off"]
NO_FMAD = {"gg_tree_gpu.cu", "gg_state.cu", "gg_orb.cu"}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "gasoline_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.rsplit(".", 1)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + (["-fmad=false"] if src in NO_FMAD else []) + \
              (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs +
                          ["-lcudart", "-lpthread", "-ldl"])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
