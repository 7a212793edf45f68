#!/usr/bin/env python
"""Build libncn.so (the C-ABI hot-path library) in-tree with nvcc for sm_100a.

    python normal-clustering-nerf_b200/csrc/build.py [--force] [--verbose]

One object per .cu (parallel), linked into ``normal-clustering-nerf_b200/libncn.so``.
No torch headers are used: the library is plain CUDA runtime + (dlopen'ed) NCCL.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(PKG, "libncn.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-I" + os.path.join(ROOT, "include")]
# per-file extra flags: the bit-exact integer/indexing kernels pin every fp32 rounding with
# explicit intrinsics and are additionally compiled without FMA contraction.
EXTRA = {
    "march.cu": ["-fmad=false"],
    "intersect.cu": ["-fmad=false"],
}


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _newest_dep():
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".h", ".py"))]
    deps.append(os.path.join(ROOT, "include", "ncn.h"))
    return max(os.path.getmtime(p) for p in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_dep():
        if verbose:
            print(f"[ncn build] up to date: {LIB}")
        return LIB
    hdr_time = max(os.path.getmtime(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith((".cuh", ".h")))
    hdr_time = max(hdr_time, os.path.getmtime(os.path.join(ROOT, "include", "ncn.h")),
                   os.path.getmtime(os.path.abspath(__file__)))
    jobs = []
    for f in sources():
        src = os.path.join(HERE, f)
        obj = os.path.join(OBJ, f + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time):
            continue
        cmd = [NVCC] + ARCH + COMMON + EXTRA.get(f, []) + os.environ.get("NCN_NVCC_EXTRA", "").split() + ["-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("[ncn build] failed: " + " ".join(cmd) + "\n" + r.stdout + "\n" + r.stderr)
        if verbose:
            print(" ".join(cmd[-3:]), "\n", r.stderr)
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, f + ".o") for f in sources()]
    link = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("[ncn build] link failed\n" + r.stdout + "\n" + r.stderr)
    if verbose:
        print(f"[ncn build] built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(LIB)
