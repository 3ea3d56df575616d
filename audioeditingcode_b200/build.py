"""Build recipe for libaedit.so (hand-written sm_100a kernels behind the C ABI of include/aedit.h).

    python -m audioeditingcode_b200.build          # incremental: one object per .cu, then link

nvcc cross-compiles for sm_100a without a GPU.  The library is built IN-TREE
(audioeditingcode_b200/libaedit.so) so it travels to the GPU box with the repo snapshot; it links only
against the CUDA runtime (no torch, no cuBLAS/cuDNN).  The driver entry point for TMA descriptor
encoding is resolved at run time through cudaGetDriverEntryPoint.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libaedit.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-ffp-contract=off", "--expt-relaxed-constexpr", "-I", INCLUDE]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    mt = os.path.getmtime(obj)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps += [src, os.path.join(INCLUDE, "aedit.h")]
    return any(os.path.getmtime(d) > mt for d in deps)


def _compile(name, verbose):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name[:-3] + ".o")
    if not _stale(src, obj):
        return obj, ""
    cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {name}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda n: _compile(n, verbose), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
