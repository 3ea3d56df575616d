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
LIB = os.path.join(HERE, "libaedit.so")               # fp16 operands (default)
LIB_BF16 = os.path.join(HERE, "libaedit_bf16.so")     # -DAE_OPERAND_BF16 (AEDIT_OPERANDS=bf16)
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


def _compile(name, verbose, bf16=False):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name[:-3] + ("_bf16.o" if bf16 else ".o"))
    if not _stale(src, obj):
        return obj, ""
    cmd = [NVCC, *FLAGS, *(["-DAE_OPERAND_BF16"] if bf16 else []), "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {name}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(verbose: bool = False, force: bool = False, variants=("fp16", "bf16")) -> str:
    """Builds libaedit.so (fp16 operands) and libaedit_bf16.so (bf16 operands) from the same sources."""
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = sources()
    jobs = [(n, v == "bf16") for v in variants for n in srcs]
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        results = list(ex.map(lambda j: _compile(j[0], verbose, j[1]), jobs))
    if verbose:
        for _, log in results:
            if log:
                print(log)
    for v in variants:
        lib = LIB_BF16 if v == "bf16" else LIB
        objs = [o for (n, b), (o, _) in zip(jobs, results) if b == (v == "bf16")]
        if (not os.path.exists(lib)) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
            cmd = [NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
