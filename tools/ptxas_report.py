"""Register / barrier / spill report of the tcgen05 kernels from `ptxas -v` (no GPU needed), with the flags of build.py.
Usage: python tools/ptxas_report.py > profiles/<name>.txt"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audioeditingcode_b200 import build as B  # noqa: E402

print("# ptxas -v, nvcc " + subprocess.run([B.NVCC, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
      + ", flags of audioeditingcode_b200/build.py: registers / named barriers / stack frame / spill stores / spill loads (bytes)")
for f in ("gemm_tcgen05.cu", "attn_tc.cu", "attn_kernels.cu", "norm_kernels.cu", "pc_kernels.cu"):
    with tempfile.TemporaryDirectory() as td:
        r = subprocess.run([B.NVCC, *B.FLAGS, "-Xptxas", "-v", "-c", os.path.join(B.CSRC, f), "-o", os.path.join(td, "o.o")],
                           capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr[-2000:])
    print(f"== {f}")
    for b in re.split(r"ptxas info\s+: Compiling entry function '", r.stderr)[1:]:
        dem = subprocess.run(["c++filt", b.split("'")[0]], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem.replace("(anonymous namespace)::", "").replace("aedit::", "").replace("void ", ""))
        m = re.search(r"Used (\d+) registers, used (\d+) barriers", b)
        sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
        print(f"{dem:48s} regs {m.group(1):>3s}  barriers {m.group(2):>2s}  stack {sp.group(1):>4s}  spill st {sp.group(2):>4s}  "
              f"spill ld {sp.group(3):>4s}")
