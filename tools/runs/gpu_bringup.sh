#!/bin/bash
# First-contact script for a fresh B200 box: structured GEMM diagnostic, then the gpu test files one process each
# (a hung kernel in one file must not take the others down), then smoke().
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 300 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1; echo "gemm_diag rc=$?" >> gpurun_out/summary.txt
for f in kernels unet loops; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_$f.log 2>&1
  echo "pytest_$f rc=$?" >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/gemm_diag.log
for f in kernels unet loops; do tail -3 gpurun_out/pytest_$f.log; done
tail -3 gpurun_out/smoke.log
