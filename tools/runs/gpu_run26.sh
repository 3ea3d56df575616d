#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "groupnorm" --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gn.log 2>&1; echo "pytest_gn rc=$?"
tail -4 gpurun_out/pytest_gn.log
timeout 900 python tools/lanes_ab.py --variants "overlap=0,gnstream=0" "gnstream=0,head=0" "overlap=0" "head=0" "head=10" "head=5" "head=20" "pdlx=3" "pdlx=4" "pdlx=8" "pdlx=15" > gpurun_out/lanes_ab2.log 2> gpurun_out/lanes_ab2.err; echo "lanes_ab rc=$?"
cat gpurun_out/lanes_ab2.log; tail -5 gpurun_out/lanes_ab2.err
