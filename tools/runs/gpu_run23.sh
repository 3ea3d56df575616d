#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_loops.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_loops.log 2>&1; echo "pytest_loops rc=$?" >> gpurun_out/summary.txt
timeout 900 python tools/lanes_ab.py > gpurun_out/lanes_ab.log 2> gpurun_out/lanes_ab.err; echo "lanes_ab rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/pytest_loops.log
cat gpurun_out/lanes_ab.log; tail -5 gpurun_out/lanes_ab.err
