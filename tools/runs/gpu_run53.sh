#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_loops.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_final_loops.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_final_loops.log
