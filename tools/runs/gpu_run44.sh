#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loops.py -q -m gpu --tb=short -p no:cacheprovider -x -k "pair or ddim" > gpurun_out/pytest_pair.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -15 gpurun_out/pytest_pair.log
