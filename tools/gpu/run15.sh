#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -s -x -k "attention" > gpurun_out/r02_attn_tc_test.log 2>&1; echo "attn tests rc=$?"
grep -E "passed|failed|Error|error|assert" gpurun_out/r02_attn_tc_test.log | tail -6
timeout 400 python tools/attn_bench.py > gpurun_out/r02_attn_bench_v5.log 2>&1; cat gpurun_out/r02_attn_bench_v5.log | tail -10
