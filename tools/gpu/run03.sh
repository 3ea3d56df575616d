#!/bin/bash
# round-2 GPU call 3: pc_drift device path parity + memcheck
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pc_drift.py -q -s > gpurun_out/r02_pc_drift_tests.log 2>&1; echo "pc tests rc=$?"
grep -E "cos\(|rel-L2|passed|failed|Error|assert" gpurun_out/r02_pc_drift_tests.log | tail -40
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pc_drift.py -q -k "open_loop or apply_drift" > gpurun_out/r02_sanitizer_memcheck_pc.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r02_sanitizer_memcheck_pc.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pc_drift.py -q -k "open_loop" > gpurun_out/r02_sanitizer_racecheck_pc.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r02_sanitizer_racecheck_pc.log
