#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r02_pytest_v41.log 2>&1; echo "suite rc=$?"
grep -E "fd step|passed|failed|Error" gpurun_out/r02_pytest_v41.log | tail -12
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention_tcgen05" > gpurun_out/r02_sanitizer_memcheck_attn_tc.log 2>&1; echo "memcheck attn rc=$?"; tail -3 gpurun_out/r02_sanitizer_memcheck_attn_tc.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention_tcgen05 and (1024 or 333 or 128-128)" > gpurun_out/r02_sanitizer_racecheck_attn_tc.log 2>&1; echo "racecheck attn rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck_attn_tc.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention_tcgen05 and 1024" > gpurun_out/r02_sanitizer_synccheck_attn_tc.log 2>&1; echo "synccheck attn rc=$?"; tail -3 gpurun_out/r02_sanitizer_synccheck_attn_tc.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
