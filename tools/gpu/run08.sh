#!/bin/bash
# round-2 GPU call 8: attention fix + suite + ncu evidence (launch list of one B=2 eval + one B=100 chunk; --set full of the
# GEMM family and of the tcgen05 attention kernel)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -s -x -k "attention" > gpurun_out/r02_attn_tc_test.log 2>&1; echo "attn tests rc=$?"
grep -E "passed|failed|Error|error" gpurun_out/r02_attn_tc_test.log | tail -4
timeout 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_v3.log 2>&1; cat gpurun_out/r02_attn_bench_v3.log | tail -10
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_v40.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/r02_pytest_v40.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_v40_B2_B100.csv python tools/profile_step.py --forward-batch 50 > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ --launch-skip 60 --launch-count 14 -o gpurun_out/r02_ncu_full_gemm_v40 -f python tools/profile_step.py --only chunk --forward-batch 50 > gpurun_out/r02_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_tc --launch-count 3 -o gpurun_out/r02_ncu_full_attn_tc_v40 -f python tools/profile_step.py --only chunk --forward-batch 50 > gpurun_out/r02_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
