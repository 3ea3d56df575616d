#!/bin/bash
# round-2 GPU call 6: tcgen05 attention bring-up (guarded by timeouts: a hung mbarrier wait must not hold the box)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -s -x -k "attention_tcgen05" > gpurun_out/r02_attn_tc_test.log 2>&1; echo "attn tc test rc=$?"
grep -E "rel err|passed|failed|Error|error" gpurun_out/r02_attn_tc_test.log | tail -20
timeout 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench.log 2>&1; echo "attn bench rc=$?"
cat gpurun_out/r02_attn_bench.log | tail -12
