#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_kernels.log 2>&1; rc=$?; echo "pytest_kernels rc=$rc" >> gpurun_out/summary.txt
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_kernels.log; cat gpurun_out/summary.txt; exit 0; fi
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x --deselect tests/test_gpu_kernels.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_rest rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py > gpurun_out/microbench10.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
run_bench() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v12_$name.json 2> gpurun_out/bench_v12_$name.err; echo "bench $name rc=$?" >> gpurun_out/summary.txt
}
run_bench fused AEDIT_X=0
run_bench twolaunch AEDIT_FUSED_SPLITK=0
timeout 600 python tools/gemm_table.py --top 16 > gpurun_out/gemm_table_v12.log 2>&1; echo "gemm_table rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_kernels.log gpurun_out/pytest_gpu.log | tail -8
grep -E "split|launch" gpurun_out/microbench10.log
for f in fused twolaunch; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v12_$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'e2e', round(j['e2e']['value'],1), 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v12_$f.err').read()[-1200:])
"; done
grep -A12 "B=2:" gpurun_out/gemm_table_v12.log | cut -c1-100; grep -A12 "B=100:" gpurun_out/gemm_table_v12.log | cut -c1-100
