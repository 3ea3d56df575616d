#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
run_bench() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v10_$name.json 2> gpurun_out/bench_v10_$name.err; echo "bench $name rc=$?" >> gpurun_out/summary.txt
}
run_bench fast AEDIT_X=0
run_bench slow AEDIT_FAST_EPILOGUE=0
timeout 600 python tools/gemm_table.py > gpurun_out/gemm_table_v10.log 2>&1; echo "gemm_table rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m tests.gpu_torch_eager_baseline --steps 10 > gpurun_out/torch_eager.log 2>&1; echo "torch_eager rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -8
for f in fast slow; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v10_$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'e2e', round(j['e2e']['value'],1), 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v10_$f.err').read()[-1200:])
"; done
cat gpurun_out/torch_eager.log | cut -c1-250
head -24 gpurun_out/gemm_table_v10.log; grep -A16 "B=100" gpurun_out/gemm_table_v10.log
