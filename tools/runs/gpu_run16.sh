#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
run_bench() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v14_$name.json 2> gpurun_out/bench_v14_$name.err; echo "bench $name rc=$?" >> gpurun_out/summary.txt
}
run_bench model AEDIT_X=0
run_bench nomodel AEDIT_TILE_MODEL=0
timeout 600 python tools/gemm_table.py --batch 2 --top 30 > gpurun_out/gemm_table_v14.log 2>&1; echo "gemm_table rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -8
for f in model nomodel; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v14_$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'e2e', round(j['e2e']['value'],1), 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v14_$f.err').read()[-1200:])
"; done
head -34 gpurun_out/gemm_table_v14.log | cut -c1-100
