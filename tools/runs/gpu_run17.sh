#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider -x -k "groupnorm or layernorm" > gpurun_out/pytest_gn.log 2>&1; rc=$?; echo "pytest_gn rc=$rc" >> gpurun_out/summary.txt
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_gn.log; cat gpurun_out/summary.txt; exit 0; fi
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py > gpurun_out/microbench11.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
run_bench() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v15_$name.json 2> gpurun_out/bench_v15_$name.err; echo "bench $name rc=$?" >> gpurun_out/summary.txt
}
run_bench gnres AEDIT_X=0
run_bench gn2 AEDIT_GN_FUSED=0
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gn.log gpurun_out/pytest_gpu.log | tail -8
grep -E "groupnorm|layernorm" gpurun_out/microbench11.log
for f in gnres gn2; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v15_$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'e2e', round(j['e2e']['value'],1), 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v15_$f.err').read()[-1200:])
"; done
