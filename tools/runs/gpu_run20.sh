#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
AEDIT_PDL=1 timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_pdl1.log 2>&1; echo "pytest_gpu pdl1 rc=$?" >> gpurun_out/summary.txt
run_bench() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v18_$name.json 2> gpurun_out/bench_v18_$name.err; echo "bench $name rc=$?" >> gpurun_out/summary.txt
}
run_bench pdl2 AEDIT_PDL=2
run_bench pdl1 AEDIT_PDL=1
run_bench pdl0 AEDIT_PDL=0
timeout 300 env AEDIT_PDL=1 python tools/microbench.py > gpurun_out/microbench13_pdl1.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu_pdl1.log | tail -8
grep -E "groupnorm|layernorm|attention|tiny|alternating" gpurun_out/microbench13_pdl1.log | head -30
for f in pdl2 pdl1 pdl0; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v18_$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'e2e', round(j['e2e']['value'],1), 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v18_$f.err').read()[-1200:])
"; done
