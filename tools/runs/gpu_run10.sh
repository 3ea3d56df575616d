#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py > gpurun_out/microbench8.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
for pdl in 0 2 1; do
AEDIT_PDL=$pdl timeout 600 python bench.py --steps 2 --warmup 3 --forward-batch 50 --no-cpu-baseline > gpurun_out/bench_v8_pdl$pdl.json 2> gpurun_out/bench_v8_pdl$pdl.err; echo "bench pdl$pdl rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -8
grep -E "cluster|ws-split|no split" gpurun_out/microbench8.log
for f in pdl0 pdl2 pdl1; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v8_$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v8_$f.err').read()[-1200:])
"; done
