#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
AEDIT_PDL=2 timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_loops.py tests/test_gpu_ends.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_pdl2.log 2>&1; echo "pytest_gpu_pdl2 rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py > gpurun_out/microbench7.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --forward-batch 50 --no-cpu-baseline > gpurun_out/bench_v7_fb50.json 2> gpurun_out/bench_v7_fb50.err; echo "bench fb50 rc=$?" >> gpurun_out/summary.txt
AEDIT_PDL=2 timeout 600 python bench.py --steps 2 --warmup 3 --forward-batch 50 --no-cpu-baseline > gpurun_out/bench_v7_fb50_pdl2.json 2> gpurun_out/bench_v7_fb50_pdl2.err; echo "bench fb50 pdl2 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log gpurun_out/pytest_gpu_pdl2.log | tail -8
grep -E "layernorm|groupnorm|elementwise" gpurun_out/microbench7.log
for f in fb50 fb50_pdl2; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v7_$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v7_$f.err').read()[-1200:])
"; done
