#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py > gpurun_out/microbench2.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2_fb8.json 2> gpurun_out/bench_v2_fb8.err; echo "bench fb8 rc=$?" >> gpurun_out/summary.txt
AEDIT_PDL=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2_nopdl.json 2> gpurun_out/bench_v2_nopdl.err; echo "bench nopdl rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --forward-batch 25 --no-cpu-baseline > gpurun_out/bench_v2_fb25.json 2> gpurun_out/bench_v2_fb25.err; echo "bench fb25 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -15; tail -2 gpurun_out/smoke.log
cat gpurun_out/microbench2.log
for f in fb8 nopdl fb25; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v2_$f.json').read().strip().splitlines()[-1]); print('$f', j['value'], j['ms_per_step'], j['gpu_launches'], j['roofline']['achieved'], j['roofline']['gemm_share_of_eval_time'], j['roofline']['job_tflops'])
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v2_$f.err').read()[-800:])
"; done
