#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py > gpurun_out/microbench3.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v3_fb8.json 2> gpurun_out/bench_v3_fb8.err; echo "bench fb8 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --forward-batch 25 --no-cpu-baseline > gpurun_out/bench_v3_fb25.json 2> gpurun_out/bench_v3_fb25.err; echo "bench fb25 rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v3.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -15
cat gpurun_out/microbench3.log
for f in fb8 fb25; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v3_$f.json').read().strip().splitlines()[-1]); print('$f', j['value'], j['ms_per_step'], j['gpu_launches']); print(json.dumps(j['roofline']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v3_$f.err').read()[-1200:])
"; done
