#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py > gpurun_out/microbench4.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
for fb in 25 50 100; do
timeout 600 python bench.py --steps 2 --warmup 3 --forward-batch $fb --no-cpu-baseline > gpurun_out/bench_v4_fb$fb.json 2> gpurun_out/bench_v4_fb$fb.err; echo "bench fb$fb rc=$?" >> gpurun_out/summary.txt
done
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v4.csv python tools/profile_step.py --forward-batch 25 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error|rel-L2" gpurun_out/pytest_gpu.log | tail -22
grep -E "layernorm|groupnorm|elementwise|attention" gpurun_out/microbench4.log
for f in 25 50 100; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v4_fb$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('fb$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('fb$f ERR', e, open('gpurun_out/bench_v4_fb$f.err').read()[-1200:])
"; done
