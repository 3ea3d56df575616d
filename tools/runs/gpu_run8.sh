#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py > gpurun_out/microbench6.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --forward-batch 50 --no-cpu-baseline > gpurun_out/bench_v6_fb50.json 2> gpurun_out/bench_v6_fb50.err; echo "bench fb50 rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v6.csv python tools/profile_step.py --forward-batch 25 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -8
grep -E "layernorm|groupnorm|elementwise|attention" gpurun_out/microbench6.log
python -c "
import json
j=json.loads(open('gpurun_out/bench_v6_fb50.json').read().strip().splitlines()[-1]); r=j['roofline']; print('fb50', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
"
