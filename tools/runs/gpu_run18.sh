#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gn_stats|gn_apply" --launch-skip 20 --launch-count 8 -o gpurun_out/prof_gn_b2 python tools/profile_step.py --only b2 > gpurun_out/ncu_gn.log 2>&1; echo "ncu gn rc=$?" >> gpurun_out/summary.txt
run_bench() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v16_$name.json 2> gpurun_out/bench_v16_$name.err; echo "bench $name rc=$?" >> gpurun_out/summary.txt
}
run_bench base AEDIT_X=0
timeout 600 python tools/gemm_table.py --batch 2 --top 30 > gpurun_out/gemm_table_v16.log 2>&1; echo "gemm_table rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in base; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v16_$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'e2e', round(j['e2e']['value'],1), 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v16_$f.err').read()[-1200:])
"; done
head -20 gpurun_out/gemm_table_v16.log | cut -c1-100
