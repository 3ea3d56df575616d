#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py > gpurun_out/microbench9.log 2>&1; echo "microbench rc=$?" >> gpurun_out/summary.txt
run_bench() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v9_$name.json 2> gpurun_out/bench_v9_$name.err; echo "bench $name rc=$?" >> gpurun_out/summary.txt
}
run_bench base AEDIT_X=0
run_bench gn0 AEDIT_GN_FUSED=0
run_bench dual148 AEDIT_DUAL_STREAM=1
run_bench dual74 AEDIT_DUAL_STREAM=1 AEDIT_SPLITK_CTAS=74
timeout 600 python tools/gemm_table.py > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m tests.gpu_torch_eager_baseline --steps 10 > gpurun_out/torch_eager.log 2>&1; echo "torch_eager rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 --launch-skip 60 --launch-count 14 -o gpurun_out/prof_gemm_v9 python tools/profile_step.py --only chunk --forward-batch 50 > gpurun_out/ncu_full_v9.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -8
grep -E "groupnorm" gpurun_out/microbench9.log
for f in base gn0 dual148 dual74; do python -c "
import json
try:
    j=json.loads(open('gpurun_out/bench_v9_$f.json').read().strip().splitlines()[-1]); r=j['roofline']; print('$f', round(j['value'],1), round(j['ms_per_step'],1), j['gpu_launches'], 'e2e', round(j['e2e']['value'],1), 'achieved', round(r['achieved'],1), json.dumps(r['per_eval']))
except Exception as e: print('$f ERR', e, open('gpurun_out/bench_v9_$f.err').read()[-1200:])
"; done
cat gpurun_out/torch_eager.log | cut -c1-250
head -30 gpurun_out/gemm_table.log
