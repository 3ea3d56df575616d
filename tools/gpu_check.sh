#!/bin/bash
# The GPU validation sequence of this repo, for `gpurun -- 'bash tools/gpu_check.sh [quick|full|ncu|sanitize|synccheck][-only]'`.
# Everything it writes goes to gpurun_out/ (scratch); copy what should be kept into profiles/.
set -u
mode=${1:-quick}
mkdir -p gpurun_out
if [ "${mode%-only}" = "$mode" ]; then     # "<mode>-only" skips the suite / smoke / bench preamble
  python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -2 gpurun_out/pytest_gpu.log
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
  timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench.json
fi
mode=${mode%-only}
if [ "$mode" = full ]; then
  for c in tango-10s sdedit-30s pc-drift; do
    timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-ends --queue-group 0 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "$c rc=$?"
  done
  AEDIT_OPERANDS=bf16 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_bf16.log 2>&1; echo "bf16 suite rc=$?"
  python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1
  python tools/hbm_kernels.py --json gpurun_out/hbm_kernels.json > gpurun_out/hbm_kernels.log 2>&1
fi
if [ "$mode" = synccheck ]; then tools="synccheck"; mode=sanitize; else tools="memcheck racecheck synccheck"; fi
if [ "$mode" = sanitize ]; then
  for tool in $tools; do
    timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 100000 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pc_drift.py -x -q -k "not large" > gpurun_out/sanitizer_$tool.full 2>&1; echo "$tool rc=$?"
    # keep the summary lines and one line per distinct error site (kernel + source line), not every thread's report
    grep -E "passed|failed|SUMMARY" gpurun_out/sanitizer_$tool.full > gpurun_out/sanitizer_$tool.log
    grep "=========     at " gpurun_out/sanitizer_$tool.full | sed 's/+0x[0-9a-f]* in / in /' | sort | uniq -c >> gpurun_out/sanitizer_$tool.log
    rm -f gpurun_out/sanitizer_$tool.full; tail -6 gpurun_out/sanitizer_$tool.log
  done
fi
if [ "$mode" = ncu ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_B2_B100.csv python tools/profile_step.py --forward-batch 50 > gpurun_out/launches.log 2>&1
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ --launch-skip 60 --launch-count 14 -o gpurun_out/ncu_full_gemm -f python tools/profile_step.py --only chunk --forward-batch 50 > gpurun_out/ncu_gemm.log 2>&1
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_tc --launch-count 3 -o gpurun_out/ncu_full_attn_tc -f python tools/profile_step.py --only chunk --forward-batch 50 > gpurun_out/ncu_attn.log 2>&1
fi
