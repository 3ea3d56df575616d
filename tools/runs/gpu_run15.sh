#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_v13_default.json 2> gpurun_out/bench_v13_default.err; echo "bench default rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_v13_reference.json 2> gpurun_out/bench_v13_reference.err; echo "bench reference rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_v13.csv python tools/profile_step.py --forward-batch 50 > gpurun_out/ncu_list_v13.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 --launch-skip 60 --launch-count 14 -o gpurun_out/prof_gemm_v13 python tools/profile_step.py --only chunk --forward-batch 50 > gpurun_out/ncu_full_v13.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -8
cat gpurun_out/bench_v13_default.json | cut -c1-1500; tail -3 gpurun_out/bench_v13_default.err
cat gpurun_out/bench_v13_reference.json | cut -c1-800
tail -2 gpurun_out/smoke.log
