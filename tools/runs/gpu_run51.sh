#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v35.json 2> gpurun_out/bench_v35.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ --launch-skip 60 --launch-count 14 -o gpurun_out/prof_v35 -f python tools/profile_step.py --only chunk --forward-batch 50 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -4 gpurun_out/pytest_gpu.log
cut -c1-700 gpurun_out/bench_v35.json; tail -3 gpurun_out/bench_v35.err
tail -2 gpurun_out/smoke.log
