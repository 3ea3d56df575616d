#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v34.json 2> gpurun_out/bench_v34.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_v34_reference.json 2> gpurun_out/bench_v34_reference.err; echo "bench ref rc=$?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_v34.csv python tools/profile_step.py --forward-batch 50 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
cut -c1-900 gpurun_out/bench_v34.json; tail -3 gpurun_out/bench_v34.err
cut -c1-600 gpurun_out/bench_v34_reference.json; tail -3 gpurun_out/bench_v34_reference.err
tail -3 gpurun_out/smoke.log
