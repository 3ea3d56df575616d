#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_fb8.json 2> gpurun_out/bench_fb8.err; echo "bench fb8 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --forward-batch 1 --no-cpu-baseline > gpurun_out/bench_fb1.json 2> gpurun_out/bench_fb1.err; echo "bench fb1 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --forward-batch 25 --no-cpu-baseline > gpurun_out/bench_fb25.json 2> gpurun_out/bench_fb25.err; echo "bench fb25 rc=$?" >> gpurun_out/summary.txt
AEDIT_CUDA_GRAPH=0 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench nograph rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 40 -c 4 -o gpurun_out/prof_gemm_r1 python tools/profile_step.py --only chunk > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
cat gpurun_out/bench_fb8.json; tail -3 gpurun_out/bench_fb8.err
cat gpurun_out/bench_fb1.json gpurun_out/bench_fb25.json gpurun_out/bench_nograph.json
