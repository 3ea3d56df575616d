#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "geglu or persistent" --tb=short -p no:cacheprovider -x > gpurun_out/pytest_geglu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_geglu.log
timeout 900 python tools/lanes_ab.py --reps 4 --variants "pmt=0" "pmt=296" "pmt=0" "pmt=296" "overlap=0,pmt=0" "overlap=0,pmt=296" > gpurun_out/lanes_ab6.log 2> gpurun_out/lanes_ab6.err; echo "lanes_ab rc=$?"
cat gpurun_out/lanes_ab6.log; tail -5 gpurun_out/lanes_ab6.err
timeout 300 python tools/eval_time.py --B 2 100 --pdlx 0 2> gpurun_out/et.err; tail -2 gpurun_out/et.err
