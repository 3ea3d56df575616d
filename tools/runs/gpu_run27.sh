#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/eval_time.py --pdlx 0 1 2 3 4 7 8 11 15 > gpurun_out/eval_time.log 2> gpurun_out/eval_time.err; echo "rc=$?"
cat gpurun_out/eval_time.log; tail -5 gpurun_out/eval_time.err
