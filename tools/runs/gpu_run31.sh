#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/lanes_ab.py --reps 4 --variants "hr=0" "hr=1" "hr=0" "hr=1,rev=shared" "hr=1,rev=solo" > gpurun_out/lanes_ab5.log 2> gpurun_out/lanes_ab5.err; echo "lanes_ab rc=$?"
cat gpurun_out/lanes_ab5.log; tail -5 gpurun_out/lanes_ab5.err
