#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/lanes_ab.py --reps 4 --variants "head=10" "pdlx=4" "head=10" "pdlx=15" "head=10" "pdlx=3" "head=10,fb=40" "head=10,fb=67" "head=10,fb=100" > gpurun_out/lanes_ab3.log 2> gpurun_out/lanes_ab3.err; echo "lanes_ab rc=$?"
cat gpurun_out/lanes_ab3.log; tail -5 gpurun_out/lanes_ab3.err
