#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/lanes_ab.py --reps 4 --variants "head=10" "head=6" "head=16" "head=10,fb=40" "head=10" "head=10,pc=7" "head=10,pc=11" "overlap=0" > gpurun_out/lanes_ab8.log 2> gpurun_out/lanes_ab8.err; echo "lanes_ab rc=$?"
cat gpurun_out/lanes_ab8.log; tail -5 gpurun_out/lanes_ab8.err
