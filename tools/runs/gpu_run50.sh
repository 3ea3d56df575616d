#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/lanes_ab.py --reps 4 --variants "ssm=1,rep=0" "ssm=0,rep=0" "ssm=1,rep=1" "ssm=0,rep=1" > gpurun_out/lanes_ab10.log 2> gpurun_out/lanes_ab10.err; echo "lanes_ab rc=$?"
cat gpurun_out/lanes_ab10.log; tail -5 gpurun_out/lanes_ab10.err
