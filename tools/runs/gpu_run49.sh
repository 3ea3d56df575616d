#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/lanes_ab.py --reps 4 --variants "gpt=1,rep=0" "gpt=4,rep=0" "gpt=1,rep=1" "gpt=4,rep=1" "gpt=1,rep=2" "gpt=4,rep=2" > gpurun_out/lanes_ab9.log 2> gpurun_out/lanes_ab9.err; echo "lanes_ab rc=$?"
cat gpurun_out/lanes_ab9.log; tail -5 gpurun_out/lanes_ab9.err
