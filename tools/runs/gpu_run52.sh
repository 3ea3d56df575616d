#!/bin/bash
mkdir -p gpurun_out
for c in tango-10s audioldm2-30s audioldm-s-5s; do
  timeout 240 python tools/lanes_ab.py --config $c --reps 2 --variants "overlap=0" "overlap=1" 2> gpurun_out/cfg_$c.err | sed "s/^/$c /"; tail -2 gpurun_out/cfg_$c.err
done | tee gpurun_out/other_configs.log
