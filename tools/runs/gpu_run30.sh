#!/bin/bash
mkdir -p gpurun_out
for kb in 0 6 9 12 24; do
  timeout 300 python tools/eval_time.py --B 100 --pdlx 0 --env ae_set_shallow_kblocks=$kb 2> gpurun_out/et.err | sed "s/^/skb=$kb /" ; tail -2 gpurun_out/et.err
done | tee gpurun_out/eval_time_skb.log
timeout 900 python tools/lanes_ab.py --reps 4 --variants "pf=0" "pf=15" "pc=15" "ps=15" "pf=15,pc=15" "pf=0" "skb=9" "skb=24" > gpurun_out/lanes_ab4.log 2> gpurun_out/lanes_ab4.err; echo "lanes_ab rc=$?"
cat gpurun_out/lanes_ab4.log; tail -5 gpurun_out/lanes_ab4.err
