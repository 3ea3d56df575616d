#!/bin/bash
mkdir -p gpurun_out
for v in 2500,3000000 1500,3000000 500,3000000 0,6000000 1500,6000000 4000,3000000; do
  timeout 300 python tools/eval_time.py --B 2 --pdlx 0 --env ae_set_tile_model_reduce=$v 2> gpurun_out/et.err | sed "s/^/reduce=$v /"; tail -2 gpurun_out/et.err
done | tee gpurun_out/eval_time_reduce_model.log
