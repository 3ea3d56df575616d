#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/eval_time.py --B 2 --pdlx 0 2 4 6 8 0 2> gpurun_out/et.err | tee gpurun_out/eval_time_pdlx_v34.log; tail -2 gpurun_out/et.err
