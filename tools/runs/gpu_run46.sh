#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_bench.py 2> gpurun_out/ab.err | tee gpurun_out/attn_bench.log; tail -3 gpurun_out/ab.err
