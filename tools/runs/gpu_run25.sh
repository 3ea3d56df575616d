#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/kernel_share.py > gpurun_out/kernel_share.log 2> gpurun_out/kernel_share.err; echo "share rc=$?"
cat gpurun_out/kernel_share.log; tail -8 gpurun_out/kernel_share.err
