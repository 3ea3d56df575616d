#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/bimodal_probe.py 2> gpurun_out/bp.err | tee gpurun_out/bimodal_probe.log; tail -3 gpurun_out/bp.err
