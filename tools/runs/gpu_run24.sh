#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/contention_probe.py > gpurun_out/contention_probe.log 2> gpurun_out/contention_probe.err; echo "probe rc=$?"
cat gpurun_out/contention_probe.log; tail -8 gpurun_out/contention_probe.err
