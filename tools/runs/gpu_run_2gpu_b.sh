#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/dist_check.log | tail -12 | cut -c1-300
