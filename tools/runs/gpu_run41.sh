#!/bin/bash
mkdir -p gpurun_out
AEDIT_GEMM_MULTICAST=0 timeout 300 python tools/gemm_table.py --batch 2 --top 40 > gpurun_out/gemm_table_mc0.log 2>&1; echo rc=$?
AEDIT_GEMM_MULTICAST=1 timeout 300 python tools/gemm_table.py --batch 2 --top 40 > gpurun_out/gemm_table_mc1.log 2>&1; echo rc=$?
