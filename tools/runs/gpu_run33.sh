#!/bin/bash
mkdir -p gpurun_out
AEDIT_PERSIST_MIN_TILES=0 timeout 400 python tools/gemm_table.py --batch 100 --top 22 > gpurun_out/gemm_table_p0.log 2>&1; echo rc=$?
AEDIT_PERSIST_MIN_TILES=296 timeout 400 python tools/gemm_table.py --batch 100 --top 22 > gpurun_out/gemm_table_p1.log 2>&1; echo rc=$?
paste -d'|' <(cut -c1-78 gpurun_out/gemm_table_p0.log | tail -24) <(cut -c48-78 gpurun_out/gemm_table_p1.log | tail -24)
