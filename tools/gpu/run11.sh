#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pc_drift.py -q -s -k "fd_const" > gpurun_out/r02_pc_fd_const.log 2>&1; echo "fd_const rc=$?"
grep -E "fd step|passed|failed|Error|assert" gpurun_out/r02_pc_fd_const.log | tail
for Q in 4 6 8; do timeout 900 python bench.py --steps 2 --warmup 3 --queue-group $Q --no-cpu-baseline --no-ends > gpurun_out/r02_bench_v40_q$Q.json 2> gpurun_out/r02_bench_v40_q$Q.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_v40_q$Q.json").read().strip().splitlines()[-1]); print("Q=$Q value", round(d["value"],1), "queue", d.get("throughput_queue"))
except Exception as e: print("Q=$Q ERR", e); print(open("gpurun_out/r02_bench_v40_q$Q.err").read()[-1500:])
PY
done
