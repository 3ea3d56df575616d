#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r02_pytest_v41.log 2>&1; echo "suite rc=$?"
grep -E "fd step|passed|failed|Error" gpurun_out/r02_pytest_v41.log | tail -12
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_v41.json 2> gpurun_out/r02_bench_v41.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_v41.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "queue", d.get("throughput_queue",{}).get("value"), "ends", d.get("ends",{}).get("per_edit_ms"), "frac", d["roofline"]["frac"], d["roofline"]["frac_in_situ"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_v41.err").read()[-2000:])
PY
