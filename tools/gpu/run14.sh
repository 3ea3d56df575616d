#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r02_pytest_v42.log 2>&1; echo "suite rc=$?"
grep -E "batched vocoder|fd step|passed|failed|Error" gpurun_out/r02_pytest_v42.log | tail -12
python tools/hbm_kernels.py --json gpurun_out/r02_hbm_kernels_v42.json > gpurun_out/r02_hbm_kernels_v42.log 2>&1; grep -E "gn_apply|gn stats|layernorm_kernel \[102" gpurun_out/r02_hbm_kernels_v42.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v42.json 2> gpurun_out/r02_bench_v42.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_v42.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "queue", d.get("throughput_queue",{}).get("value"), "frac", d["roofline"]["frac"], d["roofline"]["frac_in_situ"])
    print("per_eval", {k:(round(v["eval_ms"],2)) for k,v in d["roofline"]["per_eval"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_v42.err").read()[-2000:])
PY
