#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_v43.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/r02_pytest_v43.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_v43.json 2> gpurun_out/r02_bench_v43.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_v43.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "queue", d.get("throughput_queue",{}).get("value"), "frac", d["roofline"]["frac"], d["roofline"]["frac_in_situ"])
    print("per_eval", {k:(round(v["eval_ms"],2), round(v["gemm_ms"],2)) for k,v in d["roofline"]["per_eval"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_v43.err").read()[-2000:])
PY
for c in tango-10s sdedit-30s; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-ends --queue-group 0 > gpurun_out/r02_bench_v43_$c.json 2> gpurun_out/r02_bench_v43_$c.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_v43_$c.json").read().strip().splitlines()[-1]); print("$c", d["value"], d["ms_per_step"])
except Exception as e: print("$c ERR", e)
PY
done
