#!/bin/bash
# round-2 GPU call 7: tcgen05 attention v2 (O in TMEM, conditional rescale) + whole suite + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -s -x -k "attention" > gpurun_out/r02_attn_tc_test.log 2>&1; echo "attn tests rc=$?"
grep -E "rel err|passed|failed|Error|error" gpurun_out/r02_attn_tc_test.log | tail -14
timeout 300 python tools/attn_bench.py > gpurun_out/r02_attn_bench_v2.log 2>&1; echo "attn bench rc=$?"
cat gpurun_out/r02_attn_bench_v2.log | tail -10
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_v39.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/r02_pytest_v39.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_v39.json 2> gpurun_out/r02_bench_v39.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_v39.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "queue", d.get("throughput_queue",{}).get("value"), "ends", d.get("ends"))
    print("per_eval", {k:(round(v["eval_ms"],2)) for k,v in d["roofline"]["per_eval"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_v39.err").read()[-2000:])
PY
for c in tango-10s sdedit-30s; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-ends --queue-group 0 > gpurun_out/r02_bench_v39_$c.json 2> gpurun_out/r02_bench_v39_$c.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_v39_$c.json").read().strip().splitlines()[-1]); print("$c", d["value"], d["ms_per_step"])
except Exception as e: print("$c ERR", e)
PY
done
