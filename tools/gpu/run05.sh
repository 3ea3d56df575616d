#!/bin/bash
# round-2 GPU call 5: clip queue (pipelined / grouped), new STFT kernel, int16, full GPU suite, default bench with throughput_queue
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_v38.log 2>&1; echo "suite rc=$?"
tail -3 gpurun_out/r02_pytest_v38.log
python tools/hbm_kernels.py --json gpurun_out/r02_hbm_kernels_v38.json > gpurun_out/r02_hbm_kernels_v38.log 2>&1; tail -2 gpurun_out/r02_hbm_kernels_v38.log
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_v38.json 2> gpurun_out/r02_bench_v38.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_v38.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "queue", d.get("throughput_queue"), "cpu", d.get("cpu_baseline"))
    print("roofline frac", d["roofline"]["frac"], "in situ", d["roofline"]["frac_in_situ"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_v38.err").read()[-2000:])
PY
timeout 600 python bench.py --steps 2 --warmup 3 --queue-group 8 --no-cpu-baseline > gpurun_out/r02_bench_v38_q8.json 2> gpurun_out/r02_bench_v38_q8.err; echo "bench q8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_v38_q8.json").read().strip().splitlines()[-1])
    print("q8", d.get("throughput_queue"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02_bench_v38_q8.err").read()[-2000:])
PY
