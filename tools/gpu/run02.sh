#!/bin/bash
# round-2 GPU call 2: fp16-operand build (default) vs bf16-operand build: whole GPU suite, full-size parity, bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -x > gpurun_out/r02_pytest_fp16.log 2>&1; echo "fp16 suite rc=$?"
grep -E "rel-L2|passed|failed|Error" gpurun_out/r02_pytest_fp16.log | tail -40
AEDIT_OPERANDS=bf16 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_bf16.log 2>&1; echo "bf16 suite rc=$?"
grep -E "passed|failed" gpurun_out/r02_pytest_bf16.log | tail -5
python tools/error_attribution.py audioldm2-large 501 > gpurun_out/r02_error_attribution_audioldm2_large_fp16.log 2>&1
tail -2 gpurun_out/r02_error_attribution_audioldm2_large_fp16.log
python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_fp16.json 2> gpurun_out/r02_bench_fp16.err; echo "bench fp16 rc=$?"
AEDIT_OPERANDS=bf16 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_bf16.json 2> gpurun_out/r02_bench_bf16.err; echo "bench bf16 rc=$?"
python - <<'PY'
import json
for n in ("fp16","bf16"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"])
    except Exception as e: print(n, "ERR", e)
PY
