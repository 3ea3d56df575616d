#!/bin/bash
# round-2 GPU call 4: multi-clip batched entry points, bench lines for BASELINE configs 2/3/4 geometries
mkdir -p gpurun_out
python -m pytest tests/test_gpu_loops.py -q -s -k "multi_clip" > gpurun_out/r02_multiclip_test.log 2>&1; echo "multiclip test rc=$?"
grep -E "clip [0-9]|passed|failed|Error" gpurun_out/r02_multiclip_test.log | tail
run() { name=$1; shift; timeout 900 python bench.py "$@" --no-cpu-baseline > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "launches", d["gpu_launches"], d.get("roofline",{}).get("per_eval"))
except Exception as e:
    print("$name ERR", e); print(open("gpurun_out/r02_bench_$name.err").read()[-1500:])
PY
}
run tiny_k2 --config tiny --clips 2 --steps 3 --warmup 3
run large_k1 --config audioldm2-large-10s --steps 5 --warmup 3
run large_k4 --config audioldm2-large-10s --clips 4 --steps 3 --warmup 3
run tango_k1 --config tango-10s --steps 3 --warmup 3
run tango_k4 --config tango-10s --clips 4 --steps 3 --warmup 3
run sdedit_30s --config sdedit-30s --steps 3 --warmup 3
run pc_drift --config pc-drift --steps 2 --warmup 3
