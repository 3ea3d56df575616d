#!/bin/bash
# N-GPU validation, for `gpurun --gpus N -- 'bash tools/multi_gpu_check.sh N'`: NCCL protocol check, then the bench lines of
# the sharded geometries (BASELINE configs 2-4 + the strong-scaling mode).  Output: gpurun_out/ (scratch).
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1; echo "dist check rc=$?"; tail -4 gpurun_out/dist_check_${N}gpu.log
B="--gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-ends --queue-group 0"
i=0
for args in "--shard-forward" "--config tango-10s --clips 4" "--config pc-drift" "--config sdedit-30s"; do
  i=$((i+1)); tag=$(echo $args | tr -d '-' | tr ' ' '_')
  timeout 900 $TR --master-port $((29520+i)) bench.py $B $args > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.err; echo "$tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${N}gpu_$tag.json").read().strip().splitlines()[-1]); print("  ", d["config"]["workload"], d["scaling"], round(d["value"],1), "steps/s", round(d["ms_per_step"],1), "ms", d["details"]["parallelism"])
except Exception as e: print("  ERR", e)
PY
done
