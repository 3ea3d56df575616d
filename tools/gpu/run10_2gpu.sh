#!/bin/bash
# round-2 GPU call 10 (2 GPUs): NCCL checks of the sharded paths + torchrun bench lines
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r02_dist_check_2gpu.log 2>&1; echo "dist check rc=$?"
grep -E "world|Error|error" gpurun_out/r02_dist_check_2gpu.log | tail -5
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --config pc-drift --steps 2 --warmup 3 > gpurun_out/r02_bench_pc_drift_2gpu.json 2> gpurun_out/r02_bench_pc_drift_2gpu.err; echo "pc 2gpu rc=$?"
tail -c 700 gpurun_out/r02_bench_pc_drift_2gpu.json
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_v40_2gpu.json 2> gpurun_out/r02_bench_v40_2gpu.err; echo "bench 2gpu rc=$?"
python - <<'PY'
import json
for n in ("r02_bench_v40_2gpu","r02_bench_pc_drift_2gpu"):
    try:
        d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1]); print(n, d["value"], d["n_gpus"], d["scaling"], d["ms_per_step"])
    except Exception as e: print(n,"ERR",e); print(open(f"gpurun_out/{n}.err").read()[-1500:])
PY
