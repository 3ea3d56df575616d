#!/bin/bash
# run with: gpurun --gpus 2 -- 'bash tools/gpu_run_2gpu.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary2.txt
nvidia-smi -L > gpurun_out/smi2.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check.log 2>&1; echo "dist_check rc=$?" >> gpurun_out/summary2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu rc=$?" >> gpurun_out/summary2.txt
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_samebox.json 2> gpurun_out/bench_1gpu_samebox.err; echo "bench 1gpu rc=$?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt
tail -5 gpurun_out/dist_check.log
tail -2 gpurun_out/bench_2gpu.json | cut -c1-700
tail -3 gpurun_out/bench_2gpu.err
tail -1 gpurun_out/bench_1gpu_samebox.json | cut -c1-300
