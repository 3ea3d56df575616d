#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_k.log
timeout 300 python tools/eval_time.py --B 2 100 --pdlx 0 2> gpurun_out/et.err; tail -2 gpurun_out/et.err
