#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "multicast" --tb=short -p no:cacheprovider -x > gpurun_out/pytest_mc.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -15 gpurun_out/pytest_mc.log
if [ $rc -eq 0 ]; then
for mc in 0 1; do
  AEDIT_GEMM_MULTICAST=$mc timeout 200 python tools/eval_time.py --B 2 --pdlx 0 2> gpurun_out/et.err | sed "s/^/multicast=$mc /"; tail -2 gpurun_out/et.err
done | tee gpurun_out/eval_time_multicast.log
fi
nvidia-smi --query-gpu=name,utilization.gpu --format=csv | tail -1
