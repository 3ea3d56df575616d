#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "persistent" --tb=short -p no:cacheprovider -x > gpurun_out/pytest_persist.log 2>&1; echo "pytest_persist rc=$?"
tail -25 gpurun_out/pytest_persist.log
for mt in 0 296; do
  timeout 300 python tools/eval_time.py --B 2 100 --pdlx 0 --env ae_set_persistent_min_tiles=$mt 2> gpurun_out/et.err | sed "s/^/min_tiles=$mt /" ; tail -2 gpurun_out/et.err
done | tee gpurun_out/eval_time_persist.log
