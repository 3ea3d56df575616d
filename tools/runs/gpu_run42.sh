#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_unet.py tests/test_gpu_loops.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_fold.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -15 gpurun_out/pytest_fold.log
for f in 0 1; do
  AEDIT_FOLD_CROSS_ATTN=$f timeout 300 python tools/eval_time.py --B 2 100 --pdlx 0 2> gpurun_out/et.err | sed "s/^/fold=$f /"; tail -2 gpurun_out/et.err
done | tee gpurun_out/eval_time_fold.log
