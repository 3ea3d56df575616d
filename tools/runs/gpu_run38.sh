#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" --tb=short -p no:cacheprovider -x > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_attn.log
for ns in 1 0 2 3 4 6; do
  AEDIT_ATTN_SPLIT=$ns timeout 300 python tools/eval_time.py --B 2 --pdlx 0 2> gpurun_out/et.err | sed "s/^/attn_split=$ns /"; tail -2 gpurun_out/et.err
done | tee gpurun_out/eval_time_attn_split.log
