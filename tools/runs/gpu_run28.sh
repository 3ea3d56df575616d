#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "column_statistics or groupnorm" --tb=short -p no:cacheprovider > gpurun_out/pytest_cs.log 2>&1; echo "pytest_cs rc=$?"
tail -30 gpurun_out/pytest_cs.log
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_loops.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_unet.log 2>&1; echo "pytest_unet rc=$?"
tail -8 gpurun_out/pytest_unet.log
timeout 600 python tools/eval_time.py --pdlx 0 > gpurun_out/eval_time_cs.log 2> gpurun_out/eval_time_cs.err; echo "rc=$?"
cat gpurun_out/eval_time_cs.log; tail -5 gpurun_out/eval_time_cs.err
AEDIT_GN_COLSTATS=0 timeout 600 python tools/eval_time.py --pdlx 0 > gpurun_out/eval_time_nocs.log 2> gpurun_out/eval_time_nocs.err; echo "rc=$?"
cat gpurun_out/eval_time_nocs.log; tail -5 gpurun_out/eval_time_nocs.err
