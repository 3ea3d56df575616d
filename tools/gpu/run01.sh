#!/bin/bash
# round-2 GPU call 1: full-size parity at the benchmarked geometries, per-block error attribution, HBM-bound kernel
# bandwidths (+ ncu dram bytes), compute-sanitizer over the kernel unit tests.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize.py -x -q -s > gpurun_out/r02_fullsize.log 2>&1; echo "fullsize rc=$?" 
python -m pytest tests/test_gpu_fullsize.py -q -s >> gpurun_out/r02_fullsize.log 2>&1
tail -5 gpurun_out/r02_fullsize.log
python tools/error_attribution.py audioldm2-large 501 > gpurun_out/r02_error_attribution_audioldm2_large.log 2>&1
python tools/error_attribution.py tango 501 > gpurun_out/r02_error_attribution_tango.log 2>&1
tail -3 gpurun_out/r02_error_attribution_audioldm2_large.log
python tools/hbm_kernels.py --json gpurun_out/r02_hbm_kernels.json > gpurun_out/r02_hbm_kernels.log 2>&1; echo "hbm rc=$?"
cat gpurun_out/r02_hbm_kernels.log | tail -20
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r02_hbm_kernels_ncu.csv python tools/hbm_kernels.py --iters 2 > gpurun_out/r02_hbm_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "not large" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r02_sanitizer_memcheck.log
