#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/tests17.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -8 gpurun_out/tests17.log
for dt in mxfp8 fp8; do
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --dtype $dt > gpurun_out/bench17_$dt.log 2>&1; tail -1 gpurun_out/bench17_$dt.log | cut -c1-250
done
cat gpurun_out/summary.txt
