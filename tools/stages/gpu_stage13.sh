#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python tools/sampler_time.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/tests13.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -6 gpurun_out/tests13.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 > gpurun_out/bench13.log 2>&1; tail -1 gpurun_out/bench13.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --batch 1 > gpurun_out/bench13_b1.log 2>&1; tail -1 gpurun_out/bench13_b1.log | cut -c1-200
cat gpurun_out/summary.txt
