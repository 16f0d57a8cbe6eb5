#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python tools/op_breakdown.py 2>&1 | tail -14 | tee gpurun_out/op_breakdown2.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/tests20.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/tests20.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 > gpurun_out/bench20.log 2>&1; tail -1 gpurun_out/bench20.log | cut -c1-260
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --batch 1 > gpurun_out/bench20_b1.log 2>&1; tail -1 gpurun_out/bench20_b1.log | cut -c1-200
cat gpurun_out/summary.txt
