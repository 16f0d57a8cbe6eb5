#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python tools/pf_debug.py > gpurun_out/pf_debug.log 2>&1; tail -6 gpurun_out/pf_debug.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q --timeout=300 -p no:cacheprovider -x -k "streamk or both_kernels" > gpurun_out/tests_sk.log 2>&1; echo "sk tests exit $?" >> gpurun_out/summary.txt; tail -15 gpurun_out/tests_sk.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/tests10.log 2>&1; echo "all tests exit $?" >> gpurun_out/summary.txt; tail -8 gpurun_out/tests10.log
for sk in 1 0; do
  B2B_STREAMK=$sk timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e > gpurun_out/bench_sk$sk.log 2>&1; echo "sk=$sk: $(tail -1 gpurun_out/bench_sk$sk.log | cut -c1-170)"
  B2B_STREAMK=$sk timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --batch 1 > gpurun_out/bench_sk${sk}_b1.log 2>&1; echo "sk=$sk b1: $(tail -1 gpurun_out/bench_sk${sk}_b1.log | cut -c1-170)"
done
cat gpurun_out/summary.txt
