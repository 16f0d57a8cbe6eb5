#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=200 -p no:cacheprovider -x > gpurun_out/tests_pdl.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests_pdl.log
timeout 900 python tools/layer_sweep.py > gpurun_out/layer_sweep.log 2>&1; echo "sweep exit $?" >> gpurun_out/summary.txt
cat gpurun_out/layer_sweep.log | tail -40
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e > gpurun_out/bench_8b_pdl.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_8b_pdl.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --batch 1 > gpurun_out/bench_8b_b1_pdl.log 2>&1
tail -1 gpurun_out/bench_8b_b1_pdl.log
cat gpurun_out/summary.txt
