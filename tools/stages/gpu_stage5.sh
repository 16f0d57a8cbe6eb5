#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=200 -p no:cacheprovider -x > gpurun_out/tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/tests.log
timeout 300 python tools/gemm_timeline.py > gpurun_out/gemm_timeline2.log 2>&1; cat gpurun_out/gemm_timeline2.log
timeout 600 python tools/layer_sweep.py > gpurun_out/layer_sweep2.log 2>&1; echo "sweep32 exit $?" >> gpurun_out/summary.txt; B=1 timeout 600 python tools/layer_sweep.py >> gpurun_out/layer_sweep2.log 2>&1; echo "sweep1 exit $?" >> gpurun_out/summary.txt
grep "us/layer" gpurun_out/layer_sweep2.log; tail -3 gpurun_out/layer_sweep2.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e > gpurun_out/bench_8b_v3.log 2>&1; tail -1 gpurun_out/bench_8b_v3.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --batch 1 > gpurun_out/bench_8b_b1_v3.log 2>&1; tail -1 gpurun_out/bench_8b_b1_v3.log
cat gpurun_out/summary.txt
