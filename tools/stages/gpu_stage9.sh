#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/tests9.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/tests9.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 > gpurun_out/bench_8b_v4.log 2>&1; tail -1 gpurun_out/bench_8b_v4.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --batch 1 > gpurun_out/bench_8b_b1_v4.log 2>&1; tail -1 gpurun_out/bench_8b_b1_v4.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --batch 64 > gpurun_out/bench_8b_b64_v4.log 2>&1; tail -1 gpurun_out/bench_8b_b64_v4.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_v4.csv \
   env B2B_STEPS=1 python tools/decode_eager.py > gpurun_out/ncu_launches_v4.log 2>&1; echo "ncu_launches exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
