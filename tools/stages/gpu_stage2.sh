#!/bin/bash
# Llama-3-8B on one B200: bench (ours), launch list + ncu of the top kernels, reference arm
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_model_gpu.py -q -k engine --timeout=200 -p no:cacheprovider > gpurun_out/test_engine.log 2>&1; echo "test_engine exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --gpus 1 --steps 64 --warmup 8 > gpurun_out/bench_8b.log 2>&1; echo "bench_8b exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_8b.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --batch 1 --no-e2e > gpurun_out/bench_8b_b1.log 2>&1; echo "bench_8b_b1 exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_8b_b1.log
# per-launch device times of one decode step (eager, no graphs): shares, not absolutes
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv \
   env B2B_STEPS=1 python tools/decode_eager.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu_launches exit $?" >> gpurun_out/summary.txt
# full capture of the GEMM kernel (gate/up + down shapes) and attention
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -c 5 -o gpurun_out/prof_gemm \
   env B2B_STEPS=1 python tools/decode_eager.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu_gemm exit $?" >> gpurun_out/summary.txt
timeout 1500 python bench.py --impl reference --gpus 1 --steps 64 --warmup 8 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_ref.log
cat gpurun_out/summary.txt
