#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python tools/pf_debug.py 2>&1 | tail -3
timeout 300 python tools/sampler_time.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/tests12.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -6 gpurun_out/tests12.log
for sk in 1 0; do
  B2B_STREAMK=$sk timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e > gpurun_out/bench_sk$sk.log 2>&1; echo "sk=$sk: $(tail -1 gpurun_out/bench_sk$sk.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['p50_ttft_ms'])")"
  B2B_STREAMK=$sk timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --batch 1 > gpurun_out/bench_sk${sk}_b1.log 2>&1; echo "sk=$sk b1: $(tail -1 gpurun_out/bench_sk${sk}_b1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['p50_ttft_ms'])")"
done
cat gpurun_out/summary.txt
