#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/tests25.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/tests25.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 > gpurun_out/bench25.log 2>&1; tail -1 gpurun_out/bench25.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','p50_ttft_ms')}, d['e2e']['value'], d['roofline'])"
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --batch 1 > gpurun_out/bench25_b1.log 2>&1; tail -1 gpurun_out/bench25_b1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','p50_ttft_ms')})"
for dt in fp8 mxfp8; do timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 --no-e2e --dtype $dt > gpurun_out/bench25_$dt.log 2>&1; tail -1 gpurun_out/bench25_$dt.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$dt', {k:d[k] for k in ('value','ms_per_step','p50_ttft_ms')})"; done
cat gpurun_out/summary.txt
