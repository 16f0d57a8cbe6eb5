#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
ORDER=1,1,2 timeout 300 python tools/pf_debug.py 2>&1 | grep "call\|attn row" | head -8
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/tests21.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/tests21.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 > gpurun_out/bench21.log 2>&1; tail -1 gpurun_out/bench21.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','p50_ttft_ms','gpu_launches')}, d['e2e'])"
cat gpurun_out/summary.txt
