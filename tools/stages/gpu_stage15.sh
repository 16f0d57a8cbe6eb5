#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout=300 -p no:cacheprovider > gpurun_out/tests15.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -8 gpurun_out/tests15.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_prefill -s 1 -c 1 -o gpurun_out/prof_attn_tc2 python tools/attn_tc_one.py 4096 2>&1 | tail -3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sample_kernel -s 2 -c 1 -o gpurun_out/prof_sampler python tools/sampler_time.py 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 > gpurun_out/bench15.log 2>&1; tail -1 gpurun_out/bench15.log | cut -c1-300
cat gpurun_out/summary.txt
