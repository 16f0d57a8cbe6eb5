#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/op_breakdown.py 2>&1 | tail -14 | tee gpurun_out/op_breakdown.txt
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/attn_tc_one.py 512 2>&1 | tail -4 | tee gpurun_out/sanitizer_attn.txt
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py -q -k "sampler_greedy or sampler_padded" -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/sanitizer_sampler.txt
