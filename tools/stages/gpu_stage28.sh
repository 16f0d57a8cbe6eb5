#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --model gemma-2-2b --prompt-len 4096 --steps 32 --warmup 4 --batch 4 --no-e2e > gpurun_out/bench28_gemma.log 2>&1; tail -1 gpurun_out/bench28_gemma.log | cut -c1-700
timeout 400 python tools/load_test.py --spawn zephyr-7b-beta --clients 32 --requests 2 --max-new-tokens 64 --max-batch 32 --max-seq-len 1024 > gpurun_out/load28.log 2>&1; tail -1 gpurun_out/load28.log
