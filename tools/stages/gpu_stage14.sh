#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 -p no:cacheprovider > gpurun_out/tests14.log 2>&1; echo "tests exit $?" >> gpurun_out/summary.txt; tail -6 gpurun_out/tests14.log
timeout 600 python bench.py --gpus 1 --steps 64 --warmup 8 > gpurun_out/bench14.log 2>&1; tail -1 gpurun_out/bench14.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 8 > gpurun_out/bench14_n2.log 2>&1; tail -1 gpurun_out/bench14_n2.log | cut -c1-400
cat gpurun_out/summary.txt
